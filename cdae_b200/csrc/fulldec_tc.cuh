// cdae_b200/csrc/fulldec_tc.cuh — full-item-decode TRAINING on the
// 5th-generation tensor cores.  The reference scores positives + sampled negatives
// (cdae.hpp:225-293); with the output set = ALL items the same per-output arithmetic
//     y = W'[i].z_u + b'[i]      g = l'(y, t_ui)      hg_u += g W'[i]      gW'[i] += g z_u
// is three dense contractions over a minibatch of B users (6*B*I*K flops):
//     S  = Z  W'^T   (B x I)     G = l'(S + b', T)            fd_score_kernel
//     HG = G  W'     (B x K)                                   fd_gemm_kernel<KB, false>
//     gW'= G^T Z     (I x K)   gb' = G^T 1                     fd_gemm_kernel<KB, true>
// Operands are bf16 (fp32 accumulation in TMEM), laid out as for the recommend path
// (topn_tc.cuh): Zb [B_pad][Kp], Wb [I_pad][Kp], Kp = round_up(K+2, 64), column K / K+1 of Zb
// hold 1 and of Wb the bf16 hi / lo halves of b' — so the bias rides in the first contraction
// and column K of the third one IS the output-bias gradient.  G is kept in bf16 [B_pad][I_pad]
// and is written once and read twice; each kernel overlaps that HBM stream with its MMAs.
//
// MN-major operands.  The second and third contraction read G and the packed tables ACROSS their
// storage order (the contraction index is the slow one in memory), which tcgen05 supports for
// 16-bit types through MN-major shared-memory descriptors: a TMA box of {64 elements (128 bytes) x
// 64 rows}, 128-byte swizzled, is a [kdim = 64][MN = 64] slab; the descriptor's stride byte offset
// (1024) steps over groups of eight kdim rows and its leading byte offset (8192 = one box) over
// 64-wide MN blocks (cute::UMMA canonical layout "((8,n),(8,k)):((1,LBO),(8,SBO))" in 16-byte units).
#pragma once
#include "topn_tc.cuh"

namespace cdae {
namespace fd {

using namespace tc;   // PTX wrappers, TILE_U = 128, TILE_I = 256, KBLK = 64

constexpr int SC_STAGE = 3;                           // score kernel: W' k-blocks in flight
constexpr int SC_STG_BYTES = 32 * 128;                // score kernel: one warp's staging tile
constexpr int GM_STAGE = 4;                           // gemm kernels: contraction steps in flight
constexpr int GM_A_BYTES = 128 * KBLK * 2;            // 16 KB: 128 (M) x 64 (kdim), either major
constexpr int GM_BOX_BYTES = 64 * KBLK * 2;           // 8 KB : one {64, 64} box
constexpr int MAX_KB = 4;                             // Kp <= 256: K <= 256 (bias outside the contraction when K + 2 > Kp)

__host__ __device__ constexpr size_t score_smem(int kb) {
  return 1024 + (size_t)kb * A_BLK_BYTES + (size_t)SC_STAGE * B_BLK_BYTES + 16 * SC_STG_BYTES + 256;
}
__host__ __device__ constexpr size_t gemm_smem(int kb) {
  return 1024 + (size_t)GM_STAGE * (GM_A_BYTES + (size_t)kb * GM_BOX_BYTES) + GM_BOX_BYTES /*ones slab*/ + 256;
}
__host__ __device__ constexpr int tmem_cols(int kb) { return kb == 1 ? 64 : kb == 2 ? 128 : 256; }

// MN-major operand descriptor, SWIZZLE_128B: see the header comment.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_major(int M, int N, int a_mn, int b_mn) {
  return umma_idesc_bf16(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

// ---------------------------------------------------------------------------------------
// Zb[r][0..K) = bf16(Z[r]), Zb[r][K] = Zb[r][K+1] = 1, rest 0; rows >= n are 0.  Z is the
// minibatch-local hidden matrix [n][ld].  One thread per 8 columns.
// bias_cols = 0: no bias / ones columns (K is a multiple of 64 or one short of it, e.g. config E's
// K = 256: the bias is then added in the score epilogue and its gradient comes from a separate
// ones-operand MMA in the item-gradient kernel).
__global__ void __launch_bounds__(256) pack_z_train_kernel(const float* __restrict__ Z, int n, int64_t n_pad, int K,
                                                           int ld, int Kp, int bias_cols,
                                                           __nv_bfloat16* __restrict__ out) {
  const int g8 = Kp / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pad * g8) return;
  const int64_t r = idx / g8;
  const int c0 = (int)(idx % g8) * 8;
  __align__(16) __nv_bfloat16 o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    float v = 0.f;
    if (r < n) v = c < K ? Z[r * ld + c] : ((bias_cols && c <= K + 1) ? 1.f : 0.f);
    o[j] = __float2bfloat16_rn(v);
  }
  *reinterpret_cast<uint4*>(out + r * Kp + c0) = *reinterpret_cast<const uint4*>(o);
}

// Wb for the bias-outside mode: plain bf16 copy of W' (pad rows / columns 0) and b' zero-padded to
// I_pad floats for the score epilogue.  One thread per 8 columns.
__global__ void __launch_bounds__(256) pack_w_plain_kernel(const float* __restrict__ W, const float* __restrict__ bp,
                                                           int64_t I, int64_t I_pad, int K, int ld, int Kp,
                                                           __nv_bfloat16* __restrict__ out, float* __restrict__ bias_pad) {
  const int g8 = Kp / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= I_pad * g8) return;
  const int64_t r = idx / g8;
  const int c0 = (int)(idx % g8) * 8;
  __align__(16) __nv_bfloat16 o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    o[j] = __float2bfloat16_rn((r < I && c < K) ? W[r * ld + c] : 0.f);
  }
  *reinterpret_cast<uint4*>(out + r * Kp + c0) = *reinterpret_cast<const uint4*>(o);
  if (c0 == 0) bias_pad[r] = r < I ? bp[r] : 0.f;
}

// Target bitmap of a minibatch slice: bit i of row r <=> item i is in the train row of user uids[r]
// (t = 1, cdae.hpp:225-228).  bits [n_pad][words] must be zero; one warp per user.
__global__ void __launch_bounds__(256) fd_bitmap_kernel(const int32_t* __restrict__ uids, int n,
                                                        const int64_t* __restrict__ row_ptr,
                                                        const int32_t* __restrict__ col, int64_t words,
                                                        uint32_t* __restrict__ bits) {
  const int r = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const int64_t uid = uids[r];
  const int64_t p0 = row_ptr[uid], p1 = row_ptr[uid + 1];
  uint32_t* row = bits + (int64_t)r * words;
  for (int64_t p = p0 + lane; p < p1; p += 32) {
    const int it = __ldg(col + p);
    atomicOr(row + (it >> 5), 1u << (it & 31));
  }
}

// ---------------------------------------------------------------------------------------
// The epilogue is bound by issue slots and by the FMA and SFU pipes together, so sigma is computed
// two ways, on alternating pairs of scores (FD_SIGMOID below; 598 TFLOP/s against 537 polynomial-only
// and 558 SFU-only at config C).
// sigma(y) for two scores at once on the FMA pipe: a = e^-|y| (one MUFU.EX2 each), then
// 1/(1+a) on [0,1] as a degree-6 polynomial (Chebyshev fit, |rel err| < 9e-6 evaluated in fp32)
// with packed fp32x2 FMAs; sigma(-|y|) = a/(1+a), sigma(|y|) = 1 - sigma(-|y|).
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void sigmoid2(float y0, float y1, float& s0, float& s1) {
  const float a0 = ex2_approx(-1.4426950408889634f * fabsf(y0));
  const float a1 = ex2_approx(-1.4426950408889634f * fabsf(y1));
  const uint64_t a = pack2(a0, a1);
  uint64_t p = pack2(0.07170680165290833f, 0.07170680165290833f);
  p = fma2(p, a, pack2(-0.32268059253692627f, -0.32268059253692627f));
  p = fma2(p, a, pack2(0.6677695512771606f, 0.6677695512771606f));
  p = fma2(p, a, pack2(-0.9030575156211853f, -0.9030575156211853f));
  p = fma2(p, a, pack2(0.9854083061218262f, 0.9854083061218262f));
  p = fma2(p, a, pack2(-0.9991334080696106f, -0.9991334080696106f));
  p = fma2(p, a, pack2(0.999991238117218f, 0.999991238117218f));
  float q0, q1;
  unpack2(mul2(p, a), q0, q1);          // sigma(-|y|) in (0, 0.5]
  // (a comparison-free form, copysign(0.5 - q, y) + (0.5 - t), needs fewer issue slots but was not
  // faster — the kernel is not issue-bound — and loses the relative precision of small sigma)
  s0 = y0 >= 0.f ? 1.f - q0 : q0;
  s1 = y1 >= 0.f ? 1.f - q1 : q1;
}
// sigma(y) with two SFU operations and nothing else: 1 / (1 + 2^(-y log2 e)).  ex2 saturates to
// +inf / 0 at the ends and rcp(inf) = 0, so no clamp, no |y|, no select.
__device__ __forceinline__ float sigmoid_sfu(float y) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + ex2_approx(-1.4426950408889634f * y)));
  return r;
}
#ifndef FD_EPI
#define FD_EPI 0       // fd_score_kernel epilogue loop: 0 one 32-column TMEM load at a time; 1 the same with the waits moved under
                       // the arithmetic; 2 16-column loads, one always in flight; 3 = 2 + next tile / target words fetched ahead
                       // (A/B: CDAE_NVCC_FLAGS=-DFD_EPI=n)
#endif
#ifndef FD_SIGMOID
#define FD_SIGMOID 2   // 0: polynomial on the FMA pipe, 1: SFU reciprocal, 2: alternate pairs, 3: 3 of 4 pairs on the SFU, 4: 1 of 4
                       // (A/B: CDAE_NVCC_FLAGS=-DFD_SIGMOID=n)
#endif
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

struct ScoreArgs {
  int n_users;               // valid rows of Zb (this rank's slice of the minibatch)
  int64_t I, I_pad;          // valid items; leading dimension of G (multiple of 256)
  int n_tiles;               // I_pad / 256
  int tiles_per_split;       // gridDim.y CTAs share the item tiles of one user tile
  int ksteps;                // ceil((K + 2) / 16): MMA k-steps that hold non-zero operands
  const uint32_t* bits;      // [B_pad][I_pad / 32] target bitmap of the slice (fd_bitmap_kernel)
  const float* bias;         // [I_pad] b' zero-padded, bias-outside mode only (BOUT)
  __nv_bfloat16* G;          // [B_pad][I_pad] loss gradients dl/dy, 0 in pad rows / pad columns
  unsigned long long* outputs;  // stats: scored outputs
};

// One 32-column chunk: y -> g = l'(y, t) -> bf16 -> 64 bytes of this thread's row of the warp's
// staging tile ([32 rows][128 bytes], 128-byte swizzled like every TMA tile: 16-byte piece c of
// row r lives at piece c ^ (r & 7)); cbase = 0 / 4 selects the half of the row.
template <int LT, bool BOUT = false>
__device__ __forceinline__ void grad_compute(const uint32_t (&v)[32], uint32_t pos, uint32_t valid, bool row_ok,
                                             uint32_t (&o)[16], const float* bias = nullptr) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float y0 = __uint_as_float(v[j]), y1 = __uint_as_float(v[j + 1]);
    if (BOUT) {                      // b' outside the contraction: the same 8 bytes for every lane (L1 broadcast)
      const float2 b = __ldg(reinterpret_cast<const float2*>(bias + j));
      y0 += b.x;
      y1 += b.y;
    }
    // target as a float without a conversion: bit j of pos -> 0x3f800000 (1.0f) or 0
    const float t0 = __uint_as_float(((pos >> j) & 1u) * 0x3f800000u);
    const float t1 = __uint_as_float(((pos >> (j + 1)) & 1u) * 0x3f800000u);
    float g0, g1;
    if (LT == LOSS_CE) {             // loss.hpp:141-147: sigma(y) - t
      if (FD_SIGMOID == 1 || (FD_SIGMOID == 2 && (j & 2)) || (FD_SIGMOID == 3 && (j & 6)) || (FD_SIGMOID == 4 && !(j & 6))) {
        g0 = sigmoid_sfu(y0);
        g1 = sigmoid_sfu(y1);
      } else {
        sigmoid2(y0, y1, g0, g1);
      }
      g0 -= t0;
      g1 -= t1;
    } else {                         // SQUARE, loss.hpp:53-55: -2 (t - y)
      g0 = 2.f * (y0 - t0);
      g1 = 2.f * (y1 - t1);
    }
    o[j >> 1] = pack_bf16x2(g0, g1);
  }
  if (valid != 0xffffffffu) {        // warp-uniform: only the padded tail of the last tile
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (!((valid >> (2 * j)) & 1u)) o[j] &= 0xffff0000u;
      if (!((valid >> (2 * j + 1)) & 1u)) o[j] &= 0x0000ffffu;
    }
  }
  if (!row_ok) {
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = 0u;
  }
}
__device__ __forceinline__ void grad_store(const uint32_t (&o)[16], unsigned char* srow, int cbase, int sw) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(srow + (((cbase + j) ^ sw) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
}
template <int LT, bool BOUT = false>
__device__ __forceinline__ void grad_chunk(const uint32_t (&v)[32], uint32_t pos, uint32_t valid, bool row_ok,
                                           unsigned char* srow, int cbase, int sw, const float* bias = nullptr) {
  uint32_t o[16];
  grad_compute<LT, BOUT>(v, pos, valid, row_ok, o, bias);
  grad_store(o, srow, cbase, sw);
}

// The same arithmetic for a chunk of N = 16 or 32 columns (FD_EPI >= 2: 16-column chunks, two TMEM loads in
// flight).  pos / valid hold the chunk's N bits in their low bits.  The sigma path of a pair depends on
// (j & 2) / (j & 6) only, so a 16-column chunk that starts at a multiple of 16 takes, score by score, the
// path the 32-column form takes.
template <int LT, bool BOUT, int N>
__device__ __forceinline__ void grad_compute_n(const uint32_t (&v)[N], uint32_t pos, uint32_t valid, bool row_ok,
                                               uint32_t (&o)[N / 2], const float* bias = nullptr) {
  constexpr uint32_t FULL = N == 32 ? 0xffffffffu : ((1u << (N & 31)) - 1u);
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    float y0 = __uint_as_float(v[j]), y1 = __uint_as_float(v[j + 1]);
    if (BOUT) {
      const float2 b = __ldg(reinterpret_cast<const float2*>(bias + j));
      y0 += b.x;
      y1 += b.y;
    }
    const float t0 = __uint_as_float(((pos >> j) & 1u) * 0x3f800000u);
    const float t1 = __uint_as_float(((pos >> (j + 1)) & 1u) * 0x3f800000u);
    float g0, g1;
    if (LT == LOSS_CE) {
      if (FD_SIGMOID == 1 || (FD_SIGMOID == 2 && (j & 2)) || (FD_SIGMOID == 3 && (j & 6)) || (FD_SIGMOID == 4 && !(j & 6))) {
        g0 = sigmoid_sfu(y0);
        g1 = sigmoid_sfu(y1);
      } else {
        sigmoid2(y0, y1, g0, g1);
      }
      g0 -= t0;
      g1 -= t1;
    } else {
      g0 = 2.f * (y0 - t0);
      g1 = 2.f * (y1 - t1);
    }
    o[j >> 1] = pack_bf16x2(g0, g1);
  }
  if ((valid & FULL) != FULL) {      // warp-uniform: only the padded tail of the last tile
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
      if (!((valid >> (2 * j)) & 1u)) o[j] &= 0xffff0000u;
      if (!((valid >> (2 * j + 1)) & 1u)) o[j] &= 0x0000ffffu;
    }
  }
  if (!row_ok) {
#pragma unroll
    for (int j = 0; j < N / 2; ++j) o[j] = 0u;
  }
}
// NW packed words = NW / 4 16-byte pieces of this thread's staging row, starting at piece p0
template <int NW>
__device__ __forceinline__ void grad_store_n(const uint32_t (&o)[NW], unsigned char* srow, int p0, int sw) {
#pragma unroll
  for (int j = 0; j < NW / 4; ++j)
    *reinterpret_cast<uint4*>(srow + (((p0 + j) ^ sw) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
}
// 16 consecutive fp32 columns of this thread's TMEM lane (tcgen05.ld 32x32b.x16) and the matching wait
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
// 2-CTA cluster helpers: a TMA load whose bytes (and mbarrier completion) land at the same
// shared-memory offsets of every CTA in ctaMask, and a tcgen05.commit that arrives on the same
// barrier offset of every CTA in ctaMask.
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// S = Zb Wb^T for 128 users x all items of this CTA's range, tile by tile (128 x 256), fused with
// the loss gradient: the scores never leave TMEM, G goes to global memory in bf16.
// warp 0 TMA · warp 1 MMA issue · warp 2 owns TMEM · warps 4-19 epilogue: four per TMEM lane
// quadrant, warp (q, h) turns columns 64h..64h+63 of rows 32q..32q+31 into a [32][128-byte]
// staging tile and hands it to the TMA store engine (whole 128-byte lines reach L2; the first
// version stored 16 bytes per lane straight from registers and ran the L2 at 62 % with the
// tensor pipe at 21 %, profiles/r01_k_*).  The targets come from a bitmap of the slice built
// beforehand (fd_bitmap_kernel): walking the CSR rows inside this kernel, as the recommend kernel
// does, serialises one dependent global load per positive and made two helper warps the pace of
// the whole kernel at config C's 145 items per user (profiles/r01_j_*).
// CL2: launched as clusters of two CTAs (two user tiles, same item range).  Every W' k-block is
// needed by both, so each CTA fetches HALF of it (128 of the 256 item rows) and multicasts it into
// both shared memories: the L2 -> SM stream of W' (148 x |Wb| per launch, the kernel's largest L2
// consumer) is cut in two.  A stage is released to the producers when BOTH tensor cores are done with it.
// Measured neutral at config C (0.805 vs 0.799 ms per two launches: this kernel is not bound by
// that stream), so it is opt-in (CDAE_B200_FD_CLUSTER=1); kept as the building block the fused
// kernel needs, where the W' slots ARE the limit.
template <int KB, int LT, bool BOUT, bool CL2>
__global__ void __launch_bounds__(640, 1) fd_score_kernel(const __grid_constant__ CUtensorMap map_a,
                                                          const __grid_constant__ CUtensorMap map_b,
                                                          const __grid_constant__ CUtensorMap map_g, ScoreArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                                   // KB x [128][64] bf16, swizzled
  unsigned char* sB = sA + KB * A_BLK_BYTES;                  // SC_STAGE x [256][64] bf16
  unsigned char* sG = sB + SC_STAGE * B_BLK_BYTES;            // 16 x [32][64] bf16 staging tiles, swizzled
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + 16 * SC_STG_BYTES);
  uint64_t* full = bars;                  // [SC_STAGE]
  uint64_t* empty = bars + SC_STAGE;      // [SC_STAGE]
  uint64_t* a_full = bars + 2 * SC_STAGE;
  uint64_t* t_full = a_full + 1;          // [2]
  uint64_t* t_empty = t_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * TILE_U;
  const int t_lo = blockIdx.y * a.tiles_per_split;
  const int n_t = max(0, min(a.n_tiles, t_lo + a.tiles_per_split) - t_lo);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_g);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < SC_STAGE; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, CL2 ? 2 : 1);
    }
    mbar_init(a_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(t_full + b, 1);
      mbar_init(t_empty + b, 16);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && a.outputs)
    atomicAdd(a.outputs, (unsigned long long)a.n_users * (unsigned long long)a.I);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = CL2 ? cluster_ctarank() : 0u;
  if (CL2) cluster_sync_all();            // the peer's barriers are initialised before anything is multicast

  if (warp == 0) {
    if (lane == 0 && n_t > 0) {
      mbar_arrive_expect_tx(a_full, KB * A_BLK_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sA + kb * A_BLK_BYTES, &map_a, a_full, kb * KBLK, u0);
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_t; ++t) {
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(empty + s, ph ^ 1);
          mbar_arrive_expect_tx(full + s, B_BLK_BYTES);
          if (CL2)   // map_b has 128-row boxes: my half of the k-block, into both CTAs
            tma_load_2d_mc(sB + s * B_BLK_BYTES + crank * (B_BLK_BYTES / 2), &map_b, full + s, kb * KBLK,
                           (t_lo + t) * TILE_I + (int)crank * (TILE_I / 2), (uint16_t)0x3);
          else
            tma_load_2d(sB + s * B_BLK_BYTES, &map_b, full + s, kb * KBLK, (t_lo + t) * TILE_I);
          if (++s == SC_STAGE) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_t > 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_U, TILE_I);
      mbar_wait(a_full, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_t; ++t) {
        const int buf = t & 1;
        mbar_wait(t_empty + buf, ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)buf * TILE_I;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + kb * A_BLK_BYTES), b0 = smem_u32(sB + s * B_BLK_BYTES);
          const int nk = min(KBLK / 16, a.ksteps - kb * (KBLK / 16));   // trailing k-steps are all zero
          for (int k = 0; k < nk; ++k)
            umma_bf16(d, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, (kb | k) != 0);
          if (CL2) umma_commit_mc(empty + s, (uint16_t)0x3);
          else umma_commit(empty + s);
          if (++s == SC_STAGE) { s = 0; ph ^= 1; }
        }
        umma_commit(t_full + buf);
      }
    }
  } else if (warp >= 4) {
    const int e = warp - 4;
    const int q = e & 3;                    // TMEM lane quadrant (= warp % 4)
    const int h = e >> 2;                   // columns 64h .. 64h+63 of every tile
    const int row = q * 32 + lane;
    const bool row_ok = u0 + row < a.n_users;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    unsigned char* stg = sG + e * SC_STG_BYTES;
    unsigned char* srow = stg + lane * 128;
    const int sw = lane & 7;
    // this row's 64 target bits of a tile are 8 bytes of the slice's bitmap; the words of tile t + 1
    // are fetched at the END of tile t's work (a load issued before the current words are consumed
    // would share their scoreboard and stall the consumer)
    const uint2* brow = reinterpret_cast<const uint2*>(a.bits + (int64_t)(u0 + row) * (a.I_pad / 32)) + (int64_t)t_lo * 4 + h;
#if FD_EPI == 0
    uint2 nb = make_uint2(0u, 0u);
    if (n_t > 0) nb = __ldg(brow);
    for (int t = 0; t < n_t; ++t) {
      const int buf = t & 1;
      const uint32_t par = (t >> 1) & 1;
      const uint2 cb = nb;
      const int64_t item0 = (int64_t)(t_lo + t) * TILE_I + h * 64;
      const uint32_t col0 = lane_addr + (uint32_t)(buf * TILE_I + h * 64);
      mbar_wait(t_full + buf, par);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32_issue(col0, v);
      // the TMA store of the previous tile must have read the staging tile before it is rewritten
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      tmem_ld_wait(v);
      {
        const uint32_t valid = item0 + 32 <= a.I ? 0xffffffffu : (item0 >= a.I ? 0u : ((1u << (int)(a.I - item0)) - 1u));
        grad_chunk<LT, BOUT>(v, cb.x, valid, row_ok, srow, 0, sw, a.bias + item0);
      }
      tmem_ld32_issue(col0 + 32u, v);
      tmem_ld_wait(v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + buf);   // this warp's TMEM columns are in registers
      {
        const int64_t first = item0 + 32;
        const uint32_t valid = first + 32 <= a.I ? 0xffffffffu : (first >= a.I ? 0u : ((1u << (int)(a.I - first)) - 1u));
        grad_chunk<LT, BOUT>(v, cb.y, valid, row_ok, srow, 4, sw, a.bias + first);
      }
      fence_proxy_async();                         // staging writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&map_g, stg, (int)item0, u0 + q * 32);
        bulk_commit();
      }
      if (t + 1 < n_t) nb = __ldg(brow + 4 * (t + 1));
    }
#elif FD_EPI == 1
    // As FD_EPI = 0, but nothing of the tile's work waits behind the TMA engine: the wait for "the previous
    // tile's store has read the staging tile" sits between the first chunk's arithmetic and its stores, and
    // the second TMEM load is issued before it, so both latencies run under each other.
    uint2 nb = make_uint2(0u, 0u);
    if (n_t > 0) nb = __ldg(brow);
    for (int t = 0; t < n_t; ++t) {
      const int buf = t & 1;
      const uint32_t par = (t >> 1) & 1;
      const uint2 cb = nb;
      const int64_t item0 = (int64_t)(t_lo + t) * TILE_I + h * 64;
      const uint32_t col0 = lane_addr + (uint32_t)(buf * TILE_I + h * 64);
      mbar_wait(t_full + buf, par);
      tc_fence_after();
      uint32_t v[32], o[16];
      tmem_ld32_issue(col0, v);
      tmem_ld_wait(v);
      {
        const uint32_t valid = item0 + 32 <= a.I ? 0xffffffffu : (item0 >= a.I ? 0u : ((1u << (int)(a.I - item0)) - 1u));
        grad_compute<LT, BOUT>(v, cb.x, valid, row_ok, o, a.bias + item0);
      }
      tmem_ld32_issue(col0 + 32u, v);
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      grad_store(o, srow, 0, sw);
      tmem_ld_wait(v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + buf);   // this warp's TMEM columns are in registers
      {
        const int64_t first = item0 + 32;
        const uint32_t valid = first + 32 <= a.I ? 0xffffffffu : (first >= a.I ? 0u : ((1u << (int)(a.I - first)) - 1u));
        grad_compute<LT, BOUT>(v, cb.y, valid, row_ok, o, a.bias + first);
        grad_store(o, srow, 4, sw);
      }
      fence_proxy_async();                         // staging writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&map_g, stg, (int)item0, u0 + q * 32);
        bulk_commit();
      }
      if (t + 1 < n_t) nb = __ldg(brow + 4 * (t + 1));
    }
#else
    // FD_EPI = 2: the tile's 64 columns are read in four 16-column TMEM loads, always one in flight under
    // the arithmetic of the previous chunk (registers: two 16-word buffers instead of one 32-word buffer), and
    // the wait for the staging tile sits between the first chunk's arithmetic and its stores.
    // FD_EPI = 3: in addition the first TMEM load of tile t + 1 is issued before the last chunk of tile t is
    // worked on, and the target words are fetched two tiles ahead.
    constexpr bool AHEAD = FD_EPI >= 3;
    uint2 nb = make_uint2(0u, 0u), nb2 = make_uint2(0u, 0u);
    if (n_t > 0) nb = __ldg(brow);
    if (AHEAD && n_t > 1) nb2 = __ldg(brow + 4);
    uint32_t va[16], vb[16];
    if (AHEAD && n_t > 0) {
      mbar_wait(t_full + 0, 0u);
      tc_fence_after();
      tmem_ld16_issue(lane_addr + (uint32_t)(h * 64), va);
    }
    for (int t = 0; t < n_t; ++t) {
      const int buf = t & 1;
      const uint32_t par = (t >> 1) & 1;
      const uint2 cb = nb;
      const int64_t item0 = (int64_t)(t_lo + t) * TILE_I + h * 64;
      const uint32_t col0 = lane_addr + (uint32_t)(buf * TILE_I + h * 64);
      if (!AHEAD) {
        mbar_wait(t_full + buf, par);
        tc_fence_after();
        tmem_ld16_issue(col0, va);
      }
      tmem_ld_wait16(va);
      tmem_ld16_issue(col0 + 16u, vb);
      {
        const uint32_t valid = item0 + 16 <= a.I ? 0xffffu : (item0 >= a.I ? 0u : ((1u << (int)(a.I - item0)) - 1u));
        uint32_t o[8];
        grad_compute_n<LT, BOUT, 16>(va, cb.x & 0xffffu, valid, row_ok, o, a.bias + item0);
        // the TMA store of the previous tile must have read the staging tile before it is rewritten
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        grad_store_n<8>(o, srow, 0, sw);
      }
      tmem_ld_wait16(vb);
      tmem_ld16_issue(col0 + 32u, va);
      {
        const int64_t first = item0 + 16;
        const uint32_t valid = first + 16 <= a.I ? 0xffffu : (first >= a.I ? 0u : ((1u << (int)(a.I - first)) - 1u));
        uint32_t o[8];
        grad_compute_n<LT, BOUT, 16>(vb, cb.x >> 16, valid, row_ok, o, a.bias + first);
        grad_store_n<8>(o, srow, 2, sw);
      }
      tmem_ld_wait16(va);
      tmem_ld16_issue(col0 + 48u, vb);
      {
        const int64_t first = item0 + 32;
        const uint32_t valid = first + 16 <= a.I ? 0xffffu : (first >= a.I ? 0u : ((1u << (int)(a.I - first)) - 1u));
        uint32_t o[8];
        grad_compute_n<LT, BOUT, 16>(va, cb.y & 0xffffu, valid, row_ok, o, a.bias + first);
        grad_store_n<8>(o, srow, 4, sw);
      }
      tmem_ld_wait16(vb);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + buf);   // this warp's TMEM columns are in registers
      if (AHEAD && t + 1 < n_t) {                  // va is free: the first chunk of the next tile
        mbar_wait(t_full + (buf ^ 1), (uint32_t)(((t + 1) >> 1) & 1));
        tc_fence_after();
        tmem_ld16_issue(lane_addr + (uint32_t)((buf ^ 1) * TILE_I + h * 64), va);
      }
      {
        const int64_t first = item0 + 48;
        const uint32_t valid = first + 16 <= a.I ? 0xffffu : (first >= a.I ? 0u : ((1u << (int)(a.I - first)) - 1u));
        uint32_t o[8];
        grad_compute_n<LT, BOUT, 16>(vb, cb.y >> 16, valid, row_ok, o, a.bias + first);
        grad_store_n<8>(o, srow, 6, sw);
      }
      fence_proxy_async();                         // staging writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&map_g, stg, (int)item0, u0 + q * 32);
        bulk_commit();
      }
      if (AHEAD) {
        nb = nb2;
        if (t + 2 < n_t) nb2 = __ldg(brow + 4 * (t + 2));
      } else if (t + 1 < n_t) {
        nb = __ldg(brow + 4 * (t + 1));
      }
    }
    if (AHEAD && n_t > 0) tmem_ld_wait16(va);      // (nothing outstanding; keeps va's last definition consumed)
#endif
    if (lane == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (CL2) cluster_sync_all();            // no CTA leaves while its peer may still multicast into it
}

// ---------------------------------------------------------------------------------------
// Fused score + hidden-gradient kernel.  fd_score_kernel is paced by its epilogue (one MUFU and ~12
// issue slots per score against 0.05 clk of tensor time), so the tensor pipe idles 80 % of the time,
// while the separate HG = G W' launch re-reads all of G from HBM.  Here the SAME CTA that turns a
// 128 x 64 score tile into loss gradients hands that tile — already in shared memory as a K-major
// UMMA operand, because that is also the layout the TMA store wants — straight back to the tensor
// core: HG[128 x Kp] += G_tile[128 x 64] . W'_tile[64 x Kp], with W'_tile the very k-blocks the
// first contraction just used, re-read MN-major.  The second contraction hides completely under
// the epilogue; G still goes out once (for the item gradient), but is read once instead of twice.
//
//   per 64-item tile t (slot = t % 3, s = t & 1):
//     TMA      W' tile -> slot                         (KB boxes {64 k, 64 items}, 8 KB each)
//     MMA      S[s]  = Zb . W'_tile^T                  (N = 64; 128 x 64 fp32 in TMEM columns 64 s ..)
//     epilogue set s (8 warps): S[s] -> g -> bf16 -> Gs[s] (shared, swizzled) -> TMA store to G
//     MMA      HG   += Gs[s] . W'_tile                 (N = Kp; TMEM columns 128 .. 128 + Kp)
//   the two epilogue sets work on alternate tiles, so both S buffers / G buffers are in flight.
// warp 0 TMA loads · warp 1 MMA issue · warp 2 owns TMEM · warp 3 TMA stores of G · warps 4-19
// epilogue: warp = (set s, quadrant q, half h) handles rows 32q.., columns 32h.. of the tiles with
// t & 1 == s.  All hand-offs are mbarriers (no CTA-wide or set-wide bar.sync in the tile loop).
constexpr int FU_TILE_I = 64;
constexpr int FU_SLOTS = 4;                           // W' tiles in flight: t (2nd contraction) .. t + 3 (landing)
constexpr int FU_G_BYTES = 128 * 128;                 // one G tile: [128 users][64 items] bf16
__host__ __device__ constexpr size_t fused_smem(int kb) {
  return 1024 + (size_t)kb * A_BLK_BYTES + (size_t)FU_SLOTS * kb * GM_BOX_BYTES + 2 * FU_G_BYTES + 256;
}

struct FusedArgs {
  int n_users;
  int64_t I, I_pad;
  int n_tiles;               // I_pad / 64
  int tiles_per_split;
  int ksteps;                // ceil((K + 2) / 16)
  int K, ld;
  const uint32_t* bits;      // [B_pad][I_pad / 32]
  float* HG;                 // [n_users][ld], zero or partial sums (reduced into)
  unsigned long long* outputs;
};

// CL2: clusters of two CTAs (two user tiles, same item range) fetch half of every W' tile each and
// multicast it into both shared memories; a slot is refilled when both tensor cores released it.
template <int KB, int LT, bool CL2>
__global__ void __launch_bounds__(640, 1) fd_fused_kernel(const __grid_constant__ CUtensorMap map_a,
                                                          const __grid_constant__ CUtensorMap map_w,
                                                          const __grid_constant__ CUtensorMap map_g, FusedArgs a) {
  constexpr int NCOL = KB * KBLK;                 // Kp
  constexpr int W_TILE_BYTES = KB * GM_BOX_BYTES;
  constexpr uint32_t HG_COL = 128;                // TMEM column of the HG accumulator
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                                   // KB x [128][64] bf16
  unsigned char* sW = sA + KB * A_BLK_BYTES;                  // FU_SLOTS x KB x [64][64] bf16
  unsigned char* sG = sW + FU_SLOTS * W_TILE_BYTES;           // 2 x [128][64] bf16
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + 2 * FU_G_BYTES);
  uint64_t* w_full = bars;                 // [FU_SLOTS]
  uint64_t* w_empty = bars + FU_SLOTS;     // [FU_SLOTS]
  uint64_t* a_full = bars + 2 * FU_SLOTS;
  uint64_t* t_full = a_full + 1;           // [2] scores ready
  uint64_t* t_empty = t_full + 2;          // [2] scores in registers
  uint64_t* g_full = t_empty + 2;          // [2] gradient tile written
  uint64_t* g_empty = g_full + 2;          // [2] gradient tile consumed by the second contraction
  uint64_t* hg_full = g_empty + 2;
  uint64_t* s_free = hg_full + 1;          // [2] gradient tile read by the TMA store
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * TILE_U;
  const int t_lo = blockIdx.y * a.tiles_per_split;
  const int n_t = max(0, min(a.n_tiles, t_lo + a.tiles_per_split) - t_lo);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_g);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < FU_SLOTS; ++i) {
      mbar_init(w_full + i, 1);
      mbar_init(w_empty + i, CL2 ? 2 : 1);
    }
    mbar_init(a_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(t_full + b, 1);
      mbar_init(t_empty + b, 8);
      mbar_init(g_full + b, 8);       // one arrive per epilogue warp of the set
      mbar_init(g_empty + b, 1);
      mbar_init(s_free + b, 1);
    }
    mbar_init(hg_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && a.outputs)
    atomicAdd(a.outputs, (unsigned long long)a.n_users * (unsigned long long)a.I);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = CL2 ? cluster_ctarank() : 0u;
  if (CL2) cluster_sync_all();

  if (warp == 0) {
    if (lane == 0 && n_t > 0) {
      mbar_arrive_expect_tx(a_full, KB * A_BLK_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sA + kb * A_BLK_BYTES, &map_a, a_full, kb * KBLK, u0);
      for (int t = 0; t < n_t; ++t) {
        const int slot = t % FU_SLOTS;
        mbar_wait(w_empty + slot, ((t / FU_SLOTS) & 1) ^ 1);
        mbar_arrive_expect_tx(w_full + slot, W_TILE_BYTES);
        for (int kb = 0; kb < KB; ++kb) {
          unsigned char* dst = sW + slot * W_TILE_BYTES + kb * GM_BOX_BYTES;
          if (CL2)   // map_w has 32-row boxes: my half of the rows, into both CTAs
            tma_load_2d_mc(dst + crank * (GM_BOX_BYTES / 2), &map_w, w_full + slot, kb * KBLK,
                           (t_lo + t) * FU_TILE_I + (int)crank * (FU_TILE_I / 2), (uint16_t)0x3);
          else
            tma_load_2d(dst, &map_w, w_full + slot, kb * KBLK, (t_lo + t) * FU_TILE_I);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_t > 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(TILE_U, FU_TILE_I);                 // S: both K-major
      constexpr uint32_t idesc2 = umma_idesc_bf16_major(TILE_U, NCOL, 0, 1);          // HG: B MN-major
      mbar_wait(a_full, 0);
      // first contraction of tile t
      auto scores = [&](int t) {
        const int sb = t & 1, slot = t % FU_SLOTS;
        mbar_wait(t_empty + sb, ((t >> 1) & 1) ^ 1);
        mbar_wait(w_full + slot, (t / FU_SLOTS) & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(sb * FU_TILE_I);
        const uint32_t w0 = smem_u32(sW + slot * W_TILE_BYTES);
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t a0 = smem_u32(sA + kb * A_BLK_BYTES), b0 = w0 + kb * GM_BOX_BYTES;
          const int nk = min(KBLK / 16, a.ksteps - kb * (KBLK / 16));
          for (int k = 0; k < nk; ++k)
            umma_bf16(d, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc1, (kb | k) != 0);
        }
        umma_commit(t_full + sb);
      };
      scores(0);
      if (n_t > 1) scores(1);
      for (int t = 0; t < n_t; ++t) {
        const int sb = t & 1, slot = t % FU_SLOTS;
        // scores of tile t + 2 go in FIRST: they only need S[sb] to be in the epilogue's registers
        // (early in its work on tile t), and must be ready when that set comes back for more; the
        // second contraction of tile t is needed by nobody for two more tiles
        if (t + 2 < n_t) scores(t + 2);
        mbar_wait(g_full + sb, (t >> 1) & 1);
        tc_fence_after();
        const uint32_t g0 = smem_u32(sG + sb * FU_G_BYTES), w0 = smem_u32(sW + slot * W_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < FU_TILE_I / 16; ++k)
          umma_bf16(tmem_base + HG_COL, umma_desc_sw128(g0 + k * 32), umma_desc_mn_sw128(w0 + k * 2048, GM_BOX_BYTES),
                    idesc2, (t | k) != 0);
        umma_commit(g_empty + sb);
        if (CL2) umma_commit_mc(w_empty + slot, (uint16_t)0x3);
        else umma_commit(w_empty + slot);
      }
      umma_commit(hg_full);
    }
  } else if (warp == 3) {
    // ===== G store: as soon as a gradient tile is complete it goes out through the TMA engine
    if (lane == 0) {
      for (int t = 0; t < n_t; ++t) {
        const int sb = t & 1;
        mbar_wait(g_full + sb, (t >> 1) & 1);
        tma_store_2d(&map_g, sG + sb * FU_G_BYTES, (t_lo + t) * FU_TILE_I, u0);
        bulk_commit();
        bulk_wait_read0();                  // the engine has read the tile out of shared memory
        mbar_arrive(s_free + sb);
      }
      bulk_wait0();
    }
  } else if (warp >= 4) {
    const int e = warp - 4;
    const int q = e & 3;                    // TMEM lane quadrant (= warp % 4)
    const int h = (e >> 2) & 1;             // columns 32h .. 32h+31 of the tile
    const int sb = e >> 3;                  // tiles with t & 1 == sb
    const int row = q * 32 + lane;
    const bool row_ok = u0 + row < a.n_users;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    unsigned char* gt = sG + sb * FU_G_BYTES;
    unsigned char* srow = gt + row * 128;
    const int sw = row & 7;
    const uint32_t* brow = a.bits + (int64_t)(u0 + row) * (a.I_pad / 32) + (int64_t)t_lo * 2 + h;
    uint32_t nb = 0u;
    if (sb < n_t) nb = __ldg(brow + 2 * sb);
    for (int t = sb; t < n_t; t += 2) {
      const uint32_t par = (t >> 1) & 1;
      const uint32_t cb = nb;
      const int64_t item0 = (int64_t)(t_lo + t) * FU_TILE_I;
      mbar_wait(t_full + sb, par);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32_issue(lane_addr + (uint32_t)(sb * FU_TILE_I + h * 32), v);
      tmem_ld_wait(v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + sb);
      // the gradient tile of tile t - 2 must have been read by the second contraction and by the
      // TMA store before it is overwritten (checked AFTER the arithmetic, when both are long done);
      // no warp waits for its siblings — each one publishes its 32 x 32 block and moves on
      uint32_t o[16];
      {
        const int64_t first = item0 + h * 32;
        const uint32_t valid = first + 32 <= a.I ? 0xffffffffu : (first >= a.I ? 0u : ((1u << (int)(a.I - first)) - 1u));
        grad_compute<LT>(v, cb, valid, row_ok, o);
      }
      mbar_wait(g_empty + sb, par ^ 1);
      mbar_wait(s_free + sb, par ^ 1);
      grad_store(o, srow, h * 4, sw);
      fence_proxy_async();                  // generic-proxy writes -> visible to UMMA and TMA
      __syncwarp();
      if (lane == 0) mbar_arrive(g_full + sb);
      if (t + 2 < n_t) nb = __ldg(brow + 2 * (t + 2));
    }
    // ---- hidden gradient of this CTA's item range: thread = user row, warps of a quadrant share the columns
    if (n_t > 0) {
      mbar_wait(hg_full, 0);
      tc_fence_after();
      const int part = e >> 2;              // 0..3
      for (int c = part; c * 32 < a.K; c += 4) {
        uint32_t v[32];
        tmem_ld32_issue(lane_addr + HG_COL + (uint32_t)(c * 32), v);
        tmem_ld_wait(v);
        if (!row_ok) continue;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int col = c * 32 + j;
          if (col >= a.K) break;
          float x[4] = {__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])};
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (col + i >= a.K) x[i] = 0.f;                    // bias columns: not part of the row
          red_add_v4(a.HG + (int64_t)(u0 + row) * a.ld + col, make_float4(x[0], x[1], x[2], x[3]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (CL2) cluster_sync_all();
}

// ---------------------------------------------------------------------------------------
struct GemmArgs {
  int n_steps;          // contraction steps of 64 (items for the hidden gradient, users for the item gradient)
  int steps_per_split;  // gridDim.y CTAs share the contraction; partial sums meet in 16-byte reductions
  int n_rows;           // valid output rows: users of the slice / items
  int K, ld;
  float* out;           // HG [n_users][ld]  |  gW' [I][ld]      (zero, or holding other contributions)
  float* out_bias;      // item gradient only: gb' [I]
  const float* W;       // item gradient only: W' and b' for the lambda terms
  const float* bp;
  float nlambda;        // (#users of the slice) * lambda: every user contributes lambda*theta per item
};

// ITEMGRAD = false:  HG[u][k]  = sum_i G[u][i] Wb[i][k]     A = G   (K-major),  B = Wb (MN-major)
// ITEMGRAD = true :  gW'[i][k] = sum_u G[u][i] Zb[u][k]     A = G^T (MN-major), B = Zb (MN-major)
//                    (+ n*lambda*W'[i][k]; column K -> gb'[i] + n*lambda*b'[i])
// One CTA = one 128-row output tile x all Kp columns (accumulator: Kp TMEM columns), walking its
// share of the contraction in steps of 64 through a GM_STAGE-deep TMA ring.
// warp 0 TMA · warp 1 MMA issue · warp 2 TMEM owner · warps 4-7 epilogue (thread = output row).
// CL2: clusters of two CTAs (adjacent output tiles, same share of the contraction) read the SAME
// B slabs (Wb / Zb rows of the contraction step); each fetches every other 64-column box and
// multicasts it into both shared memories, halving the B stream out of L2.
template <int KB, bool ITEMGRAD, bool BOUT = false, bool CL2 = false>
__global__ void __launch_bounds__(256, 1) fd_gemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                         const __grid_constant__ CUtensorMap map_b, GemmArgs a) {
  constexpr int STAGE_BYTES = GM_A_BYTES + KB * GM_BOX_BYTES;
  constexpr int NCOL = KB * KBLK;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // bias-outside mode: gb' = G^T 1 comes from a second, 16-column accumulator fed by a constant
  // all-ones MN-major slab (64 contraction rows x 128 bytes)
  unsigned char* sOnes = smem + GM_STAGE * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + (BOUT ? GM_BOX_BYTES : 0));
  uint64_t* full = bars;
  uint64_t* empty = bars + GM_STAGE;
  uint64_t* t_full = bars + 2 * GM_STAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_full + 1);
  constexpr int TCOLS = BOUT ? 512 : tmem_cols(KB);
  constexpr uint32_t ONES_COL = 256;
  if (BOUT) {
    for (int i = threadIdx.x; i < GM_BOX_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3f803f80u;
    fence_proxy_async();
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int s_lo = blockIdx.y * a.steps_per_split;
  const int n_s = max(0, min(a.n_steps, s_lo + a.steps_per_split) - s_lo);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GM_STAGE; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, CL2 ? 2 : 1);
    }
    mbar_init(t_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = CL2 ? cluster_ctarank() : 0u;
  if (CL2) cluster_sync_all();

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_s; ++t) {
        mbar_wait(empty + s, ph ^ 1);
        unsigned char* st = smem + s * STAGE_BYTES;
        mbar_arrive_expect_tx(full + s, STAGE_BYTES);
        const int k0 = (s_lo + t) * KBLK;   // first contraction index of the step
        if (ITEMGRAD) {
          // G^T: rows = users k0.., 128 bytes = 64 items; two boxes cover the tile's 128 items
          tma_load_2d(st, &map_a, full + s, m0, k0);
          tma_load_2d(st + GM_BOX_BYTES, &map_a, full + s, m0 + 64, k0);
        } else {
          // G: rows = the tile's 128 users, 128 bytes = items k0..k0+63
          tma_load_2d(st, &map_a, full + s, k0, m0);
        }
        for (int kb = 0; kb < KB; ++kb) {   // rows = contraction index, 128 bytes = columns kb*64..
          if (!CL2) tma_load_2d(st + GM_A_BYTES + kb * GM_BOX_BYTES, &map_b, full + s, kb * KBLK, k0);
          else if ((uint32_t)(kb & 1) == crank)
            tma_load_2d_mc(st + GM_A_BYTES + kb * GM_BOX_BYTES, &map_b, full + s, kb * KBLK, k0, (uint16_t)0x3);
        }
        if (++s == GM_STAGE) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_s > 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_major(128, NCOL, ITEMGRAD ? 1 : 0, 1);
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_s; ++t) {
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), b0 = a0 + GM_A_BYTES;
#pragma unroll
        for (int k = 0; k < KBLK / 16; ++k) {
          // 16 contraction indices: 32 bytes along a K-major row, 16 rows (2048 bytes) of an MN-major slab
          const uint64_t ad = ITEMGRAD ? umma_desc_mn_sw128(a0 + k * 2048, GM_BOX_BYTES) : umma_desc_sw128(a0 + k * 32);
          const uint64_t bd = umma_desc_mn_sw128(b0 + k * 2048, GM_BOX_BYTES);
          umma_bf16(tmem_base, ad, bd, idesc, (t | k) != 0);
          if (BOUT) {
            constexpr uint32_t idesc1 = umma_idesc_bf16_major(128, 16, 1, 1);
            umma_bf16(tmem_base + ONES_COL, ad, umma_desc_mn_sw128(smem_u32(sOnes) + k * 2048, GM_BOX_BYTES), idesc1, (t | k) != 0);
          }
        }
        if (CL2) umma_commit_mc(empty + s, (uint16_t)0x3);
        else umma_commit(empty + s);
        if (++s == GM_STAGE) { s = 0; ph ^= 1; }
      }
      umma_commit(t_full);
    }
  } else if (warp >= 4 && n_s > 0) {
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < a.n_rows;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    mbar_wait(t_full, 0);
    tc_fence_after();
    const int last_col = (ITEMGRAD && !BOUT) ? a.K : a.K - 1;     // column K of the item gradient is gb'
    if (BOUT) {
      uint32_t v1[32];
      tmem_ld32_issue(lane_addr + ONES_COL, v1);   // 16 identical columns (+ 16 unused ones)
      tmem_ld_wait(v1);
      if (row_ok) {
        float gb = __uint_as_float(v1[0]);
        if (blockIdx.y == 0 && a.nlambda != 0.f) gb = fmaf(a.nlambda, a.bp[row], gb);
        red_add_f32(a.out_bias + row, gb);
      }
    }
    const bool add_l2 = ITEMGRAD && blockIdx.y == 0 && a.nlambda != 0.f;
    for (int c = 0; c * 32 <= last_col; ++c) {
      uint32_t v[32];
      tmem_ld32_issue(lane_addr + (uint32_t)(c * 32), v);
      tmem_ld_wait(v);
      if (!row_ok) continue;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int col = c * 32 + j;
        if (col > last_col) break;
        float x[4] = {__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])};
        if (ITEMGRAD && !BOUT && col <= a.K && col + 3 >= a.K) {
          const int kk = a.K - col;
          float gb = kk == 0 ? x[0] : kk == 1 ? x[1] : kk == 2 ? x[2] : x[3];
          if (add_l2) gb = fmaf(a.nlambda, a.bp[row], gb);
          red_add_f32(a.out_bias + row, gb);
        }
        if (col < a.K) {
          if (add_l2) {
            const float4 w = ld4(a.W + (int64_t)row * a.ld + col);   // pad columns of W' are 0
            x[0] = fmaf(a.nlambda, w.x, x[0]); x[1] = fmaf(a.nlambda, w.y, x[1]);
            x[2] = fmaf(a.nlambda, w.z, x[2]); x[3] = fmaf(a.nlambda, w.w, x[3]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (col + i >= a.K) x[i] = 0.f;                        // bias columns: not part of the row
          red_add_v4(a.out + (int64_t)row * a.ld + col, make_float4(x[0], x[1], x[2], x[3]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
  if (CL2) cluster_sync_all();
}

}  // namespace fd
}  // namespace cdae
static_assert(cdae::fd::score_smem(4) <= 232448 && cdae::fd::gemm_smem(4) <= 232448 && cdae::fd::fused_smem(4) <= 232448,
              "full-decode kernels exceed the 227 KB of shared memory a CTA can opt into");
