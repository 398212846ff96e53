// cdae_b200/csrc/topn_tc.cuh — full-item decode of CDAE::recommend (cdae.hpp:176-186) on the
// 5th-generation tensor cores: S = Z · W'ᵀ (+ b') for a tile of 128 users against ALL items, with
// the top-candidate selection fused into the epilogue so the U x I score matrix never exists.
//
//   operands  bf16, K-major, 128-byte-swizzled in shared memory (TMA, cp.async.bulk.tensor)
//             A = Zb  [users_pad][Kp]   Kp = round_up(K + 2, 64); columns K, K+1 hold 1.0
//             B = Wb  [items_pad][Kp]   columns K, K+1 hold bf16 hi / lo of b'  (bias rides in
//                                       the contraction, no epilogue add)
//   MMA       tcgen05.mma.cta_group::1.kind::f16, M = 128 (users) x N = 256 (items) x K = 16,
//             fp32 accumulators in TMEM, two 256-column buffers (all 512 columns)
//   roles     warp 0 TMA producer · warp 1 MMA issuer · warps 2-3 rated-item bitmaps (warp 2 also
//             owns the TMEM allocation) · warps 4-7 epilogue (tcgen05.ld, one TMEM lane = one user
//             per thread)
//
// The scores are APPROXIMATE (bf16 operands).  Exactness of the final lists is restored
// afterwards (topn_api.inl): the epilogue keeps, per user, every unrated item whose approximate
// score exceeds a running threshold `thr` (raised by periodic compaction, never lowered), so at
// the end every unrated item that is NOT a candidate has approx score <= thr; the re-rank kernel
// scores the candidates in fp64 and accepts the user's list only if  thr + eps_u < (k-th best
// exact score), eps_u being a rigorous bound on |approx - exact| (pack_z_bf16_kernel).  Users
// that fail the test are re-done by the exact fp32 kernel.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace cdae {
namespace tc {

constexpr int TILE_U = 128;   // users per CTA  (UMMA M)
constexpr int TILE_I = 256;   // items per accumulator tile (UMMA N)
constexpr int KBLK = 64;      // bf16 elements per 128-byte swizzle row
constexpr int NSTAGE = 3;     // B k-blocks in flight
constexpr int A_BLK_BYTES = TILE_U * KBLK * 2;   // 16 KB
constexpr int B_BLK_BYTES = TILE_I * KBLK * 2;   // 32 KB
constexpr int MAX_KB = 5;     // Kp <= 320  (K <= 318)
constexpr int CAND_MAX = 96;  // largest per-user candidate buffer

// Epilogue warps per TMEM lane quadrant: two while shared memory leaves room for two candidate
// buffers per user (each warp of a pair takes every other 32-column chunk and keeps its own
// buffer and threshold), one for the widest operands.
#ifndef TC_KB4_EPI
#define TC_KB4_EPI 1   // A/B switch (CDAE_NVCC_FLAGS=-DTC_KB4_EPI=2): paired epilogue warps for 4 k-blocks too
#endif
#ifndef TC_KB1_EPI
#define TC_KB1_EPI 2   // A/B switch (-DTC_KB1_EPI=4): at K <= 62 a tile is 4 UMMAs followed by 32 K score reads, so the
#endif                 // epilogue is the critical path — but FOUR warps per TMEM lane quadrant measured 3.95 ms against
                       // 2.27 ms with two at config B (24-slot candidate buffers compact far more often, 96 registers
                       // with spills, a 16-warp barrier per tile); profiles/r02_j_topn_epilogue_ab.json
__host__ __device__ constexpr int epi_warps(int kb) { return kb == 1 ? TC_KB1_EPI : kb <= 3 ? 2 : (kb == 4 ? TC_KB4_EPI : 1); }
// candidate slots per user (all buffers together): what is left of the 227 KB after A and B
__host__ __device__ constexpr int cand_slots(int kb) {
  return kb == 1 ? 96 : kb == 2 ? 80 : kb == 3 ? 64 : kb == 4 ? (TC_KB4_EPI == 2 ? 54 : 48) : 36;
}
__host__ __device__ constexpr int buf_slots(int kb) { return cand_slots(kb) / epi_warps(kb); }
// a compaction keeps between keep_lo and keep_hi entries of a buffer
__host__ __device__ constexpr int keep_hi(int kb) { return buf_slots(kb) / 2; }
__host__ __device__ constexpr int keep_lo(int kb) { return keep_hi(kb) - 8 > 12 ? keep_hi(kb) - 8 : (keep_hi(kb) > 12 ? 12 : 10); }
__host__ __device__ constexpr int n_threads(int kb) { return 128 + 128 * epi_warps(kb); }
constexpr int BM_BYTES = 2 * 8 * TILE_U * 4;  // two rated bitmaps [8 words][128 rows]
__host__ __device__ constexpr size_t smem_bytes(int kb) {
  return 1024 /*alignment slack*/ + (size_t)kb * A_BLK_BYTES + (size_t)NSTAGE * B_BLK_BYTES +
         (size_t)cand_slots(kb) * TILE_U * 8 + BM_BYTES + (size_t)(epi_warps(kb) < 2 ? 2 : epi_warps(kb)) * TILE_U * 8 /*merge*/ +
         (size_t)(epi_warps(kb) < 2 ? 2 : epi_warps(kb)) * 4 * 32 /*hit exchange*/ + 256 /*barriers*/;
}

// ---------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane: issue, then wait.  The wait names the
// destination registers as in/out operands so the compiler cannot move their uses above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),
                 "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]),
                 "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]),
                 "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B (cute::UMMA::SmemDescriptor):
// start address >> 4 in [0,14); LBO (unused for swizzled K-major) in [16,30); SBO = 1024 B (eight
// 128-byte rows) >> 4 in [32,46); descriptor version 1 in [46,48); layout type 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (bits 4-5 = 1), A and B bf16
// (bits 7-9 = 1, 10-12 = 1), both K-major (bits 15, 16 = 0), N >> 3 in [17,23), M >> 4 in [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------
// Operand packing.
// Column-wise max |W'| (K values) and max |b'|, for the error bound.  out[K] = bmax.
__global__ void __launch_bounds__(256) absmax_cols_kernel(const float* __restrict__ W, const float* __restrict__ bp,
                                                          int64_t I, int K, int ld, float* __restrict__ out) {
  // block handles a slab of rows; thread t handles columns t, t+256, ...
  const int64_t rows_per_block = (I + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = blockIdx.x * rows_per_block, r1 = min(I, r0 + rows_per_block);
  for (int k = threadIdx.x; k <= K; k += blockDim.x) {
    float m = 0.f;
    if (k < K) {
      for (int64_t r = r0; r < r1; ++r) m = fmaxf(m, fabsf(W[r * ld + k]));
    } else {
      for (int64_t r = r0; r < r1; ++r) m = fmaxf(m, fabsf(bp[r]));
    }
    atomicMax(reinterpret_cast<int*>(out + k), __float_as_int(m));  // non-negative floats order as ints
  }
}

// out[K+1] = max_i ||W'[i]||_2^2 (one warp per row), the Cauchy-Schwarz side of the error bound.
__global__ void __launch_bounds__(256) rownorm_max_kernel(const float* __restrict__ W, int64_t I, int K, int ld,
                                                          float* __restrict__ out) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  if (r < I)
    for (int k = lane; k < K; k += 32) {
      const float v = W[r * ld + k];
      s = fmaf(v, v, s);
    }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0 && r < I) atomicMax(reinterpret_cast<int*>(out + K + 1), __float_as_int(s));
}

// Wb[i][0..K) = bf16(W'[i]), Wb[i][K] = bf16(b'[i]), Wb[i][K+1] = bf16(b'[i] - hi); rows >= I and
// the remaining pad columns are 0.  One thread per (row, 8-column group).
__global__ void __launch_bounds__(256) pack_w_bf16_kernel(const float* __restrict__ W, const float* __restrict__ bp,
                                                          int64_t I, int64_t I_pad, int K, int ld, int Kp,
                                                          __nv_bfloat16* __restrict__ out) {
  const int g8 = Kp / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= I_pad * g8) return;
  const int64_t r = idx / g8;
  const int c0 = (int)(idx % g8) * 8;
  __align__(16) __nv_bfloat16 o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    float v = 0.f;
    if (r < I) {
      if (c < K) v = W[r * ld + c];
      else if (c == K) v = bp[r];
      else if (c == K + 1) {
        const float b = bp[r];
        v = b - __bfloat162float(__float2bfloat16_rn(b));
      }
    }
    o[j] = __float2bfloat16_rn(v);
  }
  *reinterpret_cast<uint4*>(out + r * Kp + c0) = *reinterpret_cast<const uint4*>(o);
}

// Zb[r][0..K) = bf16(z of user r), Zb[r][K] = Zb[r][K+1] = 1; rows >= n are 0.  Also the per-user
// error bound: approx = sum_k bf16(z_k) bf16(w_k) + bhi + blo accumulated in fp32 by the tensor core;
//   |approx - exact| <= (2^-8 * 1.01 + (Kp + 2) * 2^-22) * S   (operand rounding 2^-9 each, product
//   exact, fp32 accumulation)  +  (2^-16 + (Kp + 2) * 2^-22) * bmax   (bias residual), where
//   S >= sum_k |z_k||w_ik| for every item i:  S = min( sum_k |z_k| max_i|w_ik| ,  ||z||_2 max_i||w_i||_2 ).
// One warp per user.  wmax = [K column maxima | bmax | max row norm squared].
__global__ void __launch_bounds__(256) pack_z_bf16_kernel(const float* __restrict__ Z, const int32_t* __restrict__ users,
                                                          int n, int64_t n_pad, int K, int ld, int Kp,
                                                          const float* __restrict__ wmax /*[K+1]*/,
                                                          __nv_bfloat16* __restrict__ out, float* __restrict__ eps) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_pad) return;
  float acc = 0.f, zz = 0.f;
  const float* z = nullptr;
  if (r < n) z = Z + (int64_t)(users ? users[r] : r) * ld;
  for (int c = lane; c < Kp; c += 32) {
    float v = 0.f;
    if (z) {
      if (c < K) {
        v = z[c];
        acc += fabsf(v) * wmax[c];
        zz = fmaf(v, v, zz);
      } else if (c <= K + 1) {
        v = 1.f;
      }
    }
    out[r * Kp + c] = __float2bfloat16_rn(v);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, off);
    zz += __shfl_xor_sync(0xffffffffu, zz, off);
  }
  if (lane == 0 && r < n) {
    const float slack = (float)(Kp + 2) * 2.3841858e-7f;  // 2^-22
    const float cs = sqrtf(zz) * sqrtf(wmax[K + 1]) * 1.0001f;  // Cauchy-Schwarz, rounded up
    eps[r] = (0.00390625f * 1.01f + slack) * fminf(acc, cs) + (1.52587890625e-5f + slack) * wmax[K];
  }
}

// Rated bitmap of a user list for the prepass mode of topn_tc_kernel: bit i of row r <=> item i is in
// the train row of user users[r] (or r) OR i >= I (padded tail columns are never candidates).
// bits [n_pad][words] must be zero; one warp per user.
__global__ void __launch_bounds__(256) topn_bitmap_kernel(const int32_t* __restrict__ users, int n,
                                                          const int64_t* __restrict__ row_ptr,
                                                          const int32_t* __restrict__ col, int64_t I, int64_t words,
                                                          uint32_t* __restrict__ bits) {
  const int r = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const int64_t uid = users ? users[r] : r;
  const int64_t p0 = row_ptr[uid], p1 = row_ptr[uid + 1];
  uint32_t* row = bits + (int64_t)r * words;
  for (int64_t p = p0 + lane; p < p1; p += 32) {
    const int it = __ldg(col + p);
    atomicOr(row + (it >> 5), 1u << (it & 31));
  }
  for (int64_t w = (I >> 5) + lane; w < words; w += 32) {
    const int64_t first = w * 32;
    atomicOr(row + w, first >= I ? 0xffffffffu : (0xffffffffu << (int)(I - first)));
  }
}

// ---------------------------------------------------------------------------------------
// Probe pass (cdae_topn_build, tensor-core path).  A sweep that starts from thr = -inf appends ~k ln(I/k)
// candidates per user before its threshold settles (about 220 at config B: nearly every 32-column chunk of
// every warp takes the slow append path, and that — not the MMA — is what the sweep costs).  The items that
// can reach a top-k list are few and largely the same for everybody (high output bias, high score for the
// AVERAGE user), so the build first scores every user against the M items with the largest mean-user score
// z_mean.W'[i] + b'[i] (M = 256..1024, one to four tiles) with the same kernel, takes the k-th best
// approximate score a_k of that pass and starts the real sweep at
//     thr0 = a_k - 2 eps_u - tiny:
// k unrated items have exact score >= a_k - eps_u, so an item with approx <= thr0 has exact
// <= thr0 + eps_u < (k-th best exact score) and cannot be in the list; the verification of the re-rank
// (thr_u + eps_u < k-th best exact) holds for thr0 by construction.  Nothing about the final lists depends
// on the probe being good: a poor probe only means a lower start threshold.
__global__ void __launch_bounds__(256) probe_colsum_kernel(const float* __restrict__ Z, const int32_t* __restrict__ users,
                                                           int n, int ld, int users_per_block, float* __restrict__ out) {
  const int r0 = blockIdx.x * users_per_block, r1 = min(n, r0 + users_per_block);
  for (int c = threadIdx.x; c < ld; c += blockDim.x) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += Z[(int64_t)(users ? users[r] : r) * ld + c];
    atomicAdd(out + c, s);
  }
}
// key[i] = b'[i] + W'[i].zsum * inv_n, ids[i] = i; one warp per item
__global__ void __launch_bounds__(256) probe_key_kernel(const float* __restrict__ W, const float* __restrict__ bp, int64_t I,
                                                        int K, int ld, const float* __restrict__ zsum, float inv_n,
                                                        float* __restrict__ keys, int32_t* __restrict__ ids) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= I) return;
  float s = 0.f;
  for (int c = lane; c < K; c += 32) s = fmaf(W[i * ld + c], zsum[c], s);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
    const float k = fmaf(s, inv_n, bp[i]);
    keys[i] = k == k ? k : -INFINITY;   // (a NaN key would poison the sort order)
    ids[i] = (int32_t)i;
  }
}
// pos_of[item] = its row in the probe table (pos_of is preset to -1)
__global__ void __launch_bounds__(256) probe_pos_kernel(const int32_t* __restrict__ sorted_ids, int M, int32_t* __restrict__ pos_of) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < M) pos_of[sorted_ids[p]] = p;
}
// probe table: row p = packed row of item sorted_ids[p]; one thread per 16 bytes
__global__ void __launch_bounds__(256) probe_gather_w_kernel(const __nv_bfloat16* __restrict__ wb, const int32_t* __restrict__ sorted_ids,
                                                             int M, int Kp, __nv_bfloat16* __restrict__ out) {
  const int g8 = Kp / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)M * g8) return;
  const int p = (int)(idx / g8), c0 = (int)(idx % g8) * 8;
  *reinterpret_cast<uint4*>(out + (int64_t)p * Kp + c0) =
      *reinterpret_cast<const uint4*>(wb + (int64_t)sorted_ids[p] * Kp + c0);
}
// rated bitmap over the probe table's rows: bit p of row r <=> probe item p is in the train row of user r
__global__ void __launch_bounds__(256) probe_bitmap_kernel(const int32_t* __restrict__ users, int n,
                                                           const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                                           const int32_t* __restrict__ pos_of, int64_t words,
                                                           uint32_t* __restrict__ bits) {
  const int r = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const int64_t uid = users ? users[r] : r;
  const int64_t p0 = row_ptr[uid], p1 = row_ptr[uid + 1];
  uint32_t* row = bits + (int64_t)r * words;
  for (int64_t p = p0 + lane; p < p1; p += 32) {
    const int q = __ldg(pos_of + __ldg(col + p));
    if (q >= 0) atomicOr(row + (q >> 5), 1u << (q & 31));
  }
}
// thr0[r] = (k-th largest approximate score of user r's probe candidates) - 2 eps_r - tiny, or -inf when the
// probe pass holds fewer than k candidates for the user.  One thread per user.
__global__ void __launch_bounds__(256) probe_thr_kernel(const float* __restrict__ cand_s, const int* __restrict__ cand_cnt,
                                                        const float* __restrict__ eps, int n, int seg, int topk,
                                                        float* __restrict__ thr0) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int cnt = cand_cnt[r];
  float out = -INFINITY;
  if (cnt >= topk) {
    // k-th largest by repeated "largest value below the previous one" (at most k passes over <= CAND_MAX entries)
    const float* s = cand_s + (int64_t)r * seg;
    float prev = INFINITY, kth = -INFINITY;
    int seen = 0;
    for (int pass = 0; pass < topk && seen < topk; ++pass) {
      float m = -INFINITY;
      int c = 0;
      for (int x = 0; x < cnt; ++x) {
        const float v = s[x];
        if (v < prev) {
          if (v > m) { m = v; c = 1; }
          else if (v == m) ++c;
        }
      }
      if (c == 0) break;
      seen += c;
      prev = m;
      kth = m;
    }
    if (seen >= topk && kth > -INFINITY) out = kth - 2.f * eps[r] - 1e-6f * fabsf(kth) - 1e-30f;
  }
  thr0[r] = out;
}

// Compaction of the per-user candidate buffers, called by a whole epilogue warp with every lane
// working on its own row: raise thr to a value t with KEEP_LO <= #{s > t} <= KEEP_HI (bisection on
// the value) and drop the entries <= t.  thr never decreases, so "every non-candidate has
// approx <= thr" is preserved.  If the bisection cannot separate (many equal scores) the cut goes
// ABOVE the tie group: fewer than KEEP_LO entries survive and the verification step later sends
// that user to the exact path.
template <int C, int KEEP_LO, int KEEP_HI>
__device__ __noinline__ void compact(float* bs, int* bi, int* cnt_io, float* thr_io) {
  int cnt = *cnt_io;
  float thr = *thr_io;
  // all C slots are read unconditionally (independent loads, pipelined); slots >= cnt are masked
  float mn = INFINITY, hi = -INFINITY, sum = 0.f;
#pragma unroll 8
  for (int e = 0; e < C; ++e) {
    const float s = bs[e * TILE_U];
    if (e < cnt) {
      mn = fminf(mn, s);
      hi = fmaxf(hi, s);
      sum += s;
    }
  }
  const bool need = cnt > KEEP_HI;
  bool done = !need;
  // bracket: count(s > lo) > KEEP_HI  and  count(s > hi) = 0 < KEEP_LO
  float lo = (thr == -INFINITY) ? mn - 1.f : thr;
  float tnew = thr;
  // First pivot: the entries are the upper tail of the score distribution above `lo`, roughly
  // exponential with mean excess (mean - lo), so a fraction f survives a cut at
  // lo + (mean - lo) * ln(1/f); aim at the middle of the keep window.  Then plain bisection.
  const float f = 0.5f * (float)(KEEP_LO + KEEP_HI) / (float)max(cnt, 1);
  float mid = lo + (sum / (float)max(cnt, 1) - lo) * __logf(1.f / f);
  if (!(mid > lo && mid < hi)) mid = 0.5f * (lo + hi);
  for (int it = 0; it < 28; ++it) {
    if (__all_sync(0xffffffffu, done)) break;
    int k = 0;
#pragma unroll 8
    for (int e = 0; e < C; ++e) k += (e < cnt) & (bs[e * TILE_U] > mid);
    if (!done) {
      if (k > KEEP_HI) lo = mid;
      else if (k < KEEP_LO) hi = mid;
      else { tnew = mid; done = true; }
    }
    mid = 0.5f * (lo + hi);
  }
  if (need && !done) tnew = hi;
  if (tnew > thr) {
    int w = 0;
#pragma unroll 4
    for (int e = 0; e < C; ++e) {
      const float s = bs[e * TILE_U];
      const int id = bi[e * TILE_U];
      if (e < cnt && s > tnew) {
        bs[w * TILE_U] = s;
        bi[w * TILE_U] = id;
        ++w;
      }
    }
    cnt = w;
    thr = tnew;
  }
  *cnt_io = cnt;
  *thr_io = thr;
}

// ---------------------------------------------------------------------------------------
struct TcArgs {
  int n_users;                 // valid rows of Zb
  int64_t I;                   // valid items
  int n_tiles;                 // ceil(I / 256)
  const int32_t* users;        // nullable: row r of Zb is user users[r]
  const int64_t* row_ptr;      // train CSR (rated items)
  const int32_t* col;
  int* cand_id;                // [n_users][CAND_MAX]
  float* cand_s;               // [n_users][CAND_MAX]  approximate scores
  int* cand_cnt;               // [n_users]
  float* cand_thr;             // [n_users]  final threshold (every non-candidate has approx <= thr)
  const float* init_thr;       // nullable [n_users]: start threshold of each user (second pass: the
                               // first pass proved that nothing at or below it can be in the top-k)
  const uint32_t* bits;        // nullable: rated bitmap [users_pad][words] built beforehand (rows with many rated
  int64_t words;               //   items per tile: the in-kernel CSR walk is one dependent load per item); pad columns set
  int n_splits;                // gridDim.y: the item tiles are cut into this many contiguous ranges,
  int tiles_per_split;         //   one CTA per (user tile, range); candidates land in per-range
  int seg;                     //   segments of `seg` = CAND_MAX / n_splits slots: cand_*[(u*S + s)*seg ..]
};

// One 32-column chunk of one accumulator tile for this thread's user: fast reject by the chunk
// maximum, else scan the 8-column groups that hold a score above thr.  Warp-uniform control flow
// (votes), so a compaction can run for the whole warp between groups.
// One 32-column chunk of one accumulator tile for this thread's user: fast reject by the chunk
// maximum, else look at the 8-column groups that hold a score above thr.  Warp-uniform control flow
// (votes), so a compaction can run for the whole warp between groups.
//
// A group with a hit is almost always hit by ONE lane (one user), so predicating an 8-column scan
// over the whole warp wastes 31/32 of the work: the sparse path instead lets the hit lane publish
// its 8 scores through `xch` (32 bytes of shared memory per warp) and lanes 0-7 test one column each
// and append, cooperatively, to THAT user's buffer.  Dense groups (start of a sweep) keep the
// predicated scan.
template <int C2, int KEEP_LO, int KEEP_HI>
__device__ __forceinline__ void scan_chunk(const uint32_t (&v)[32], uint32_t rated, int item_base, float* bs,
                                           int* bi, int& cnt, float& thr, float* xch, int lane) {
  float gm[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float x0 = max3(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]), __uint_as_float(v[g * 8 + 2]));
    const float x1 = max3(__uint_as_float(v[g * 8 + 3]), __uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
    gm[g] = max3(x0, x1, fmaxf(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7])));
  }
  const float m = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
  if (!__any_sync(0xffffffffu, m > thr)) return;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    unsigned hm = __ballot_sync(0xffffffffu, gm[g] > thr);
    if (hm == 0) continue;
    // invariant here: cnt <= C2 - 8
    if (__popc(hm) > 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = __uint_as_float(v[g * 8 + j]);
        if (s > thr && !((rated >> (g * 8 + j)) & 1u)) {
          bs[cnt * TILE_U] = s;
          bi[cnt * TILE_U] = item_base + g * 8 + j;
          ++cnt;
        }
      }
    } else {
      while (hm) {
        const int L = __ffs(hm) - 1;
        hm &= hm - 1;
        if (lane == L) {
          *reinterpret_cast<uint4*>(xch) = make_uint4(v[g * 8 + 0], v[g * 8 + 1], v[g * 8 + 2], v[g * 8 + 3]);
          *reinterpret_cast<uint4*>(xch + 4) = make_uint4(v[g * 8 + 4], v[g * 8 + 5], v[g * 8 + 6], v[g * 8 + 7]);
        }
        const float thrL = __shfl_sync(0xffffffffu, thr, L);
        const int cntL = __shfl_sync(0xffffffffu, cnt, L);
        const uint32_t ratedL = __shfl_sync(0xffffffffu, rated, L);
        __syncwarp();
        const int j = lane & 7;
        const float s = xch[j];
        const bool ok = lane < 8 && s > thrL && !((ratedL >> (g * 8 + j)) & 1u);
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const int slot = cntL + __popc(okm & ((1u << j) - 1u));
          // bs / bi point at this lane's own row: lane L's row is (L - lane) floats away
          bs[slot * TILE_U + (L - lane)] = s;
          bi[slot * TILE_U + (L - lane)] = item_base + g * 8 + j;
        }
        if (lane == L) cnt += __popc(okm);
        __syncwarp();   // xch is reused by the next hit lane
      }
    }
    if (__any_sync(0xffffffffu, cnt > C2 - 8)) compact<C2, KEEP_LO, KEEP_HI>(bs, bi, &cnt, &thr);
  }
}

template <int KB>
__global__ void __launch_bounds__(n_threads(KB), 1) topn_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                 const __grid_constant__ CUtensorMap map_b, TcArgs a) {
  constexpr int C = cand_slots(KB), EPI = epi_warps(KB), C2 = buf_slots(KB);
  constexpr int KEEP_LO = keep_lo(KB), KEEP_HI = keep_hi(KB);
  constexpr int SOFT = (C2 - KEEP_HI) / 2;   // ask for the shared compaction when fewer slots are free
  static_assert(KEEP_HI <= C2 - 8 && KEEP_LO >= 10 && KEEP_LO < KEEP_HI, "candidate buffer geometry");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                                   // KB x [128][64] bf16, swizzled
  unsigned char* sB = sA + KB * A_BLK_BYTES;                  // NSTAGE x [256][64] bf16
  float* cs = reinterpret_cast<float*>(sB + NSTAGE * B_BLK_BYTES);  // [EPI][C2][128]
  int* ci = reinterpret_cast<int*>(cs + C * TILE_U);               // [EPI][C2][128]
  uint32_t* bm = reinterpret_cast<uint32_t*>(ci + C * TILE_U);     // [2][8][128]
  constexpr int EPW = EPI < 2 ? 2 : EPI;                            // rows of the merge arrays
  int* mcnt = reinterpret_cast<int*>(bm + 2 * 8 * TILE_U);          // [EPW][128] final counts per buffer
  float* mthr = reinterpret_cast<float*>(mcnt + EPW * TILE_U);      // [EPW][128] final thresholds
  volatile int* creq = reinterpret_cast<volatile int*>(mthr + EPW * TILE_U);  // [2] tile that asked for a compaction
  float* xch_all = mthr + EPW * TILE_U + 4;                          // [4 * EPW epilogue warps][8] hit exchange (16-byte aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch_all + 4 * EPW * 8);
  uint64_t* full = bars;                 // [NSTAGE] TMA -> MMA
  uint64_t* empty = bars + NSTAGE;       // [NSTAGE] MMA -> TMA
  uint64_t* a_full = bars + 2 * NSTAGE;  // A tile landed
  uint64_t* t_full = a_full + 1;         // [2] accumulator ready      MMA -> epilogue
  uint64_t* t_empty = t_full + 2;        // [2] accumulator drained    epilogue -> MMA
  uint64_t* b_full = t_empty + 2;        // [2] bitmap ready           helpers -> epilogue
  uint64_t* b_empty = b_full + 2;        // [2] bitmap consumed        epilogue -> helpers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * TILE_U;
  const int sp = blockIdx.y;                                   // item range of this CTA
  const int t_lo = sp * a.tiles_per_split;
  const int n_t = max(0, min(a.n_tiles, t_lo + a.tiles_per_split) - t_lo);   // tiles of this CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(a_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(t_full + b, 1);
      mbar_init(t_empty + b, 4 * EPI);  // one arrive per epilogue warp
      mbar_init(b_full + b, 64);        // every helper thread
      mbar_init(b_empty + b, 4 * EPI);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (threadIdx.x == 0) creq[0] = creq[1] = -1;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(a_full, KB * A_BLK_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sA + kb * A_BLK_BYTES, &map_a, a_full, kb * KBLK, u0);
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_t; ++t) {
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(empty + s, ph ^ 1);
          mbar_arrive_expect_tx(full + s, B_BLK_BYTES);
          tma_load_2d(sB + s * B_BLK_BYTES, &map_b, full + s, kb * KBLK, (t_lo + t) * TILE_I);
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
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
#pragma unroll
          for (int k = 0; k < KBLK / 16; ++k)  // 16 bf16 = 32 bytes along K inside the swizzle row
            umma_bf16(d, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, (kb | k) != 0);
          umma_commit(empty + s);  // frees the B stage once these MMAs have read it
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
        umma_commit(t_full + buf);
      }
    }
  } else if (warp < 4 && a.bits == nullptr) {
    // ===== rated-item bitmaps: thread h owns rows h and h + 64 =====
    // A cursor walks each user's ascending train row once per sweep.  `cur` is the next rated item,
    // `nxt` the one after it, loaded one step AHEAD so that the tile loop never waits on a
    // dependent global load (the first version did: ~2 L2 round trips per rated item made these
    // two warps, not the MMA or the epilogue, the pace of the whole kernel — profiles/r01_c_*).
    const int h = threadIdx.x - 64;
    const int32_t* rowp[2];
    int pos[2], end[2], cur[2], nxt[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int r = h + q * 64;
      rowp[q] = a.col;
      pos[q] = end[q] = 0;
      cur[q] = nxt[q] = 0x7fffffff;
      if (u0 + r < a.n_users) {
        const int64_t uid = a.users ? a.users[u0 + r] : (u0 + r);
        const int64_t p0 = a.row_ptr[uid], p1 = a.row_ptr[uid + 1];
        rowp[q] = a.col + p0;
        end[q] = (int)(p1 - p0);
        if (t_lo > 0) {  // first rated item inside this CTA's item range
          int lo = 0, hi = end[q];
          const int first = t_lo * TILE_I;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (rowp[q][mid] < first) lo = mid + 1; else hi = mid;
          }
          pos[q] = lo;
        }
        if (pos[q] < end[q]) cur[q] = rowp[q][pos[q]];
        if (pos[q] + 1 < end[q]) nxt[q] = rowp[q][pos[q] + 1];
      }
    }
    for (int t = 0; t < n_t; ++t) {
      const int buf = t & 1;
      mbar_wait(b_empty + buf, ((t >> 1) & 1) ^ 1);
      const int i0 = (t_lo + t) * TILE_I, i1 = i0 + TILE_I;
      uint32_t* my = bm + buf * 8 * TILE_U;
      uint32_t w[2][8];
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int c = 0; c < 8; ++c) w[q][c] = 0u;
      if ((int64_t)i1 > a.I) {
        // columns that are not items at all (the padded tail of the last tile)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int64_t first = (int64_t)i0 + c * 32;
          const uint32_t mask = first + 32 <= a.I ? 0u : (first >= a.I ? 0xffffffffu : (0xffffffffu << (int)(a.I - first)));
          w[0][c] = w[1][c] = mask;
        }
      }
      while (cur[0] < i1 || cur[1] < i1) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (cur[q] < i1) {
            const int rel = cur[q] - i0;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if ((rel >> 5) == c) w[q][c] |= 1u << (rel & 31);
            cur[q] = nxt[q];
            ++pos[q];
            nxt[q] = pos[q] + 1 < end[q] ? rowp[q][pos[q] + 1] : 0x7fffffff;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int c = 0; c < 8; ++c) my[c * TILE_U + h + q * 64] = w[q][c];
      mbar_arrive(b_full + buf);
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = one user (TMEM lane); warp pair member `hf` takes chunks hf, hf+EPI, ..
    const int e = warp - 4;
    const int q = e & 3;                    // TMEM lane quadrant (= warp % 4)
    const int hf = e >> 2;                  // which buffer / which chunks
    const int row = q * 32 + lane;
    float* bs = cs + hf * C2 * TILE_U + row;   // entry e at bs[e * 128]
    int* bi = ci + hf * C2 * TILE_U + row;
    float* xch = xch_all + e * 8;
    float thr = (a.init_thr && u0 + row < a.n_users) ? a.init_thr[u0 + row] : -INFINITY;
    int cnt = 0;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int NCH = 8 / EPI;            // chunks per tile for this warp

    // prepass-bitmap mode: this row's 256 rated bits of a tile are 32 contiguous bytes; tile t + 1 is
    // fetched at the end of tile t's work
    const bool use_bits = a.bits != nullptr;
    const uint4* brow = use_bits ? reinterpret_cast<const uint4*>(a.bits + (int64_t)(u0 + row) * a.words) + (int64_t)t_lo * 2 : nullptr;
    uint4 nb0 = make_uint4(0u, 0u, 0u, 0u), nb1 = nb0;
    if (use_bits && n_t > 0) { nb0 = __ldg(brow); nb1 = __ldg(brow + 1); }
    for (int t = 0; t < n_t; ++t) {
      const int buf = t & 1;
      const uint32_t par = (t >> 1) & 1;
      const uint4 cb0 = nb0, cb1 = nb1;
      mbar_wait(t_full + buf, par);
      if (!use_bits) mbar_wait(b_full + buf, par);
      tc_fence_after();
      const uint32_t* my_bm = bm + buf * 8 * TILE_U + row;
      auto rated_word = [&](int c) -> uint32_t {
        if (!use_bits) return my_bm[c * TILE_U];
        const uint4 q = c < 4 ? cb0 : cb1;
        const int k = c & 3;
        return k == 0 ? q.x : k == 1 ? q.y : k == 2 ? q.z : q.w;
      };
      const int item0 = (t_lo + t) * TILE_I;
      const uint32_t col0 = lane_addr + (uint32_t)(buf * TILE_I);
      // two chunks in flight: the next tcgen05.ld is issued before the current chunk is scanned
      uint32_t va[32], vb[32];
      tmem_ld32_issue(col0 + (uint32_t)(hf * 32), va);
#pragma unroll 1
      for (int i = 0; i < NCH; i += 2) {
        const int c0 = hf + i * EPI, c1 = hf + (i + 1) * EPI;
        tmem_ld_wait(va);
        tmem_ld32_issue(col0 + (uint32_t)(c1 * 32), vb);
        scan_chunk<C2, KEEP_LO, KEEP_HI>(va, rated_word(c0), item0 + c0 * 32, bs, bi, cnt, thr, xch, lane);
        tmem_ld_wait(vb);
        if (i + 2 < NCH) tmem_ld32_issue(col0 + (uint32_t)((c1 + EPI) * 32), va);
        scan_chunk<C2, KEEP_LO, KEEP_HI>(vb, rated_word(c1), item0 + c1 * 32, bs, bi, cnt, thr, xch, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(t_empty + buf);
        if (!use_bits) mbar_arrive(b_empty + buf);
      }
      if (use_bits && t + 1 < n_t) { nb0 = __ldg(brow + 2 * (t + 1)); nb1 = __ldg(brow + 2 * (t + 1) + 1); }
      // Compactions are taken TOGETHER by all epilogue warps at a tile boundary: a compaction costs
      // about as much as scanning a whole tile, and with only two accumulator buffers a warp
      // that compacts on its own stalls the MMA and, through it, the seven other warps — ~150
      // separate stalls per sweep in the first version (profiles/r01_c_*), ~25 shared ones now.
      if (cnt > C2 - SOFT) creq[t & 1] = t;
      named_bar_sync(2, 128 * EPI);
      if (creq[t & 1] == t) compact<C2, KEEP_LO, KEEP_HI>(bs, bi, &cnt, &thr);
    }
    // ---- merge the buffers of the warp pair and write the candidates out
    mcnt[hf * TILE_U + row] = cnt;
    mthr[hf * TILE_U + row] = thr;
    named_bar_sync(1, 128 * EPI);
    if (u0 + row < a.n_users) {
      int off = 0, total = 0;
      float tmax = -INFINITY;
#pragma unroll
      for (int x = 0; x < EPI; ++x) {
        if (x < hf) off += mcnt[x * TILE_U + row];
        total += mcnt[x * TILE_U + row];
        tmax = fmaxf(tmax, mthr[x * TILE_U + row]);
      }
      const int64_t slot = (int64_t)(u0 + row) * a.n_splits + sp;
      const bool fits = total <= a.seg;
      if (fits) {
        const int64_t o = slot * a.seg + off;
        for (int x = 0; x < cnt; ++x) {
          a.cand_id[o + x] = bi[x * TILE_U];
          a.cand_s[o + x] = bs[x * TILE_U];
        }
      }
      if (hf == 0) {
        // every non-candidate of this item range has approx <= the threshold of the buffer that
        // saw it; a segment that cannot hold the candidates voids the user's proof (thr = +inf)
        a.cand_cnt[slot] = fits ? total : 0;
        a.cand_thr[slot] = fits ? tmax : INFINITY;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc
}  // namespace cdae
static_assert(cdae::tc::smem_bytes(1) <= 232448 && cdae::tc::smem_bytes(2) <= 232448 && cdae::tc::smem_bytes(3) <= 232448 &&
                  cdae::tc::smem_bytes(4) <= 232448 && cdae::tc::smem_bytes(5) <= 232448,
              "topn_tc_kernel shared memory exceeds the 227 KB a CTA can opt into");
