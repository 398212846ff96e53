// cdae_b200/csrc/train_kernels.cuh — the per-user CDAE step as sm_100a kernels.
//
// One frozen minibatch of users (SURVEY.md Appendix A "frozen-batch") runs as
//   gather(+mask) -> activate -> decode(+negatives) -> hidden_backward -> scatter -> apply
//
// Layout: all item/user tables are row-major [rows][ld] fp32 with ld = a whole number of
// 128-byte lines that one <G,NV> geometry covers EXACTLY (ld = 4*G*NV: 32, 64, 96, 128, 192, 256,
// 384 or 512 floats, row_stride() in api.cu; pad columns are kept at exactly 0), so no kernel
// needs a column bound check for correctness.  gather / scatter skip the 16-byte pieces that lie
// ENTIRELY in the padding (14 of 64 columns at K = 50, 56 of 256 at K = 200: loads that return 0,
// reductions that add 0) with a per-lane predicate; in decode_kernel the same predicate costs more
// issue slots than the L2 transactions it saves (36.3 -> 35.1 M users/s at config B), so it is not used there.  A row is owned
// by a GROUP of G >= 8 lanes, each holding NV <= 4 float4 (column 4*(v*G+lane) .. +3), so ONE
// warp-level 16-byte load/reduction covers whole 128-byte lines of 32/G rows.  History
// (profiles/r01_*): v1 used 224-byte rows with 16 lanes per row and was issue-bound; v2 used 4
// lanes per row and was bound by the L1->crossbar REQUEST rate (66% busy, ~10 requests per
// row because 64-byte pieces of unaligned rows straddle lines); line-aligned rows with 8 lanes
// need 2 read + 2 reduction requests per row at K=50.
//
// Sampling (corruption mask, negatives) is fused into gather / decode: it is pure integer ALU
// work that overlaps with the memory stalls of those kernels, and the negatives never touch
// global memory.  Reference lines each kernel takes over are cited at the kernel.
#pragma once
#include "common.cuh"

namespace cdae {

struct ModelDev {
  // parameters and AdaGrad state (cdae.hpp:430-439), fp32, leading dimension ld
  float *W, *V, *Wu, *b, *bp, *Uu;
  float *W_ag, *V_ag, *Wu_ag, *b_ag, *bp_ag, *Uu_ag;
  // dense gradient accumulators of one minibatch (one contiguous buffer, all-reduced as one)
  float *gW, *gV, *gbp, *gb;
  float* gcnt;     // [2][I4] how often item j was a kept input in this minibatch (the lambda*W[j] terms of
                   // cdae.hpp:333-349 are added as lambda*count*W[j] by apply_kernel); slot = minibatch parity
  int64_t I4;
  float* g_steps;  // [2] user steps that contributed (n*lambda*b term); slot = minibatch parity
  int steps_slot;
  int64_t I, U;
  int K, ld;
  float lambda, lr, beta, scale;  // scale = scaled ? 1/(1-q) : 1   (cdae.hpp:202-205)
  int loss, nu;
  int adagrad, asym, user_factor, linear, tanh_act, linear_function;
  int direct_lambda;  // 1: scatter_kernel adds lambda*W[j] itself (reads the row) instead of counting occurrences for
                      // the optimiser pass — process groups: the fused combine step then needs no per-row counts
};

struct BatchDev {
  const WorkItem* in_items;   // input chunks  (<= 64 slots each)
  const WorkItem* out_items;  // output chunks (<= ch_out slots each)
  int n_in_items, n_out_items, n_users;
  const int32_t* uids;        // [n_users] global uid per local row
  const int64_t* row_ptr;     // CSR of the training set (device)
  const int32_t* col;
  uint8_t* keep;              // [minibatch slots]       1 = input item survives corruption
  const int32_t* negs;        // [minibatch slots * nu]  explicit negatives (unsampled mode only)
  float *H, *Z, *HG, *D, *GU;  // [n_users][ld]
  int ch_in;                  // slots per input chunk (64)
  int flags;                  // BATCH_BY_UID: rows of H / Z are indexed by GLOBAL uid (whole-table
                              // encode); BATCH_KEEP_ALL: no corruption mask, every input item is kept
};
enum { BATCH_BY_UID = 1, BATCH_KEEP_ALL = 2, BATCH_FUSED_ENCODE = 4 };

// Counter-based sampling parameters (same specification as oracle/cdae_oracle.h).
struct SampleArgs {
  uint64_t seed;
  uint32_t pass;       // epoch * num_corruptions + corruption index
  uint32_t keep_thr;   // floor(q * 2^32)
  int keep_mode;       // 0 keep all (q <= 0), 1 keep none (q >= 1), 2 Philox
};

constexpr int STAT_STRIPES = 64;  // same-address atomics serialise in L2: spread them
struct StatsDev {
  double loss_sum[STAT_STRIPES];
  unsigned long long outputs[STAT_STRIPES];
  unsigned long long inputs_kept[STAT_STRIPES];
  unsigned long long user_steps;
  int bad_loss;  // LOGISTIC fed a score outside (0,1)
  int bad_csr;   // validate_rows_kernel found an item id out of range or an unsorted row
};

template <int G, int NV>
struct RowMap {
  static constexpr int NG = 32 / G;               // rows a warp handles at once
  static constexpr int UNR = NV <= 2 ? 4 : 2;     // row batches in flight per warp
  __device__ static __forceinline__ int col4(int gl, int v) { return (v * G + gl) * 4; }
};

template <int G>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}
// sum a per-group partial vector over the 32/G groups of the warp (result valid in all lanes)
template <int G>
__device__ __forceinline__ float4 cross_group_sum(float4 a) {
#pragma unroll
  for (int off = G; off < 32; off <<= 1) {
    a.x += __shfl_xor_sync(0xffffffffu, a.x, off);
    a.y += __shfl_xor_sync(0xffffffffu, a.y, off);
    a.z += __shfl_xor_sync(0xffffffffu, a.z, off);
    a.w += __shfl_xor_sync(0xffffffffu, a.w, off);
  }
  return a;
}

__device__ __forceinline__ WorkItem load_item(const WorkItem* p) {
  WorkItem w;
  const int4 a = __ldg(reinterpret_cast<const int4*>(p));
  const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
  w.uid = a.x; w.u_local = a.y; w.n = a.z; w.aux0 = a.w;
  w.s0 = (int64_t)(((uint64_t)(uint32_t)b.y << 32) | (uint32_t)b.x);
  w.first = b.z; w.row_off = b.w;
  return w;
}

// membership test in an ascending CSR row
__device__ __forceinline__ bool row_contains(const int32_t* row, int n, int item) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(row + mid) < item) lo = mid + 1; else hi = mid;
  }
  return lo < n && __ldg(row + lo) == item;
}

// ---------------------------------------------------------------------------------------
// Device-side check of a CSR that arrived through cdae_train_epoch_csr (one warp per trained user):
// item ids must lie in [0, I) and rows must be strictly ascending (the negative sampler and the
// top-N exclusion binary-search them).  Out-of-range ids are clamped in the device copy, so that no
// later kernel of the call can index outside a table; any violation sets stats->bad_csr (read by the
// host after the call) and *flag (part of the all-reduced gradient buffer: every rank of a process
// group sees it), which turn hidden_backward / uu_update / apply into no-ops.
__global__ void __launch_bounds__(256) validate_rows_kernel(const int32_t* uids, int64_t n_users,
                                                            const int64_t* row_ptr, int32_t* col, int64_t I,
                                                            StatsDev* stats, float* flag) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_users) return;
  const int lane = threadIdx.x & 31;
  const int64_t uid = uids[w];
  const int64_t s0 = row_ptr[uid], s1 = row_ptr[uid + 1];
  bool bad = false;
  for (int64_t s = s0 + lane; s < s1; s += 32) {
    const int32_t c = col[s];
    if (c < 0 || (int64_t)c >= I) {
      bad = true;
      col[s] = 0;
    } else if (s > s0 && col[s - 1] >= c) {
      bad = true;
    }
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) {
    stats->bad_csr = 1;
    red_add_f32(flag, 1.f);
  }
}

// ---------------------------------------------------------------------------------------
// H3, cdae.hpp:361-371: an input item is kept iff uniform > q.  Slot s of the user's row takes
// word (s&3) of philox({uid, s>>2, pass, 0}); chunks start at multiples of 64 slots, so slots
// 4q..4q+3 of a chunk share one call.  Warp-cooperative; returns this lane's kept count.
__device__ __forceinline__ int sample_keep_chunk(const WorkItem& wi, const SampleArgs& sa,
                                                 uint8_t* keep, int lane) {
  int kept = 0;
  for (int q = lane; q * 4 < wi.n; q += 32) {
    Philox4 p = {0u, 0u, 0u, 0u};
    if (sa.keep_mode == 2)
      p = philox4x32(sa.seed, (uint32_t)wi.uid, (uint32_t)((wi.row_off >> 2) + q), sa.pass, 0u);
    const int lim = min(4, wi.n - q * 4);
    for (int i = 0; i < lim; ++i) {
      uint8_t k;
      if (sa.keep_mode == 0) k = 1;
      else if (sa.keep_mode == 1) k = 0;
      else k = philox_word(p, i) > sa.keep_thr;
      keep[q * 4 + i] = k;
      kept += (int)k;
    }
  }
  return kept;
}

// H5, recsys_model_base.hpp:46-57 + cdae.hpp:217-220: n_u*num_neg draws with replacement,
// uniform over items, redrawn while the item is one of the user's positives.  Draw d of the user
// takes word (d&3) of philox({uid, d>>2, pass, 1}) — one call serves four draws — and, only if
// that item is a positive (probability n_u/I), retries a = 1,2,.. from word ((a-1)&3) of
// philox({uid, d, pass, 2 + ((a-1)>>2)}).  Warp-cooperative: fills out[0 .. n*nu).
__device__ __forceinline__ void sample_negs_chunk(const WorkItem& wi, const SampleArgs& sa, int nu,
                                                  int64_t I, const int64_t* row_ptr,
                                                  const int32_t* col, int32_t* out, int lane) {
  const int ndraw = wi.n * nu;
  if (ndraw == 0) return;
  const int64_t r0 = __ldg(row_ptr + wi.uid);
  const int n_u = (int)(__ldg(row_ptr + wi.uid + 1) - r0);
  const int32_t* row = col + r0;
  const uint32_t d_base = (uint32_t)(wi.row_off * nu);
  const int d_head = (int)((4u - (d_base & 3u)) & 3u);  // draws before the first aligned quad
  const int nquad = (d_head > 0 ? 1 : 0) + (ndraw - min(d_head, ndraw) + 3) / 4;
  for (int q = lane; q < nquad; q += 32) {
    int j0, j1;  // [j0, j1) = this quad's draws inside the chunk
    if (d_head > 0) {
      j0 = q == 0 ? 0 : d_head + (q - 1) * 4;
      j1 = q == 0 ? d_head : j0 + 4;
    } else {
      j0 = q * 4;
      j1 = j0 + 4;
    }
    j1 = min(j1, ndraw);
    if (j0 >= j1) continue;
    const Philox4 p0 = philox4x32(sa.seed, (uint32_t)wi.uid, (d_base + (uint32_t)j0) >> 2, sa.pass, 1u);
    for (int j = j0; j < j1; ++j) {
      const uint32_t d = d_base + (uint32_t)j;
      int32_t item = (int32_t)(((uint64_t)philox_word(p0, d & 3) * (uint64_t)I) >> 32);
      for (uint32_t a = 1; row_contains(row, n_u, item); ++a) {
        const Philox4 p = philox4x32(sa.seed, (uint32_t)wi.uid, d, sa.pass, 2u + ((a - 1) >> 2));
        item = (int32_t)(((uint64_t)philox_word(p, (a - 1) & 3) * (uint64_t)I) >> 32);
      }
      out[j] = item;
    }
  }
}

// ---------------------------------------------------------------------------------------
// H3 + H4 first half, cdae.hpp:361-380: (SAMPLED: draw the chunk's corruption mask, store it for
// decode / scatter) then H[u] += sum over kept inputs of W[item]  (the scale factor is applied in
// activate_kernel).  One warp per input chunk; chunks of one user add up with 16-byte
// reductions (H is zeroed per minibatch).
template <int G, int NV, bool SAMPLED>
__global__ void __launch_bounds__(256) gather_kernel(ModelDev m, BatchDev bt, SampleArgs sa,
                                                     StatsDev* stats) {
  using RM = RowMap<G, NV>;
  constexpr int NG = RM::NG, UNR = RM::UNR;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= bt.n_in_items) return;
  const int lane = threadIdx.x & 31, grp = lane / G, gl = lane % G;
  const WorkItem wi = load_item(bt.in_items + warp);
  const int32_t* items = bt.col + wi.s0;
  uint8_t* keep = bt.keep + wi.aux0;
  if (SAMPLED) {
    int kept = sample_keep_chunk(wi, sa, keep, lane);
    kept = (int)group_sum<32>((float)kept);
    if (lane == 0 && stats && kept)
      atomicAdd(&stats->inputs_kept[warp % STAT_STRIPES], (unsigned long long)kept);
    __syncwarp();  // the mask bytes written above are read by other lanes below
  }
  const bool keep_all = !SAMPLED && (bt.flags & BATCH_KEEP_ALL);
  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = f4zero();
  for (int base = 0; base < wi.n; base += NG * UNR) {
    int it[UNR];
#pragma unroll
    for (int t = 0; t < UNR; ++t) {
      const int r = base + t * NG + grp;
      it[t] = (r < wi.n && (keep_all || keep[r])) ? __ldg(items + r) : -1;
    }
    float4 w[UNR][NV];
#pragma unroll
    for (int t = 0; t < UNR; ++t)
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = RM::col4(gl, v);
        w[t][v] = (it[t] >= 0 && c < m.K) ? ld4(m.W + (int64_t)it[t] * m.ld + c) : f4zero();
      }
#pragma unroll
    for (int t = 0; t < UNR; ++t)
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] = add4(acc[v], w[t][v]);
  }
  const int64_t hrow = (bt.flags & BATCH_BY_UID) ? (int64_t)wi.uid : (int64_t)wi.u_local;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    acc[v] = cross_group_sum<G>(acc[v]);
    const int c = RM::col4(gl, v);
    if (grp == 0) red_add_v4(bt.H + hrow * m.ld + c, acc[v]);
  }
}

// ---------------------------------------------------------------------------------------
// The north-star form of the encode (H3 + H4 in ONE kernel): corrupt, gather-reduce, bias, user row,
// activation.  One warp per input chunk.  A user whose whole row is one chunk (<= 64 items: ~93 % of the
// users at config B) is finished here and its z is stored; longer rows still add into H and are finished
// by activate_kernel (which skips the users this kernel completed: BATCH_FUSED_ENCODE).
//   TMA = false: rows go straight from L2 into the registers that sum them (gather_kernel's loads);
//   TMA = true : the kept rows are STAGED IN SHARED MEMORY by bulk asynchronous copies (cp.async.bulk,
//                one per row, issued by up to 32 lanes at once, completion on the warp's mbarrier) and
//                summed from there — "TMA-staged W rows in shared memory, warp-shuffle reductions along K".
// A/B at config B in profiles/r02_e_*; api.cu selects with CDAE_B200_ENCODE = split | fused | tma.
constexpr int ENC_STAGE_ROWS = 16;   // rows staged per pass and warp (TMA = true; static shared memory: ld <= 64 only)
__device__ __forceinline__ void mbar_init_w(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_w(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_w(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
      ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(sdst)),
               "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}

template <int G, int NV, bool TMA>
__global__ void __launch_bounds__(256) encode_fused_kernel(ModelDev m, BatchDev bt, SampleArgs sa, StatsDev* stats, float scale) {
  using RM = RowMap<G, NV>;
  constexpr int NG = RM::NG, UNR = RM::UNR, LD = 4 * G * NV;
  __shared__ int32_t kept_s[8][64];
  constexpr bool STAGED = TMA && LD <= 64;     // 8 warps x 16 rows x 256 B = 32 KB of static shared memory
  __shared__ __align__(128) float stage_s[STAGED ? 8 : 1][STAGED ? ENC_STAGE_ROWS : 1][STAGED ? LD : 4];
  __shared__ uint64_t bar_s[8];
  const int wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, grp = lane / G, gl = lane % G;
  if (STAGED) {
    if (lane == 0) mbar_init_w(&bar_s[wib], 1);
    __syncwarp();
  }
  if (warp >= bt.n_in_items) return;
  const WorkItem wi = load_item(bt.in_items + warp);
  const int32_t* items = bt.col + wi.s0;
  uint8_t* keep = bt.keep + wi.aux0;
  int kept = sample_keep_chunk(wi, sa, keep, lane);
  kept = (int)group_sum<32>((float)kept);
  if (lane == 0 && kept) atomicAdd(&stats->inputs_kept[warp % STAT_STRIPES], (unsigned long long)kept);
  __syncwarp();
  // compact the kept item ids (ballot + prefix): the row loop then has no holes
  int32_t* mine = kept_s[wib];
  int nk = 0;
  for (int b0 = 0; b0 < wi.n; b0 += 32) {
    const int r = b0 + lane;
    const bool k = r < wi.n && keep[r];
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    if (k) mine[nk + __popc(bal & ((1u << lane) - 1u))] = __ldg(items + r);
    nk += __popc(bal);
  }
  __syncwarp();
  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = f4zero();
  if (!STAGED) {
    for (int base = 0; base < nk; base += NG * UNR) {
      float4 w[UNR][NV];
#pragma unroll
      for (int t = 0; t < UNR; ++t) {
        const int r = base + t * NG + grp;
#pragma unroll
        for (int v = 0; v < NV; ++v)
          w[t][v] = (r < nk && RM::col4(gl, v) < m.K) ? ld4(m.W + (int64_t)mine[r < nk ? r : 0] * LD + RM::col4(gl, v)) : f4zero();
      }
#pragma unroll
      for (int t = 0; t < UNR; ++t)
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = add4(acc[v], w[t][v]);
    }
  } else {
    const uint32_t row_bytes = (uint32_t)((m.K * 4 + 15) / 16 * 16);
    uint32_t phase = 0;
    for (int base = 0; base < nk; base += ENC_STAGE_ROWS) {
      const int nr = min(ENC_STAGE_ROWS, nk - base);
      if (lane == 0) mbar_expect_tx_w(&bar_s[wib], row_bytes * (uint32_t)nr);
      __syncwarp();
      if (lane < nr) bulk_g2s(&stage_s[wib][lane][0], m.W + (int64_t)mine[base + lane] * LD, row_bytes, &bar_s[wib]);
      mbar_wait_w(&bar_s[wib], phase);
      phase ^= 1u;
      for (int r = grp; r < nr; r += NG)
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (RM::col4(gl, v) < m.K) acc[v] = add4(acc[v], *reinterpret_cast<const float4*>(&stage_s[wib][r][RM::col4(gl, v)]));
      __syncwarp();   // the staging rows are overwritten by the next pass
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = cross_group_sum<G>(acc[v]);
  const bool whole_row = wi.first && (int64_t)wi.n == __ldg(bt.row_ptr + wi.uid + 1) - __ldg(bt.row_ptr + wi.uid);
  if (grp != 0) return;
  if (!whole_row) {
#pragma unroll
    for (int v = 0; v < NV; ++v) red_add_v4(bt.H + (int64_t)wi.u_local * LD + RM::col4(gl, v), acc[v]);
    return;
  }
  // the whole user is in this warp: bias, user row, activation (cdae.hpp:382-414)
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = RM::col4(gl, v);
    float4 h = scale4(scale, acc[v]);
    if (m.linear_function) h = mul4(ld4(m.Uu + (int64_t)wi.uid * LD + c), h);
    h = add4(h, ld4(m.b + c));
    if (m.user_factor) h = add4(h, ld4(m.Wu + (int64_t)wi.uid * LD + c));
    float x[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!m.linear) x[i] = m.tanh_act ? act_tanh(x[i]) : act_sigmoid(x[i]);
      if (c + i >= m.K) x[i] = 0.f;
    }
    st4(bt.Z + (int64_t)wi.u_local * LD + c, make_float4(x[0], x[1], x[2], x[3]));
  }
}

// H4 second half, cdae.hpp:382-414: z = act([Uu (.)] scale*H + b [+ Wu[u]]); pad columns -> 0.
__global__ void __launch_bounds__(256) activate_kernel(ModelDev m, BatchDev bt, float scale) {
  const int ld4n = m.ld / 4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)bt.n_users * ld4n) return;
  const int u = (int)(idx / ld4n), c = (int)(idx % ld4n) * 4;
  const int64_t uid = bt.uids[u];
  // users encode_fused_kernel finished (their whole row was one input chunk) already have their z
  if ((bt.flags & BATCH_FUSED_ENCODE) && bt.row_ptr[uid + 1] - bt.row_ptr[uid] <= bt.ch_in) return;
  const int64_t zrow = (bt.flags & BATCH_BY_UID) ? uid : (int64_t)u;
  float4 h = scale4(scale, ld4(bt.H + zrow * m.ld + c));
  if (m.linear_function) h = mul4(ld4(m.Uu + uid * m.ld + c), h);
  h = add4(h, ld4(m.b + c));
  if (m.user_factor) h = add4(h, ld4(m.Wu + uid * m.ld + c));
  float x[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (!m.linear) x[i] = m.tanh_act ? act_tanh(x[i]) : act_sigmoid(x[i]);
    if (c + i >= m.K) x[i] = 0.f;
  }
  st4(bt.Z + zrow * m.ld + c, make_float4(x[0], x[1], x[2], x[3]));
}

// ---------------------------------------------------------------------------------------
// H5 + H6 + H7, cdae.hpp:217-293 (+ get_output_values :418-426, Loss::gradient): (SAMPLED: draw
// the chunk's negatives into shared memory) then for every output o of the chunk (its positives,
// then their negatives): y = W'[o].z + b'[o]; g = l'(y,t); hg += g*W'[o];
// gW'[o] += g*z + lambda*W'[o]; gb'[o] += g + lambda*b'[o].
// Tied weights and o in the corrupted input: the lambda term is left to scatter_kernel so the
// row receives ONE lambda per merged occurrence (cdae.hpp:249-250, 342-343).
// TRAIN=false scores the positives only and accumulates loss(y,1): CDAE::data_loss :93-96.
// Resident CTAs per SM the register allocation must allow, and row batches in flight per warp.  The kernel
// is latency-bound (dependent chain LDS -> LDG -> dot -> shuffles -> SFU -> RED), so occupancy beats
// per-warp unrolling: measured at config B (profiles/r02_b_*) UNR 4 / 3 CTAs 104.6 us per 8,192 users,
// UNR 2 / 4 CTAs 100.4, UNR 1 / 5 CTAs 99.4, UNR 2 / 5 CTAs (spills) 118.7, UNR 4 / 2 CTAs 119.1; with the
// hand-pipelined loop (DECODE_PIPE) UNR 1 / 4 CTAs 95.6, UNR 2 / 3 CTAs 102.6.
#ifndef DECODE_MIN_BLOCKS
#define DECODE_MIN_BLOCKS 0  // 0 = by geometry: 4 for NV <= 2, else 3
#endif
#ifndef DECODE_PIPE
#define DECODE_PIPE 1      // 1: loads of the next row batch are issued before the current batch is scored (see decode_kernel)
#endif
#ifndef DECODE_STAGE
#define DECODE_STAGE 0     // > 0: rows staged in shared memory by cp.async, that many batches ahead (A/B; see decode_kernel)
#endif
#ifndef DECODE_UNR
#define DECODE_UNR 0       // row batches in flight per warp; 0 = by geometry (pipelined: 2 for NV = 1, else 1 — two batches live in registers)
#endif
constexpr int DECODE_MAX_NEGS = 96;  // ch_out * num_neg <= 96 by construction (api.cu)
constexpr int DECODE_MAX_ROWS = 96;  // outputs of one chunk: n * (1 + num_neg) <= 96

// SFU forms of the CROSS_ENTROPY loss (loss.hpp:132-147): one ex2, one rcp, one lg2 — the library
// __expf / __frcp_rn wrap these in range fix-ups (denormal scaling, a Newton step behind a slow-path
// branch) that cost ~12 issue slots per output for nothing at |y| <= 18 (profiles/r02_a_*).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// dl/dy and l(y, t) for CROSS_ENTROPY, branch-free (same functions as loss_grad's LOSS_CE case)
__device__ __forceinline__ float ce_grad(float y, float t, float* loss) {
  const float a = ex2_approx(-1.4426950408889634f * fabsf(y));   // e^-|y|
  const float d = 1.f + a;
  const float r = rcp_approx(d);
  *loss = fmaf(lg2_approx(d), 0.6931471805599453f, fmaf(1.f - t, y, fmaxf(-y, 0.f)));
  return (y >= 0.f ? r : a * r) - t;
}
// 16-byte reduction issued only where `on` != 0: a predicated instruction, not a branch
__device__ __forceinline__ void red_add_v4_if(float* addr, float4 v, int on) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\t@q red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(addr),
      "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(on)
      : "memory");
}
__device__ __forceinline__ void red_add_f32_if(float* addr, float v, int on) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q red.global.add.f32 [%0], %1;\n\t}" ::"l"(addr), "f"(v), "r"(on)
               : "memory");
}

// cp.async (LDGSTS): 16 bytes global -> shared without passing through registers; per-thread groups
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// LT >= 0 fixes the loss at compile time (CROSS_ENTROPY and SQUARE, the two that are meaningful for
// CDAE); LT = -1 reads m.loss at run time.
//
// Structure (round 2; profiles/r02_a_*: the round-1 kernel spent 122 issue slots per 4-row batch, ~70 of
// them on index plumbing — four BSSY/BRA blocks choosing between the positive list in global memory and
// the negative list in shared memory, 64-bit row-address arithmetic with the run-time leading dimension,
// library exp / rcp fix-ups — and spilled under its 80-register cap):
//   * the chunk's output list (positives, then negatives) is staged ONCE in shared memory, one word per
//     output: item id | merged flag in bit 31; the row loop reads it with one LDS per batch, no branches;
//   * the leading dimension is the compile-time constant 4*G*NV, so a row address is one IMAD.WIDE;
//   * 16-byte pieces that lie entirely in the padding (columns >= K: 3 of 16 at K = 50) are neither
//     loaded nor reduced — a per-lane predicate hoisted out of the loop, applied with predicated
//     instructions (the L2's vector-reduction throughput is this kernel's roofline, bench.py `peak_l2`);
//   * rows past the end of the chunk ride along with g = 0 and their reductions predicated off.
template <int G, int NV, bool TRAIN, bool SAMPLED, int LT>
__global__ void __launch_bounds__(256, DECODE_MIN_BLOCKS > 0 ? DECODE_MIN_BLOCKS : (NV <= 2 ? 4 : 3))
    decode_kernel(ModelDev m, BatchDev bt, SampleArgs sa, StatsDev* stats) {
  using RM = RowMap<G, NV>;
  constexpr int NG = RM::NG, UNR = DECODE_UNR > 0 ? DECODE_UNR : (DECODE_PIPE ? (NV == 1 ? 2 : 1) : (NV == 1 ? 4 : NV <= 3 ? 2 : 1)), LD = 4 * G * NV;
  constexpr bool PIPE = DECODE_PIPE && NV <= 3;   // NV = 4: two batches of 4 float4 per lane do not fit the register cap
  // DECODE_STAGE > 0 (A/B, profiles/r02_l_*): rows travel global -> shared by cp.async, DECODE_STAGE batches ahead,
  // every lane copying exactly the 16-byte pieces it will read back itself (no cross-lane hand-off, so a per-thread
  // cp.async.wait_group is the only synchronisation); nothing is held in registers while in flight.
  constexpr int NST = (DECODE_STAGE > 0 && NV <= 2 && TRAIN) ? DECODE_STAGE : 0;
  __shared__ __align__(16) float4 stage_s[NST > 0 ? 8 : 1][NST > 0 ? NST : 1][NST > 0 ? UNR * NV : 1][NST > 0 ? 32 : 1];
  constexpr int LIST = DECODE_MAX_ROWS + 1 + NG * UNR;   // (num_neg = 96: 97 rows) + one batch of slack
  __shared__ int32_t rows_s[8][LIST];    // the chunk's outputs: positives, then negatives
  __shared__ float lam_s[8][LIST];       // lambda of each output's row term (0 for a merged positive)
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= bt.n_out_items) return;
  const int lane = threadIdx.x & 31, grp = lane / G, gl = lane % G;
  const WorkItem wi = load_item(bt.out_items + warp);
  const int n = wi.n;
  const int R = TRAIN ? n * (1 + m.nu) : n;
  int32_t* mine = rows_s[threadIdx.x >> 5];
  float* lam_mine = lam_s[threadIdx.x >> 5];
  const float lambda = m.lambda;
  {
    // merged occurrence (tied weights, the positive is also a kept input): its lambda term is left to
    // scatter_kernel so the row receives ONE lambda per merged occurrence (cdae.hpp:249-250, 342-343)
    const bool tied = !m.asym;
    const int32_t* pos = bt.col + wi.s0;
    const uint8_t* keep = bt.keep + wi.aux0;
    for (int r = lane; r < R + NG * UNR; r += 32) lam_mine[r] = r < R ? lambda : 0.f;
    __syncwarp();
    for (int r = lane; r < n; r += 32) {
      mine[r] = __ldg(pos + r);
      if (TRAIN && tied && keep[r]) lam_mine[r] = 0.f;
    }
    // slack behind the list: a valid row, so the last (partial) batch loads without bound checks
    if (lane < NG * UNR) mine[R + lane] = __ldg(pos);
    if (TRAIN) {
      if (SAMPLED) {
        sample_negs_chunk(wi, sa, m.nu, m.I, bt.row_ptr, bt.col, mine + n, lane);
      } else {
        const int32_t* neg = bt.negs + (int64_t)wi.aux0 * m.nu;
        for (int j = lane; j < R - n; j += 32) mine[n + j] = neg[j];
      }
    }
    __syncwarp();
  }
  // compile-time leading dimension: a row address is base + id * (4*LD) = one IMAD.WIDE
  const float* Wl = (m.asym ? m.V : m.W) + gl * 4;
  float* gWl = (m.asym ? m.gV : m.gW) + gl * 4;

  float4 z[NV], hg[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    z[v] = ld4(bt.Z + (int64_t)wi.u_local * LD + RM::col4(gl, v));
    hg[v] = f4zero();
  }
  // Only the LAST piece of a lane can lie entirely in the padding: DISPATCH_LD picks the smallest NV
  // with 4*G*NV >= K, so K > 4*G*(NV-1) and the pieces v < NV-1 hold real columns in every lane.
  const int last_live = RM::col4(gl, NV - 1) < m.K;
  float loss_acc = 0.f;
  int bad = 0;
  const int first = gl == 0;

  // The row loop is software-pipelined by hand: the loads of batch i + 1 (ids from shared memory, rows and
  // b' from L2) are issued BEFORE batch i is scored and reduced.  The reductions are `asm volatile` with a
  // memory clobber, which the compiler will not move loads across, so without this every batch waited out
  // its own L2 round trip (ncu r02_f: long_scoreboard 6.8 of ~12 stalled warps per issue).
  constexpr int STEP = NG * UNR;
  auto fetch = [&](int base, int (&id)[UNR], float4 (&w)[UNR][NV], float (&bp)[UNR]) {
#pragma unroll
    for (int t = 0; t < UNR; ++t) id[t] = mine[base + t * NG + grp];
#pragma unroll
    for (int t = 0; t < UNR; ++t) {
#pragma unroll
      for (int v = 0; v < NV; ++v) w[t][v] = ld4(Wl + (int64_t)id[t] * LD + v * G * 4);   // pad pieces read zeros
      bp[t] = __ldg(m.bp + id[t]);
    }
  };
  auto score = [&](int base, const int (&id)[UNR], const float4 (&w)[UNR][NV], const float (&bp)[UNR]) {
    float y[UNR];
#pragma unroll
    for (int t = 0; t < UNR; ++t) {
      float p = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) p += dot4(w[t][v], z[v]);
      y[t] = p;
    }
#pragma unroll
    for (int off_ = G / 2; off_ > 0; off_ >>= 1)
#pragma unroll
      for (int t = 0; t < UNR; ++t) y[t] += __shfl_xor_sync(0xffffffffu, y[t], off_);
#pragma unroll
    for (int t = 0; t < UNR; ++t) {
      if (base + t * NG >= R) break;          // warp-uniform: whole batches past the end of the chunk
      const int r = base + t * NG + grp;
      const float ok = r < R ? 1.f : 0.f;     // a row past the end inside the last batch adds exact zeros
      const float truth = r < n ? 1.f : 0.f;
      float l, g;
      if (LT == LOSS_CE) g = ce_grad(y[t] + bp[t], truth, &l);
      else g = loss_grad(LT >= 0 ? LT : m.loss, y[t] + bp[t], truth, &l, &bad);
      g *= ok;
      loss_acc = fmaf(ok, l, loss_acc);
      if (!TRAIN) continue;
      const float lam = lam_mine[r];          // 0 in the slack
      float* grow = gWl + (int64_t)id[t] * LD;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        hg[v] = fma4(g, w[t][v], hg[v]);
        const float4 gr = fma4(g, z[v], scale4(lam, w[t][v]));
        if (v < NV - 1) red_add_v4(grow + v * G * 4, gr);
        else red_add_v4_if(grow + v * G * 4, gr, last_live);
      }
      red_add_f32_if(m.gbp + id[t], ok * fmaf(lambda, bp[t], g), first);
    }
  };
  if constexpr (NST > 0) {
    float4 (*st)[UNR * NV][32] = stage_s[threadIdx.x >> 5];
    auto issue = [&](int base, int slot) {
#pragma unroll
      for (int t = 0; t < UNR; ++t) {
        const int idr = mine[base + t * NG + grp];
#pragma unroll
        for (int v = 0; v < NV; ++v) cp_async16(&st[slot][t * NV + v][lane], Wl + (int64_t)idr * LD + v * G * 4);
      }
    };
    const int n_b = (R + STEP - 1) / STEP;
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) {
      if (p < n_b) issue(p * STEP, p);
      cp_async_commit();
    }
    for (int b = 0; b < n_b; ++b) {
      if (b + NST - 1 < n_b) issue((b + NST - 1) * STEP, (b + NST - 1) % NST);
      cp_async_commit();
      int id[UNR];
      float bp[UNR];
#pragma unroll
      for (int t = 0; t < UNR; ++t) {
        id[t] = mine[b * STEP + t * NG + grp];
        bp[t] = __ldg(m.bp + id[t]);
      }
      cp_async_wait<NST - 1>();              // batch b has landed (this thread's own copies)
      float4 w[UNR][NV];
#pragma unroll
      for (int t = 0; t < UNR; ++t)
#pragma unroll
        for (int v = 0; v < NV; ++v) w[t][v] = st[b % NST][t * NV + v][lane];
      score(b * STEP, id, w, bp);
    }
  } else if constexpr (PIPE) {
    int idA[UNR], idB[UNR];
    float4 wA[UNR][NV], wB[UNR][NV];
    float bpA[UNR], bpB[UNR];
    fetch(0, idA, wA, bpA);
    for (int base = 0; base < R; base += 2 * STEP) {
      // (reads past R stay inside the list's slack only for ONE batch: guard the second prefetch)
      if (base + STEP < R) fetch(base + STEP, idB, wB, bpB);
      score(base, idA, wA, bpA);
      if (base + STEP >= R) break;
      if (base + 2 * STEP < R) fetch(base + 2 * STEP, idA, wA, bpA);
      score(base + STEP, idB, wB, bpB);
    }
  } else {
    for (int base = 0; base < R; base += STEP) {
      int id[UNR];
      float4 w[UNR][NV];
      float bp[UNR];
      fetch(base, id, w, bp);
      score(base, id, w, bp);
    }
  }
  if (TRAIN) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      hg[v] = cross_group_sum<G>(hg[v]);
      const int c = RM::col4(gl, v);
      if (grp == 0) red_add_v4(bt.HG + (int64_t)wi.u_local * LD + c, hg[v]);
    }
  }
  loss_acc = group_sum<32>(first ? loss_acc : 0.f);
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    atomicAdd(&stats->loss_sum[warp % STAT_STRIPES], (double)loss_acc);
    atomicAdd(&stats->outputs[warp % STAT_STRIPES], (unsigned long long)R);
    if (bad) stats->bad_loss = 1;
  }
}

// ---------------------------------------------------------------------------------------
// H8 user-local part, cdae.hpp:295-331: delta = hg (.) act'(z) (:208-215); the user's Wu row
// takes upd(delta + lambda*Wu[u]) in place (rows are user-private, so frozen-batch == online);
// gb += delta (the n*lambda*b term is added in apply_kernel).  Block = 32 users x ld/4 lanes.
__global__ void __launch_bounds__(256) hidden_backward_kernel(ModelDev m, BatchDev bt, StatsDev* stats) {
  extern __shared__ float4 red[];  // [blockDim.y][ld/4]
  const int ld4n = m.ld / 4;
  const int c4 = threadIdx.x;  // < ld4n
  const int u = blockIdx.x * blockDim.y + threadIdx.y;
  float4 d = f4zero();
  if (stats->bad_csr) return;  // invalid CSR (validate_rows_kernel): leave Wu untouched
  if (u < bt.n_users && c4 < ld4n) {
    const int c = c4 * 4;
    const float4 zz = ld4(bt.Z + (int64_t)u * m.ld + c);
    const float4 hg = ld4(bt.HG + (int64_t)u * m.ld + c);
    float4 dz = make_float4(1.f, 1.f, 1.f, 1.f);
    if (!m.linear) {
      if (!m.tanh_act) dz = make_float4(zz.x - zz.x * zz.x, zz.y - zz.y * zz.y, zz.z - zz.z * zz.z, zz.w - zz.w * zz.w);
      else dz = make_float4(1.f - zz.x * zz.x, 1.f - zz.y * zz.y, 1.f - zz.z * zz.z, 1.f - zz.w * zz.w);
    }
    d = mul4(hg, dz);
    // pad columns: hg is exactly 0 there (W' pad columns are 0), so d is 0.  (scatter_kernel forms the same
    // delta itself from HG and Z, so nothing is stored here.)
    if (m.user_factor) {
      const int64_t uid = bt.uids[u];
      float* wp = m.Wu + uid * m.ld + c;
      float* ap = m.Wu_ag + uid * m.ld + c;
      float4 wv = ld4(wp);
      float g[4] = {d.x + m.lambda * wv.x, d.y + m.lambda * wv.y, d.z + m.lambda * wv.z, d.w + m.lambda * wv.w};
      float wn[4] = {wv.x, wv.y, wv.z, wv.w};
      if (m.adagrad) {
        float4 av = ld4(ap);
        float a[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (g[i] == 0.f) continue;  // upd(g = 0) is a no-op; also keeps 0/0 out of pad columns
          a[i] += g[i] * g[i];
          g[i] = g[i] / (m.beta + sqrtf(a[i]));
        }
        st4(ap, make_float4(a[0], a[1], a[2], a[3]));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) wn[i] -= m.lr * g[i];
      st4(wp, make_float4(wn[0], wn[1], wn[2], wn[3]));
    }
  }
  // column sums of delta over the block's users -> one reduction per column per block
  red[threadIdx.y * blockDim.x + threadIdx.x] = d;
  __syncthreads();
  if (threadIdx.y == 0 && c4 < ld4n) {
    float4 s = f4zero();
    for (int y = 0; y < blockDim.y; ++y) s = add4(s, red[y * blockDim.x + threadIdx.x]);
    red_add_v4(m.gb + c4 * 4, s);
  }
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const int cnt = min((int)blockDim.y, bt.n_users - blockIdx.x * (int)blockDim.y);
    if (cnt > 0) {
      red_add_f32(m.g_steps + m.steps_slot, (float)cnt);
      atomicAdd(&stats->user_steps, (unsigned long long)cnt);
    }
  }
}

// H8 item part, cdae.hpp:333-349: every kept input row j gets
// gW[j] += scale*([Uu[u] (.)] delta) + lambda*W[j]; with linear_function also
// GU[u] += delta (.) W[j] (:340).  One warp per input chunk, same geometry as gather_kernel.
// The lambda*W[j] term needs no row read here: the kernel counts the occurrences of j
// (gcnt[j] += 1) and apply_kernel, which loads W[j] anyway, adds lambda*count*W[j] — that removes
// half of this kernel's L2 traffic (it is bound by L2 transactions: row reads + 16-byte reductions).
template <int G, int NV>
__global__ void __launch_bounds__(256) scatter_kernel(ModelDev m, BatchDev bt) {
  using RM = RowMap<G, NV>;
  constexpr int NG = RM::NG, UNR = RM::UNR;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= bt.n_in_items) return;
  const int lane = threadIdx.x & 31, grp = lane / G, gl = lane % G;
  const WorkItem wi = load_item(bt.in_items + warp);
  const int32_t* items = bt.col + wi.s0;
  const uint8_t* keep = bt.keep + wi.aux0;
  float* cnt = m.gcnt + (int64_t)m.steps_slot * m.I4;
  const bool count = m.lambda != 0.f && !m.direct_lambda;
  const bool direct = m.lambda != 0.f && m.direct_lambda;
  float4 d[NV], sd[NV], gu[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = RM::col4(gl, v);
    // delta = hg (.) act'(z) (cdae.hpp:208-215, 295-301), formed here from HG and Z rather than read back from
    // hidden_backward_kernel's D: the two kernels then have no dependency and run side by side (api.cu)
    {
      const float4 hgv = ld4(bt.HG + (int64_t)wi.u_local * m.ld + c);
      const float4 zz = ld4(bt.Z + (int64_t)wi.u_local * m.ld + c);
      float4 dz = make_float4(1.f, 1.f, 1.f, 1.f);
      if (!m.linear) {
        if (!m.tanh_act) dz = make_float4(zz.x - zz.x * zz.x, zz.y - zz.y * zz.y, zz.z - zz.z * zz.z, zz.w - zz.w * zz.w);
        else dz = make_float4(1.f - zz.x * zz.x, 1.f - zz.y * zz.y, 1.f - zz.z * zz.z, 1.f - zz.w * zz.w);
      }
      d[v] = mul4(hgv, dz);
    }
    sd[v] = scale4(m.scale, d[v]);
    if (m.linear_function) sd[v] = mul4(ld4(m.Uu + (int64_t)wi.uid * m.ld + c), sd[v]);
    gu[v] = f4zero();
  }
  for (int base = 0; base < wi.n; base += NG * UNR) {
    int it[UNR];
#pragma unroll
    for (int t = 0; t < UNR; ++t) {
      const int r = base + t * NG + grp;
      it[t] = (r < wi.n && keep[r]) ? __ldg(items + r) : -1;
    }
    if (m.linear_function) {
#pragma unroll
      for (int t = 0; t < UNR; ++t) {
        if (it[t] < 0) continue;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = RM::col4(gl, v);
          if (c < m.K) gu[v] = add4(gu[v], mul4(d[v], ld4(m.W + (int64_t)it[t] * m.ld + c)));
        }
      }
    }
#pragma unroll
    for (int t = 0; t < UNR; ++t) {
      if (it[t] < 0) continue;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = RM::col4(gl, v);
        if (c < m.K) {
          float4 add = sd[v];
          if (direct) add = fma4(m.lambda, ld4(m.W + (int64_t)it[t] * m.ld + c), add);
          red_add_v4(m.gW + (int64_t)it[t] * m.ld + c, add);
        }
      }
      if (count && gl == 0) red_add_f32(cnt + it[t], 1.f);
    }
  }
  if (m.linear_function) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      gu[v] = cross_group_sum<G>(gu[v]);
      const int c = RM::col4(gl, v);
      if (grp == 0) red_add_v4(bt.GU + (int64_t)wi.u_local * m.ld + c, gu[v]);
    }
  }
}

// cdae.hpp:295-299,351-357: Uu[u] takes upd(lambda*Uu[u] + GU[u]) (linear_function only).
__global__ void __launch_bounds__(256) uu_update_kernel(ModelDev m, BatchDev bt, const StatsDev* stats) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)bt.n_users * m.ld || stats->bad_csr) return;
  const int u = (int)(idx / m.ld), c = (int)(idx % m.ld);
  const int64_t uid = bt.uids[u];
  float* wp = m.Uu + uid * m.ld + c;
  float g = m.lambda * (*wp) + bt.GU[idx];
  if (c >= m.K) g = 0.f;
  if (g == 0.f) return;
  if (m.adagrad) {
    float* ap = m.Uu_ag + uid * m.ld + c;
    const float a = *ap + g * g;
    *ap = a;
    g = g / (m.beta + sqrtf(a));
  }
  *wp -= m.lr * g;
}

// ---------------------------------------------------------------------------------------
// H9, the upd() inlined at every update site of the reference (e.g. cdae.hpp:253-257):
// acc += g^2; g /= beta + sqrt(acc); w -= lr*g — applied ONCE per element with the summed
// minibatch gradient, then the accumulator is cleared.  Elements whose gradient is exactly 0
// are left untouched (upd with g = 0 is a no-op), so the dense pass equals "touched rows only".
struct ApplySeg {
  float *w, *acc, *g;
  int64_t n4;        // float4 count
  float extra_coef;  // b only: g += extra_coef * n_steps * w   (n * lambda * b)
  const float* cnt;  // W only: g[row] += cnt_coef * cnt[row] * w[row]   (lambda * occurrences * W[j])
  float cnt_coef;
  int ld4;           // float4 per row (cnt != nullptr)
};
struct ApplyArgs {
  ApplySeg seg[4];
  int nseg;
  float lr, beta;
  int adagrad;
  float* g_steps;  // [0..1]: read slot `steps_slot`, clear the other one for the next minibatch;
                   // [2] != 0: some rank's CSR failed validate_rows_kernel -> discard the gradients
  int steps_slot;
  float* cnt_clear;  // the occurrence counters of the OTHER slot (consumed by the previous apply)
  int64_t n_cnt;
  int* bad_csr_out;  // StatsDev::bad_csr: raised when the gradients were discarded (another rank's CSR was invalid)
};
__global__ void __launch_bounds__(256) apply_kernel(ApplyArgs a) {
  const float steps = a.g_steps[a.steps_slot];
  const bool discard = a.g_steps[2] != 0.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.g_steps[a.steps_slot ^ 1] = 0.f;
    if (discard) *a.bad_csr_out = 1;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_cnt / 4; i += (int64_t)gridDim.x * blockDim.x)
    st4(a.cnt_clear + i * 4, f4zero());
  // Two elements per thread and pass, and the weight / accumulator loads issued TOGETHER with the gradient load
  // (nearly every row of a minibatch is touched, so they are almost never wasted): one L2 round trip per pass
  // instead of "gradient, then — if non-zero — weight and accumulator".
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int s = 0; s < a.nseg; ++s) {
    const ApplySeg sg = a.seg[s];
    const float extra = sg.extra_coef * steps;
    auto finish = [&](int64_t i, float4 g4, float4 w4, float4 a4, float rowc) {
      if (extra == 0.f && rowc == 0.f && g4.x == 0.f && g4.y == 0.f && g4.z == 0.f && g4.w == 0.f) return;
      if (discard) {
        st4(sg.g + i * 4, f4zero());
        return;
      }
      float g[4] = {g4.x, g4.y, g4.z, g4.w};
      float w[4] = {w4.x, w4.y, w4.z, w4.w};
      const float lin = extra + rowc;
      if (lin != 0.f) {
#pragma unroll
        for (int k = 0; k < 4; ++k) g[k] += lin * w[k];
      }
      if (a.adagrad) {
        float ac[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (g[k] == 0.f) continue;  // no-op element (and no 0/0 in pad columns when beta = 0)
          ac[k] += g[k] * g[k];
          g[k] = g[k] / (a.beta + sqrtf(ac[k]));
        }
        st4(sg.acc + i * 4, make_float4(ac[0], ac[1], ac[2], ac[3]));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) w[k] -= a.lr * g[k];
      st4(sg.w + i * 4, make_float4(w[0], w[1], w[2], w[3]));
      st4(sg.g + i * 4, f4zero());
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sg.n4; i += 2 * stride) {
      const int64_t j = i + stride;
      const bool two = j < sg.n4;
      const float4 gi = ld4(sg.g + i * 4), wi = ld4(sg.w + i * 4);
      const float4 ai = a.adagrad ? ld4(sg.acc + i * 4) : f4zero();
      const float ci = sg.cnt ? sg.cnt_coef * __ldg(sg.cnt + i / sg.ld4) : 0.f;
      float4 gj = f4zero(), wj = f4zero(), aj = f4zero();
      float cj = 0.f;
      if (two) {
        gj = ld4(sg.g + j * 4);
        wj = ld4(sg.w + j * 4);
        if (a.adagrad) aj = ld4(sg.acc + j * 4);
        if (sg.cnt) cj = sg.cnt_coef * __ldg(sg.cnt + j / sg.ld4);
      }
      finish(i, gi, wi, ai, ci);
      if (two) finish(j, gj, wj, aj, cj);
    }
  }
}

// ---------------------------------------------------------------------------------------
// parameter plumbing
// CDAE::reset's Random(rows,K)*4*sqrt(6/(I+K)) (cdae.hpp:112-121) from Philox:
// element idx (row*K+k, unpadded) of block `which` = float((2u-1)*scale),
// u = (word+0.5)*2^-32, word = philox(seed, {idx lo, idx hi, which, 0xC0DE}).x
__global__ void init_uniform_kernel(float* dst, int64_t rows, int K, int ld, double scale,
                                    uint64_t seed, uint32_t which) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const int64_t r = i / K;
  const int k = (int)(i % K);
  const Philox4 p = philox4x32(seed, (uint32_t)i, (uint32_t)((uint64_t)i >> 32), which, 0xC0DEu);
  const double u = ((double)p.x + 0.5) * (1.0 / 4294967296.0);
  dst[r * ld + k] = (float)((2. * u - 1.) * scale);
}
__global__ void fill_kernel(float* dst, int64_t rows, int K, int ld, float v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  dst[i] = (int)(i % ld) < K ? v : 0.f;
}
// padded fp32 [rows][ld] <-> dense fp64 [rows][K]
__global__ void pack_from_double_kernel(float* dst, const double* src, int64_t rows, int K, int ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  const int64_t r = i / ld;
  const int k = (int)(i % ld);
  dst[i] = k < K ? (float)src[r * K + k] : 0.f;
}
__global__ void unpack_to_double_kernel(double* dst, const float* src, int64_t rows, int K, int ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const int64_t r = i / K;
  const int k = (int)(i % K);
  dst[i] = (double)src[r * ld + k];
}
__global__ void unpack_to_float_kernel(float* dst, const float* src, int64_t rows, int K, int ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const int64_t r = i / K;
  const int k = (int)(i % K);
  dst[i] = src[r * ld + k];
}
// Process-group read-back of user-private tables: copy the rows this rank trains (its slice of
// each global minibatch of B users), zero the others; a sum across ranks then yields the table.
__global__ void owned_rows_kernel(float* dst, const float* src, int64_t rows, int ld, int64_t B,
                                  int64_t U, int rank, int world) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  const int64_t uid = i / ld;
  const int64_t lo = (uid / B) * B;
  const int64_t n = min(B, U - lo);
  const int64_t a = lo + (n * rank) / world, b = lo + (n * (rank + 1)) / world;
  dst[i] = (uid >= a && uid < b) ? src[i] : 0.f;
}

// the part of an item-side block that lies inside this rank's slice [lo, hi) of the flat buffer
__global__ void owned_flat_kernel(float* dst, const float* src, int64_t n, int64_t block_off, int64_t lo, int64_t hi) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t f = block_off + i;
  dst[i] = (f >= lo && f < hi) ? src[i] : 0.f;
}
__global__ void add_double_kernel(double* dst, const double* src) { *dst += *src; }

// sum of squares of the user rows this rank trains (ownership rule of build_plan)
__global__ void __launch_bounds__(256) sumsq_owned_rows_kernel(const float* src, int64_t U, int ld, int64_t B, int rank,
                                                               int world, double* out) {
  double s = 0.;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < U * ld; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t uid = i / ld;
    const int64_t lo = (uid / B) * B;
    const int64_t nb = min(B, U - lo);
    const int64_t a = lo + (nb * rank) / world, b = lo + (nb * (rank + 1)) / world;
    if (uid >= a && uid < b) {
      const double v = src[i];
      s += v * v;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    atomicAdd(out, t);
  }
}

// sum of squares, double accumulation (CDAE::penalty_loss, cdae.hpp:103-107 / penalty.hpp:36-39)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* src, int64_t n, double* out) {
  double s = 0.;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double v = src[i];
    s += v * v;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    atomicAdd(out, t);
  }
}

}  // namespace cdae
