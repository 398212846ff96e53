// cdae_b200/csrc/topn_kernels.cuh — full-item decode for CDAE::recommend (cdae.hpp:162-196).
//
// Scores S[u][i] = W'[i].z_u + b'[i] for EVERY user against EVERY item is a dense
// (U x K)·(K x I) contraction.  It is never materialised: a CTA owns a tile of users, walks
// all item tiles, and an epilogue thread per user keeps that user's best M candidates in
// shared memory (a candidate is inserted only when it beats the current M-th best, so the
// common case is one compare per score).  Candidates are then re-scored in fp64 and the
// final top-k is taken with the reference's rule (heap.hpp:44-52: strict improvement =>
// on equal scores the lower item id stays; result sorted by score descending).
//
// This file holds the exact fp32 CUDA-core tile kernel (also the fallback of the tensor-core
// path), the fp64 re-rank, and the on-device TOPN_Evaluation metrics (evaluation.hpp:183-219).
#pragma once
#include "common.cuh"
#include "train_kernels.cuh"  // row_contains

namespace cdae {

constexpr int TOPN_M = 64;       // candidates kept per user before the exact re-rank
constexpr int TOPN_MAX_K = 32;   // largest supported topk

// Per-user candidate list in shared memory, owned by ONE thread.
struct CandList {
  float* s;   // [M] scores
  int* id;    // [M] item ids
  int cnt;
  float thr;  // current minimum once the list is full, -inf before
  __device__ __forceinline__ void offer(float score, int item, const int32_t* rated, int n_rated) {
    if (cnt == TOPN_M && !(score > thr)) return;
    if (row_contains(rated, n_rated, item)) return;  // cdae.hpp:177-179
    if (cnt < TOPN_M) {
      s[cnt] = score;
      id[cnt] = item;
      ++cnt;
      if (cnt < TOPN_M) return;
    } else {
      int mn = 0;
      for (int t = 1; t < TOPN_M; ++t)
        if (s[t] < s[mn]) mn = t;
      s[mn] = score;
      id[mn] = item;
    }
    float m = s[0];
    for (int t = 1; t < TOPN_M; ++t) m = fminf(m, s[t]);
    thr = m;
  }
};

// Exact fp32 path: 64 users x 64 items per tile on the CUDA cores, K in chunks of 32.
constexpr int TT_U = 64, TT_I = 64, TT_K = 32;
__global__ void __launch_bounds__(256) topn_tile_fp32_kernel(
    const float* __restrict__ Z, const float* __restrict__ Wd, const float* __restrict__ bp,
    int64_t I, int ld, const int32_t* __restrict__ users, int n_users,
    const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
    int* __restrict__ cand_id, float* __restrict__ cand_s, int* __restrict__ cand_cnt) {
  __shared__ float Zs[TT_U][TT_K + 1];
  __shared__ float Ws[TT_I][TT_K + 1];
  __shared__ float S[TT_U][TT_I + 1];
  extern __shared__ unsigned char dyn[];  // lists: [TT_U][M] float + [TT_U][M] int
  float* ls = reinterpret_cast<float*>(dyn);
  int* li = reinterpret_cast<int*>(ls + TT_U * TOPN_M);

  const int tid = threadIdx.x;
  const int u0 = blockIdx.x * TT_U;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads, 4 x 4 outputs each

  CandList cl;
  int my_uid = -1, n_rated = 0;
  const int32_t* rated = nullptr;
  if (tid < TT_U) {
    cl.s = ls + tid * TOPN_M;
    cl.id = li + tid * TOPN_M;
    cl.cnt = 0;
    cl.thr = -INFINITY;
    if (u0 + tid < n_users) {
      my_uid = users ? users[u0 + tid] : u0 + tid;
      const int64_t r0 = row_ptr[my_uid];
      rated = col + r0;
      n_rated = (int)(row_ptr[my_uid + 1] - r0);
    }
  }

  for (int64_t i0 = 0; i0 < I; i0 += TT_I) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int k0 = 0; k0 < ld; k0 += TT_K) {
      for (int e = tid; e < TT_U * TT_K; e += 256) {
        const int r = e / TT_K, c = e % TT_K;
        const int u = u0 + r;
        float v = 0.f;
        if (u < n_users && k0 + c < ld) {
          const int64_t uid = users ? users[u] : u;
          v = Z[uid * ld + k0 + c];
        }
        Zs[r][c] = v;
      }
      for (int e = tid; e < TT_I * TT_K; e += 256) {
        const int r = e / TT_K, c = e % TT_K;
        const int64_t it = i0 + r;
        Ws[r][c] = (it < I && k0 + c < ld) ? Wd[it * ld + k0 + c] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < TT_K; ++k) {
        float zr[4], wr[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) zr[a] = Zs[ty * 4 + a][k];
#pragma unroll
        for (int b = 0; b < 4; ++b) wr[b] = Ws[tx * 4 + b][k];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(zr[a], wr[b], acc[a][b]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int64_t it = i0 + tx * 4 + b;
        S[ty * 4 + a][tx * 4 + b] = acc[a][b] + (it < I ? bp[it] : 0.f);
      }
    __syncthreads();
    if (tid < TT_U && my_uid >= 0) {
      const int lim = (int)min((int64_t)TT_I, I - i0);
      for (int j = 0; j < lim; ++j) cl.offer(S[tid][j], (int)(i0 + j), rated, n_rated);
    }
    __syncthreads();
  }
  if (tid < TT_U && my_uid >= 0) {
    const int64_t o = (int64_t)(u0 + tid) * TOPN_M;
    for (int t = 0; t < cl.cnt; ++t) {
      cand_id[o + t] = cl.id[t];
      cand_s[o + t] = cl.s[t];
    }
    cand_cnt[u0 + tid] = cl.cnt;
  }
}

// Exact re-rank: fp64 score of each candidate, final top-k by (score desc, id asc).
// One warp per user; `stride` = candidate slots per user in cand_id (<= RERANK_MAX).
// Returns -1 ids when fewer than topk unrated items exist (the reference CHECK-aborts there,
// cdae.hpp:187; the host turns it into an error).
//
// VERIFY mode (thr != nullptr, candidates from the bf16 tensor-core kernel): every unrated item
// that is not a candidate has APPROXIMATE score <= thr[u] and |approx - exact| <= eps[u], so the
// list is provably the exact one iff  thr[u] + eps[u] < (k-th best exact candidate score).
// Users that fail (or have fewer than topk candidates) are appended to redo_list and left
// untouched; the exact fp32 kernel then handles them.
constexpr int RERANK_MAX = 96;
__global__ void __launch_bounds__(256) topn_rerank_kernel(
    const float* __restrict__ Z, const float* __restrict__ Wd, const float* __restrict__ bp, int K,
    int ld, const int32_t* __restrict__ users, int n_users, const int* __restrict__ cand_id,
    const int* __restrict__ cand_cnt, int stride, int n_splits, int topk, int32_t* __restrict__ out_id,
    float* __restrict__ out_s, int* __restrict__ short_flag, const float* __restrict__ thr,
    const float* __restrict__ eps, int32_t* __restrict__ redo_list, int* __restrict__ redo_cnt,
    float* __restrict__ redo_thr) {
  __shared__ double sc[8][RERANK_MAX];
  __shared__ int ids[8][RERANK_MAX];
  __shared__ double kth[8];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * 8 + w;
  if (u >= n_users) return;
  const int64_t uid = users ? users[u] : u;
  const float* z = Z + uid * ld;
  // candidates arrive in n_splits segments of `stride` slots (one per item range of the
  // tensor-core kernel; a single segment otherwise); gather their ids, then score them exactly:
  // 8 lanes per candidate read its row with 16-byte loads (4 candidates per warp step), partial
  // dot products in fp64, 3 shuffles.  The summation order differs from a serial loop only in
  // fp64 rounding (~1e-16 relative), far below any score gap fp32 parameters can form.
  int cnt = 0;
  float thr_u = -INFINITY;
  for (int sgm = 0; sgm < n_splits; ++sgm) {
    const int64_t slot = (int64_t)u * n_splits + sgm;
    const int n = min(cand_cnt[slot], RERANK_MAX - cnt);
    for (int c = lane; c < n; c += 32) ids[w][cnt + c] = cand_id[slot * stride + c];
    cnt += n;
    if (thr) thr_u = fmaxf(thr_u, thr[slot]);
  }
  __syncwarp();
  {
    const int grp = lane >> 3, gl = lane & 7;
    for (int c0 = 0; c0 < cnt; c0 += 4) {
      const int c = c0 + grp;
      const int it = c < cnt ? ids[w][c] : -1;
      double s = 0.;
      if (it >= 0) {
        const float* wr = Wd + (int64_t)it * ld;
        for (int k = gl * 4; k < ld; k += 32) {   // pad columns of both rows are exactly 0
          const float4 a = *reinterpret_cast<const float4*>(wr + k);
          const float4 b = *reinterpret_cast<const float4*>(z + k);
          s += (double)a.x * (double)b.x + (double)a.y * (double)b.y + (double)a.z * (double)b.z +
               (double)a.w * (double)b.w;
        }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (gl == 0 && it >= 0) sc[w][c] = s + (double)bp[it];
    }
  }
  __syncwarp();
  if (cnt < topk) {
    if (thr) {
      if (lane == 0) {
        const int slot = atomicAdd(redo_cnt, 1);
        redo_list[slot] = (int32_t)uid;
        if (redo_thr) redo_thr[slot] = -INFINITY;
      }
      return;
    }
    if (lane == 0) *short_flag = 1;
    for (int t = lane; t < topk; t += 32) {
      out_id[uid * topk + t] = -1;
      out_s[uid * topk + t] = 0.f;
    }
    return;
  }
  int my_rank[(RERANK_MAX + 31) / 32];
#pragma unroll
  for (int i = 0; i < (RERANK_MAX + 31) / 32; ++i) {
    const int c = lane + i * 32;
    my_rank[i] = 1 << 30;
    if (c < cnt) {
      const double s = sc[w][c];
      const int id = ids[w][c];
      int rank = 0;
      for (int o = 0; o < cnt; ++o) {
        const double so = sc[w][o];
        rank += (so > s) || (so == s && ids[w][o] < id);
      }
      my_rank[i] = rank;
      if (rank == topk - 1) kth[w] = s;
    }
  }
  __syncwarp();
  if (thr) {
    const bool ok = (double)thr_u + (double)eps[u] < kth[w];
    if (!ok) {
      if (lane == 0) {
        const int slot = atomicAdd(redo_cnt, 1);
        redo_list[slot] = (int32_t)uid;
        // An item of the true top-k has exact score >= kth (k candidates already reach it), hence
        // approximate score >= kth - eps: a second sweep may START at a threshold just below that
        // and then collects every possible member (rounded down, with a relative margin).
        if (redo_thr) {
          const double t0 = kth[w] - (double)eps[u];
          redo_thr[slot] = __double2float_rd(t0 - 1e-6 * fabs(t0) - 1e-30);
        }
      }
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < (RERANK_MAX + 31) / 32; ++i) {
    const int c = lane + i * 32;
    if (c < cnt && my_rank[i] < topk) {
      out_id[uid * topk + my_rank[i]] = ids[w][c];
      out_s[uid * topk + my_rank[i]] = (float)sc[w][c];
    }
  }
}

// TOPN_Evaluation::evaluate_rec_list (evaluation.hpp:183-219) per user with test items, summed.
__global__ void __launch_bounds__(256) topn_metrics_kernel(const int32_t* __restrict__ rec, int topk,
                                                           int64_t U, const int64_t* __restrict__ trp,
                                                           const int32_t* __restrict__ tcol,
                                                           double* __restrict__ out9) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double cnt = 0.;
  if (u < U) {
    const int64_t t0 = trp[u];
    const int nt = (int)(trp[u + 1] - t0);
    if (nt > 0) {
      cnt = 1.;
      double hit = 0., map5 = 0., map10 = 0.;
      const int lim = min(20, topk);
      for (int idx = 0; idx < lim; ++idx) {
        if (row_contains(tcol + t0, nt, rec[u * topk + idx])) {
          hit += 1.;
          if (idx < 5) map5 += hit / (idx + 1);
          if (idx < 10) map10 += hit / (idx + 1);
        }
        if (idx == 0) { r[0] = hit; r[3] = hit / nt; }
        else if (idx == 4) { r[1] = hit / 5.; r[4] = hit / nt; }
        else if (idx == 9) { r[2] = hit / 10.; r[5] = hit / nt; }
      }
      r[6] = map5 / (double)min(5, nt);
      r[7] = map10 / (double)min(10, nt);
    }
  }
  // block reduction, then one atomic per block per metric
  __shared__ double sm[8][9];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double v[9] = {r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], cnt};
#pragma unroll
  for (int k = 0; k < 9; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
    if (lane == 0) sm[w][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double t = 0.;
    for (int ww = 0; ww < 8; ++ww) t += sm[ww][threadIdx.x];
    atomicAdd(out9 + threadIdx.x, t);
  }
}

}  // namespace cdae
