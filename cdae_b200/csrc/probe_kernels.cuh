// cdae_b200/csrc/probe_kernels.cuh — measurement aid, not part of the training path: the L2 roofline
// of the row gather / row reduction access pattern the sampled decode (decode_kernel) is made of.
//
// At config B the item tables and their gradient buffer (38 MB) live in the 126 MB L2, so the
// kernel's ceiling is not HBM bandwidth but how fast the L2 serves 16-byte vector loads of random
// table rows and absorbs 16-byte vector reductions (red.global.add.v4.f32) into random rows of a
// second table.  This kernel issues exactly those transactions — same <G,NV> lane geometry, same
// UNR row batches in flight per warp, uniformly random rows, no arithmetic beyond keeping the loads
// alive — and bench.py reports its bytes/s as `peak_l2`, the denominator of the decode roofline.
#pragma once
#include "train_kernels.cuh"

namespace cdae {

enum { PROBE_READ = 1, PROBE_RED = 2 };

__device__ __forceinline__ uint32_t probe_hash(uint32_t x) {  // lowbias32
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// One warp handles `rows_per_warp` random rows: MODE & PROBE_READ loads the row of `src`,
// MODE & PROBE_RED reduces a vector into the same row of `dst`.
template <int G, int NV, int MODE>
__global__ void __launch_bounds__(256) l2_probe_kernel(const float* __restrict__ src, float* dst, int64_t rows,
                                                       int ld, int rows_per_warp, int n_warps, uint32_t salt,
                                                       float* sink) {
  using RM = RowMap<G, NV>;
  constexpr int NG = RM::NG, UNR = RM::UNR;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= n_warps) return;
  const int lane = threadIdx.x & 31, grp = lane / G, gl = lane % G;
  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = make_float4(1e-9f, 2e-9f, 3e-9f, 4e-9f);
  for (int base = 0; base < rows_per_warp; base += NG * UNR) {
    int64_t it[UNR];
#pragma unroll
    for (int t = 0; t < UNR; ++t) {
      const uint32_t r = (uint32_t)(base + t * NG + grp);
      it[t] = (int64_t)(((uint64_t)probe_hash((uint32_t)warp * 0x9E3779B9u + r + salt) * (uint64_t)rows) >> 32);
    }
    float4 w[UNR][NV];
    if (MODE & PROBE_READ) {
#pragma unroll
      for (int t = 0; t < UNR; ++t)
#pragma unroll
        for (int v = 0; v < NV; ++v) w[t][v] = ld4(src + it[t] * ld + RM::col4(gl, v));
#pragma unroll
      for (int t = 0; t < UNR; ++t)
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = add4(acc[v], w[t][v]);
    }
    if (MODE & PROBE_RED) {
#pragma unroll
      for (int t = 0; t < UNR; ++t)
#pragma unroll
        for (int v = 0; v < NV; ++v) red_add_v4(dst + it[t] * ld + RM::col4(gl, v), acc[v]);
    }
  }
  if (MODE == PROBE_READ) {  // keep the loads alive
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) s += acc[v].x + acc[v].y + acc[v].z + acc[v].w;
    if (s == 123.456f) sink[0] = s;
  }
}

// Same reductions issued as bulk asynchronous copies with add (cp.reduce.async.bulk: the TMA unit reads
// a whole row from shared memory and sends it to L2 as vector reductions) instead of one 16-byte RED per
// lane: `bytes` per row (a multiple of 16), NG rows per warp step, NBUF staging buffers in flight.
__device__ __forceinline__ void bulk_red_add_f32(float* gdst, const float* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int G, int NV, bool READ>
__global__ void __launch_bounds__(256) l2_probe_bulk_kernel(const float* __restrict__ src, float* dst, int64_t rows,
                                                            int rows_per_warp, int n_warps, uint32_t salt,
                                                            uint32_t bytes, float* sink) {
  using RM = RowMap<G, NV>;
  constexpr int NG = RM::NG, LD = 4 * G * NV, NBUF = NV <= 2 ? 4 : 2;   // static shared memory: 8 warps x NBUF x 512*NV bytes
  __shared__ __align__(128) float stage[8][NBUF][NG][LD];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= n_warps) return;
  const int lane = threadIdx.x & 31, grp = lane / G, gl = lane % G;
  float (*mine)[NG][LD] = stage[threadIdx.x >> 5];
  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = make_float4(1e-9f, 2e-9f, 3e-9f, 4e-9f);
  int buf = 0;
  for (int base = 0; base < rows_per_warp; base += NG) {
    const uint32_t r = (uint32_t)(base + grp);
    const int64_t it = (int64_t)(((uint64_t)probe_hash((uint32_t)warp * 0x9E3779B9u + r + salt) * (uint64_t)rows) >> 32);
    if (READ) {
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] = add4(acc[v], ld4(src + it * LD + RM::col4(gl, v)));
    }
    bulk_wait_read<NBUF - 1>();           // the buffer about to be overwritten has been read by its bulk op
    __syncwarp();
#pragma unroll
    for (int v = 0; v < NV; ++v) st4(&mine[buf][grp][RM::col4(gl, v)], acc[v]);
    fence_proxy_async_smem();
    __syncwarp();
    if (gl == 0) bulk_red_add_f32(dst + it * LD, &mine[buf][grp][0], bytes);
    bulk_commit();
    buf = (buf + 1) % NBUF;
  }
  bulk_wait_read<0>();
  if (READ) {
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) s += acc[v].x + acc[v].y + acc[v].z + acc[v].w;
    if (s == 123.456f) sink[0] = s;
  }
}

}  // namespace cdae
