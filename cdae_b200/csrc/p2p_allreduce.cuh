// cdae_b200/csrc/p2p_allreduce.cuh — the per-minibatch gradient all-reduce over NVLink peer memory
// (SURVEY.md §8e), written for this path instead of calling NCCL: every rank's gradient buffer and a
// small flag array are mapped into every other rank (CUDA IPC), and one kernel per rank
//   1. tells every peer "my gradients are complete" and waits until all peers said so,
//   2. sums ITS 1/G slice of the buffer over all ranks with 16-byte loads straight from peer memory,
//   3. writes the sum back into that slice of EVERY rank's buffer (two-shot all-reduce: reduce-scatter
//      by peer loads, all-gather by peer stores; each byte crosses NVLink once in each direction),
// followed by a one-block barrier kernel ("my stores are out" / "everybody's stores are in") before
// apply_kernel reads the buffer.  At config B the buffer is 13.6 MB; NCCL's ring all-reduce of that
// size is latency-bound at ~0.1 ms on 8 GPUs (32 % of an epoch's device time, profiles/r01_n_*).
// Opt-in (cdae_dist_p2p_export / cdae_dist_p2p_open); NCCL stays the default and the fallback.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cdae {
namespace p2p {

constexpr int MAX_RANKS = 8;

struct Args {
  float* bufs[MAX_RANKS];          // every rank's gradient buffer (own one included), same layout
  uint32_t* flags[MAX_RANKS];      // every rank's flag array: flags[p][q] = last epoch rank q announced to rank p
  int rank, world;
  uint32_t epoch;                  // value to announce / wait for (monotonic, two per minibatch)
  int64_t n4;                      // float4 count of the buffer
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_v4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// announce `epoch` to every rank (one thread per peer) — call from ONE block, after a system fence
__device__ __forceinline__ void announce(const Args& a) {
  if ((int)threadIdx.x < a.world) st_release_sys(a.flags[threadIdx.x] + a.rank, a.epoch);
}
// every calling block waits until all ranks announced `epoch` to this rank
__device__ __forceinline__ void wait_all(const Args& a) {
  if ((int)threadIdx.x < a.world)
    while ((int32_t)(ld_acquire_sys(a.flags[a.rank] + threadIdx.x) - a.epoch) < 0) {}
  __syncthreads();
}

__global__ void __launch_bounds__(512) reduce_kernel(Args a) {
  // stream order: every kernel that added to this rank's gradients has finished
  if (blockIdx.x == 0) {
    __threadfence_system();
    announce(a);
  }
  wait_all(a);
  const int64_t lo = a.n4 * a.rank / a.world, hi = a.n4 * (a.rank + 1) / a.world;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < MAX_RANKS; ++p)
      if (p < a.world) {                       // fixed order: every rank would compute the same sum, bit for bit
        const float4 v = ld_sys_v4(a.bufs[p] + i * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
#pragma unroll
    for (int p = 0; p < MAX_RANKS; ++p)
      if (p < a.world) st_sys_v4(a.bufs[p] + i * 4, s);
  }
}

// "my stores are out" -> "everybody's stores are in"; a.epoch is the SECOND value of the minibatch
__global__ void __launch_bounds__(32) barrier_kernel(Args a) {
  __threadfence_system();
  announce(a);
  wait_all(a);
}

}  // namespace p2p
}  // namespace cdae
