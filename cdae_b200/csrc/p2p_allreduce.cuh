// cdae_b200/csrc/p2p_allreduce.cuh — the per-minibatch combine step of the data-parallel path
//  as ONE kernel over NVLink peer memory instead of "NCCL all-reduce, then the same
// dense optimiser pass on every rank":
//
//     reduce-scatter (peer loads)  ->  optimiser step on the rank's 1/G slice  ->  all-gather (peer stores)
//
// Every rank's gradient buffer, item-side parameter buffer and a small flag array are mapped into
// every other rank (CUDA IPC).  The three item-side buffers share one layout (point_item_side in
// api.cu), so the whole step is a flat elementwise pass.  One launch per rank and minibatch:
//   1. block 0 tells every peer "my gradients are complete" (system-scope release store into the
//      peer's flags); every block waits until all peers said so;
//   2. the rank sums ITS 1/G slice of the gradient buffer over all ranks with 16-byte loads straight
//      from peer memory, in rank order (the sum is a pure function of the G buffers);
//   3. it applies upd() (AdaGrad or SGD, cdae.hpp:253-257) to that slice with ITS accumulator slice —
//      accumulators are sharded: only the owner of a slice keeps them current — and
//   4. stores the updated parameters into that slice of EVERY rank's parameter buffer (elements whose
//      gradient is exactly zero are skipped: nothing changed, nothing crosses the wire);
//   5. meanwhile it zeroes the OTHER gradient buffer (gradients ping-pong between two buffers, so the
//      one the previous minibatch consumed is cleared here, off the critical path);
//   6. the last block to finish announces "my stores are out" and waits for the same from every peer —
//      when the kernel ends, this rank's parameters are complete and its gradients have been read.
// Each gradient byte crosses NVLink once (as a load), each parameter byte once (as a store): the
// traffic of a two-shot all-reduce, with 1/G of the optimiser work per rank and no second pass
// over the buffer.  At config B the buffer is 13.6 MB (NCCL's all-reduce of it: ~100 us on 8 GPUs,
// latency-bound); at config D it is 102 MB and the dense apply it replaces streams 614 MB per rank.
//
// `reduce_kernel` / `barrier_kernel` (plain two-shot all-reduce followed by the replicated
// apply_kernel) are kept as the A/B baseline (CDAE_B200_P2P_FUSED=0).
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
__device__ __forceinline__ float ld_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
// Bulk data of the fused step: weak accesses.  The peers' gradients are complete and unchanging once their
// flag has been acquired (acquire + __syncthreads orders these loads after it; the L1 was invalidated at
// kernel launch and every address is read once, so no stale line can be hit), and the stores are published
// by the system-scope fence in front of the closing announcement.  STRONG.SYS accesses reached only
// ~250 GB/s per direction on 8 GPUs (profiles/r02_c_*).
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f32(const float* p) {
  float v;
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer_v4(float* p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_sys_v4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// announce `epoch` to every rank (one thread per peer) — call from ONE block, after a system fence
__device__ __forceinline__ void announce(uint32_t* const* flags, int rank, int world, uint32_t epoch) {
  if ((int)threadIdx.x < world) st_release_sys(flags[threadIdx.x] + rank, epoch);
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// A peer that never arrives (its process failed before the combine step) must not hang this GPU: the spin
// gives up after SPIN_LIMIT_NS and raises *timeout_flag = 2 (reported by the host as CDAE_E_STATE; the
// parameters of this call are then undefined).
constexpr uint64_t SPIN_LIMIT_NS = 20ull * 1000 * 1000 * 1000;
// every calling block waits until all ranks announced `epoch` to this rank
__device__ __forceinline__ void wait_all(uint32_t* const* flags, int rank, int world, uint32_t epoch,
                                         int* timeout_flag = nullptr) {
  if ((int)threadIdx.x < world) {
    const uint64_t t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(flags[rank] + threadIdx.x) - epoch) < 0) {
      if (global_ns() - t0 > SPIN_LIMIT_NS) {
        if (timeout_flag) *timeout_flag = 2;
        break;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512) reduce_kernel(Args a) {
  // stream order: every kernel that added to this rank's gradients has finished
  if (blockIdx.x == 0) {
    __threadfence_system();
    announce(a.flags, a.rank, a.world, a.epoch);
  }
  wait_all(a.flags, a.rank, a.world, a.epoch);
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
  announce(a.flags, a.rank, a.world, a.epoch);
  wait_all(a.flags, a.rank, a.world, a.epoch);
}

// ---------------------------------------------------------------------------------------
// The fused combine step (see the file comment).
struct FusedArgs {
  float* grads[MAX_RANKS];         // every rank's CURRENT gradient buffer
  float* params[MAX_RANKS];        // every rank's item-side parameter buffer
  uint32_t* flags[MAX_RANKS];
  float* acc;                      // this rank's accumulator buffer (only [lo, hi) is kept current)
  float* grad_next;                // this rank's OTHER gradient buffer: zeroed here for the next minibatch
  unsigned int* done;              // this rank's block counter (self-resetting)
  int* bad_csr_out;                // StatsDev::bad_csr
  int rank, world;
  uint32_t epoch;                  // announces epoch (gradients complete) and epoch + 1 (stores out)
  int64_t n4;                      // float4 count of one buffer
  // layout, in float4 units: [0, w_end) = W and V rows of ld4 float4; [w_end, bp_end) = b';
  // [bp_end, b_lo) = kept-input counts (two slots of cnt4 float4, no parameters); [b_lo, b_hi) = b; then steps
  int64_t w_rows_end;              // I*ld4 (only W rows carry the lambda*count*W term)
  int64_t w_end, bp_end, b_lo, b_hi;
  int ld4;
  int64_t cnt_off;                 // float offset of the active counts slot
  int64_t steps_off;               // float offset of steps[0..3] ([slot] user steps, [2] bad-CSR flag)
  int steps_slot;
  float lr, beta, lambda;
  int adagrad;
  unsigned long long* ts;          // nullable: %globaltimer at the phase boundaries of block 0 / the last block (cdae_debug_combine)
};
#define P2P_STAMP(slot) do { if (a.ts && threadIdx.x == 0) a.ts[slot] = global_ns(); } while (0)

__global__ void __launch_bounds__(512) fused_step_kernel(FusedArgs a) {
  if (blockIdx.x == 0) {
    P2P_STAMP(0);
    __threadfence_system();
    announce(a.flags, a.rank, a.world, a.epoch);
  }
  wait_all(a.flags, a.rank, a.world, a.epoch, a.bad_csr_out);
  if (blockIdx.x == 0) P2P_STAMP(1);

  // scalars every element may need: user steps of the minibatch (n * lambda * b) and the bad-CSR flag.  ONE
  // thread per peer and block fetches them (every thread doing so made 10^5 - 10^6 requests for the same two
  // words of each peer: the whole step ran at 40 GB/s on 8 GPUs, profiles/r02_c_*).
  __shared__ float sc_s[2][MAX_RANKS];
  if ((int)threadIdx.x < a.world) {
    sc_s[0][threadIdx.x] = ld_sys_f32(a.grads[threadIdx.x] + a.steps_off + a.steps_slot);
    sc_s[1][threadIdx.x] = ld_sys_f32(a.grads[threadIdx.x] + a.steps_off + 2);
  }
  __syncthreads();
  float steps = 0.f, bad = 0.f;
#pragma unroll
  for (int p = 0; p < MAX_RANKS; ++p)
    if (p < a.world) {
      steps += sc_s[0][p];
      bad += sc_s[1][p];
    }
  const bool discard = bad != 0.f;
  if (discard && blockIdx.x == 0 && threadIdx.x == 0) *a.bad_csr_out = 1;

  const int64_t lo = a.n4 * a.rank / a.world, hi = a.n4 * (a.rank + 1) / a.world;
  float* const my_params = a.params[a.rank];
  // UN elements per thread and iteration, all peer loads issued before the first use: an NVLink round trip
  // is ~2-3 us, so the loop must keep several 16-byte loads per peer in flight per thread
  constexpr int UN = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool first_pass = true;
  for (int64_t i0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 - (int64_t)threadIdx.x - (int64_t)blockIdx.x * blockDim.x < hi; i0 += stride * UN) {
    float4 s[UN];
    bool on[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t i = i0 + u * stride;
      // counts and steps have no parameter behind them
      on[u] = i < hi && !(i >= a.bp_end && i < a.b_lo) && i < a.b_hi;
      s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!on[u]) continue;
#pragma unroll
      for (int p = 0; p < MAX_RANKS; ++p)
        if (p < a.world) {                       // fixed rank order: the sum is a pure function of the G buffers
          const float4 v = ld_peer_v4(a.grads[p] + i * 4);
          s[u].x += v.x; s[u].y += v.y; s[u].z += v.z; s[u].w += v.w;
        }
    }
    if (first_pass && a.grad_next) {
      // zero the gradient buffer the PREVIOUS minibatch consumed while the peer loads above are in flight
      first_pass = false;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += stride)
        *reinterpret_cast<float4*>(a.grad_next + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (!on[u] || discard) continue;
      const int64_t i = i0 + u * stride;
      // coefficient of w in the gradient: lambda * user steps for b (item rows: scatter_kernel added lambda*W[j] itself)
      const float lin = i >= a.b_lo ? a.lambda * steps : 0.f;
      const float4 sv = s[u];
      if (lin == 0.f && sv.x == 0.f && sv.y == 0.f && sv.z == 0.f && sv.w == 0.f) continue;   // untouched: identical everywhere already
      const float4 w4 = *reinterpret_cast<const float4*>(my_params + i * 4);
      float g[4] = {sv.x, sv.y, sv.z, sv.w};
      float w[4] = {w4.x, w4.y, w4.z, w4.w};
      if (lin != 0.f) {
#pragma unroll
        for (int k = 0; k < 4; ++k) g[k] += lin * w[k];
      }
      if (a.adagrad) {
        const float4 a4 = *reinterpret_cast<const float4*>(a.acc + i * 4);
        float ac[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (g[k] == 0.f) continue;                  // upd(0) is a no-op (and no 0/0 in pad columns when beta = 0)
          ac[k] += g[k] * g[k];
          g[k] = g[k] / (a.beta + sqrtf(ac[k]));
        }
        *reinterpret_cast<float4*>(a.acc + i * 4) = make_float4(ac[0], ac[1], ac[2], ac[3]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) w[k] -= a.lr * g[k];
      const float4 wn = make_float4(w[0], w[1], w[2], w[3]);
#pragma unroll
      for (int p = 0; p < MAX_RANKS; ++p)
        if (p < a.world) st_peer_v4(a.params[p] + i * 4, wn);
    }
  }
  if (blockIdx.x == 0) P2P_STAMP(2);

  // last block out: "my stores are out" -> wait until everybody's are in.  One system-scope fence per block,
  // after the block barrier, orders every thread's stores (fence cumulativity — the grid.sync() pattern)
  // before the counter increment; 75,000 per-thread MEMBAR.SYS would each wait for the NVLink acks.
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = atomicAdd(a.done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    P2P_STAMP(3);
    if (threadIdx.x == 0) *a.done = 0u;
    __threadfence_system();
    announce(a.flags, a.rank, a.world, a.epoch + 1);
    wait_all(a.flags, a.rank, a.world, a.epoch + 1, a.bad_csr_out);
    P2P_STAMP(4);
  }
}

// ---------------------------------------------------------------------------------------
// The same combine step through the NVSwitch's multicast engine (NVLS; host side in mc_nvls.inl).  Every
// rank's [flags | parameters | gradients x2] block is bound at offset 0 of one multicast object, so
//   multimem.ld_reduce  sums a word over all ranks INSIDE THE SWITCH: the rank pulls its 1/G slice once
//                       (1/G of the buffer inbound per GPU instead of (G-1)/G),
//   multimem.st         stores the updated parameters into every rank's copy with one outbound write,
//   multimem.red        increments every rank's barrier counter.
// NVLink traffic per GPU and minibatch drops from 2 * (G-1)/G * N bytes to 2 * N/G.
__device__ __forceinline__ float4 mc_ld_sum_v4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ float mc_ld_sum_f32(const float* mc) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st_v4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_red_add_release(uint32_t* mc, uint32_t v) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc), "r"(v) : "memory");
}

struct McArgs {
  const float* mc_grad;            // multicast address of the CURRENT gradient buffer
  float* mc_params;                // multicast address of the item-side parameter buffer
  uint32_t* mc_flags;              // multicast address of the barrier counters ([0] gradients complete, [1] stores out)
  const uint32_t* flags;           // this rank's own counters (unicast)
  const float* params;             // this rank's parameters (unicast)
  float* acc;                      // this rank's accumulator buffer (only its slice is kept current)
  float* grad_next;                // this rank's OTHER gradient buffer: zeroed here
  unsigned int* done;
  int* bad_csr_out;
  int rank, world;
  uint32_t target;                 // world * (minibatches so far): the value both counters reach when everyone arrived
  int64_t n4;
  int64_t w_rows_end, w_end, bp_end, b_lo, b_hi;
  int ld4;
  int64_t cnt_off, steps_off;
  float lr, beta, lambda;
  int adagrad;
  unsigned long long* ts;          // nullable (cdae_debug_combine)
};

__device__ __forceinline__ void mc_wait(const uint32_t* counter, uint32_t target, int* timeout_flag) {
  if (threadIdx.x == 0) {
    const uint64_t t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(counter) - target) < 0) {
      if (global_ns() - t0 > SPIN_LIMIT_NS) {
        *timeout_flag = 2;
        break;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512) mc_step_kernel(McArgs a) {
  // stream order: every kernel that added to this rank's gradients has finished
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    P2P_STAMP(0);
    __threadfence_system();
    mc_red_add_release(a.mc_flags + 0, 1u);
  }
  mc_wait(a.flags + 0, a.target, a.bad_csr_out);
  if (blockIdx.x == 0) P2P_STAMP(1);
  __shared__ float sc_s[2];
  if (threadIdx.x < 2) sc_s[threadIdx.x] = mc_ld_sum_f32(a.mc_grad + a.steps_off + (threadIdx.x == 0 ? 0 : 2));
  __syncthreads();
  const float steps = sc_s[0];
  const bool discard = sc_s[1] != 0.f;
  if (discard && blockIdx.x == 0 && threadIdx.x == 0) *a.bad_csr_out = 1;

  const int64_t lo = a.n4 * a.rank / a.world, hi = a.n4 * (a.rank + 1) / a.world;
  constexpr int UN = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool first_pass = true;
  // (the loop bound is block-uniform: every thread takes part in the zeroing of the first pass)
  for (int64_t i0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 - (int64_t)threadIdx.x - (int64_t)blockIdx.x * blockDim.x < hi; i0 += stride * UN) {
    float4 s[UN];
    bool on[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t i = i0 + u * stride;
      on[u] = i < hi && !(i >= a.bp_end && i < a.b_lo) && i < a.b_hi;
      s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!on[u]) continue;
      s[u] = mc_ld_sum_v4(a.mc_grad + i * 4);
    }
    if (first_pass && a.grad_next) {
      // zero the gradient buffer the PREVIOUS minibatch consumed while the switch reductions are in flight
      first_pass = false;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += stride)
        *reinterpret_cast<float4*>(a.grad_next + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (!on[u] || discard) continue;
      const int64_t i = i0 + u * stride;
      const float lin = i >= a.b_lo ? a.lambda * steps : 0.f;   // item rows: scatter_kernel added lambda*W[j] itself
      const float4 sv = s[u];
      if (lin == 0.f && sv.x == 0.f && sv.y == 0.f && sv.z == 0.f && sv.w == 0.f) continue;
      const float4 w4 = *reinterpret_cast<const float4*>(a.params + i * 4);
      float g[4] = {sv.x, sv.y, sv.z, sv.w};
      float w[4] = {w4.x, w4.y, w4.z, w4.w};
      if (lin != 0.f) {
#pragma unroll
        for (int k = 0; k < 4; ++k) g[k] += lin * w[k];
      }
      if (a.adagrad) {
        const float4 a4 = *reinterpret_cast<const float4*>(a.acc + i * 4);
        float ac[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (g[k] == 0.f) continue;
          ac[k] += g[k] * g[k];
          g[k] = g[k] / (a.beta + sqrtf(ac[k]));
        }
        *reinterpret_cast<float4*>(a.acc + i * 4) = make_float4(ac[0], ac[1], ac[2], ac[3]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) w[k] -= a.lr * g[k];
      mc_st_v4(a.mc_params + i * 4, make_float4(w[0], w[1], w[2], w[3]));
    }
  }
  if (blockIdx.x == 0) P2P_STAMP(2);
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = atomicAdd(a.done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    P2P_STAMP(3);
    if (threadIdx.x == 0) {
      *a.done = 0u;
      __threadfence_system();
      mc_red_add_release(a.mc_flags + 1, 1u);
    }
    mc_wait(a.flags + 1, a.target, a.bad_csr_out);
    P2P_STAMP(4);
  }
}

}  // namespace p2p
}  // namespace cdae
