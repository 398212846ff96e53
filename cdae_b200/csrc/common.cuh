// cdae_b200/csrc/common.cuh — shared device helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cdae {

// One unit of work for a warp: up to `n` consecutive train slots [s0, s0+n) of one user.
// Two lists are built from these: input chunks (gather / scatter / sampling) and output
// chunks (decode).  32 bytes, read as two 16-byte loads.
struct __align__(16) WorkItem {
  int32_t uid;      // global user id (row of Wu / Uu, Philox key)
  int32_t u_local;  // row of the per-minibatch buffers H/Z/HG/D
  int32_t n;        // slots in this chunk
  int32_t aux0;     // index of slot s0 in the minibatch-local keep[] (and *num_neg in negs[])
  int64_t s0;       // absolute CSR slot of the first item
  int32_t first;    // 1 if this is the first chunk of its user
  int32_t row_off;  // s0 - row_ptr[uid]: position inside the user's row (Philox counter)
};
static_assert(sizeof(WorkItem) == 32, "WorkItem must be 32 bytes");

// ---------------------------------------------------------------------------------------
// Philox4x32-10.  Same specification as oracle/cdae_oracle.c (which tests compare against):
//   key = (seed lo, seed hi); see DESIGN.md "Sampling" for the counter layout.
struct Philox4 {
  uint32_t x, y, z, w;
};
__host__ __device__ __forceinline__ Philox4 philox4x32(uint64_t seed, uint32_t c0, uint32_t c1,
                                                       uint32_t c2, uint32_t c3) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}
__host__ __device__ __forceinline__ uint32_t philox_word(const Philox4& p, int i) {
  return i == 0 ? p.x : (i == 1 ? p.y : (i == 2 ? p.z : p.w));
}

// ---------------------------------------------------------------------------------------
// 16-byte vector reduction to global memory (sm_90+): one L2 atomic transaction per lane
// instead of four.
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float dot4(float4 a, float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float4 fma4(float s, float4 a, float4 c) {  // s*a + c
  return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float4 mul4(float4 a, float4 b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 scale4(float s, float4 a) {
  return make_float4(s * a.x, s * a.y, s * a.z, s * a.w);
}

// ---------------------------------------------------------------------------------------
// activation / loss, restating cdae.hpp:391-414 and loss.hpp in fp32
__device__ __forceinline__ float act_sigmoid(float x) {  // cdae.hpp:393-401
  if (x > 18.f) return 1.f;
  if (x < -18.f) return 0.f;
  return 1.f / (1.f + expf(-x));
}
__device__ __forceinline__ float act_tanh(float x) {  // cdae.hpp:403-412
  if (x > 9.f) return 1.f;
  if (x < -9.f) return -1.f;
  const float r = expf(-2.f * x);
  return (1.f - r) / (1.f + r);
}

enum { LOSS_SQUARE = 0, LOSS_LOGISTIC = 1, LOSS_LOG = 2, LOSS_HINGE = 3, LOSS_SQUARED_HINGE = 4,
       LOSS_CE = 5, LOSS_LOGM = 6 };

// Returns dl/dy; *loss receives l(y, t).  `bad` is set for LOGISTIC outside its domain
// (the reference CHECK-aborts, loss.hpp:85,96).  The two losses that are meaningful for CDAE
//  use the SFU exp/log (ex2.approx / lg2.approx, |rel err| ~ 2^-21 over
// the clamped range |y| <= 18) — the decode kernel is issue-bound and libm's expf/log1pf cost
// ~40 instructions per output.
__device__ __forceinline__ float loss_grad(int lt, float y, float t, float* loss, int* bad) {
  switch (lt) {
    case LOSS_CE: {  // loss.hpp:132-147
      // Branch-free: with a = e^-|y|, sigma(y) = 1/(1+a) (y >= 0) or a/(1+a) (y < 0) and
      // log(1+e^-y) = max(-y,0) + log(1+a).  The reference switches to e^-y / e^y / -y beyond
      // |y| = 18; those are the same functions to < 2e-8 absolute (1+e^-18 rounds to 1 in fp32),
      // far inside the 1e-4 parity tolerance, and this form cannot overflow.
      const float a = __expf(-fabsf(y));
      const float d = 1.f + a;
      const float r = __frcp_rn(d);
      *loss = fmaf(1.f - t, y, fmaxf(-y, 0.f)) + __logf(d);
      return (y >= 0.f ? r : a * r) - t;
    }
    case LOSS_LOGISTIC: {  // loss.hpp:84-99
      if (!(y > 0.f && y < 1.f)) {
        *bad = 1;
        *loss = 0.f;
        return 0.f;
      }
      *loss = (t == 0.f) ? -logf(fmaxf(0.0001f, 1.f - y)) : -logf(fmaxf(0.0001f, y));
      return (y - t) / (y * (1.f - y));
    }
    case LOSS_LOG: {  // loss.hpp:180-198
      const float z = y * t;
      if (z > 18.f) {
        *loss = expf(-z);
        return -t * expf(-z);
      }
      if (z < -18.f) {
        *loss = -z;
        return -t;
      }
      *loss = log1pf(expf(-z));
      return -t / (1.f + expf(z));
    }
    case LOSS_LOGM: {  // loss.hpp:230-246
      if (y > 18.f) {
        *loss = t * expf(-y);
        return -t * expf(-y);
      }
      if (y < -18.f) {
        *loss = -y * t;
        return -t;
      }
      *loss = t * log1pf(expf(-y));
      return -t / (1.f + expf(y));
    }
    case LOSS_HINGE: {  // loss.hpp:279-291
      const float z = y * t;
      if (z > 1.f) {
        *loss = 0.f;
        return 0.f;
      }
      *loss = 1.f - z;
      return -t;
    }
    case LOSS_SQUARED_HINGE: {  // loss.hpp:322-335
      const float z = y * t;
      if (z > 1.f) {
        *loss = 0.f;
        return 0.f;
      }
      const float d = 1.f - z;
      *loss = 0.5f * d * d;
      return -t * d;
    }
    default: {  // SQUARE, loss.hpp:48-55 (also Loss::create's default)
      const float err = t - y;
      *loss = err * err;
      return -2.f * err;
    }
  }
}

}  // namespace cdae
