// cdae_b200/csrc/api.cu — the extern "C" layer of include/cdae_b200.h: owns device memory,
// builds the per-minibatch work lists, queues the kernels of train_kernels.cuh /
// topn_kernels.cuh on one stream, and (optionally) all-reduces the dense gradients with NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>   // topn probe pass: items ordered by mean-user score (topn_api.inl)

#include "../../include/cdae_b200.h"
#include "handle.cuh"
#include "topn_kernels.cuh"
#include "topn_tc.cuh"
#include "fulldec_tc.cuh"
#include "p2p_allreduce.cuh"
#include "train_kernels.cuh"
#include "probe_kernels.cuh"

using namespace cdae;

// ------------------------------------------------------------------------------------------
// errors
static thread_local std::string g_last_error;

int cdae::set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return set_error(CDAE_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),  \
                       __FILE__, __LINE__);                                                   \
  } while (0)
#define TRY(call)            \
  do {                       \
    int rc__ = (call);       \
    if (rc__ != 0) return rc__; \
  } while (0)
#define KERNEL_OK(h)                                                                   \
  do {                                                                                 \
    ++(h)->launches;                                                                   \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess)                                                            \
      return set_error(CDAE_E_CUDA, "kernel launch failed: %s (%s:%d)",                \
                       cudaGetErrorString(e__), __FILE__, __LINE__);                   \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Optional per-kernel-class device timing (cdae_profile): an event pair around every launch of
// the class, summed after the call's final synchronise.  Off by default (no events recorded).
struct ProfScope {
  cdae_handle* h;
  int cls;
  ProfScope(cdae_handle* h_, int cls_) : h(h_), cls(cls_) {
    if (h->profiling) h->prof_begin(cls);
  }
  ~ProfScope() {
    if (h->profiling) h->prof_end();
  }
};

// ------------------------------------------------------------------------------------------
// NCCL, loaded lazily so a single-GPU process never needs it
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
const int kNcclFloat = 7, kNcclDouble = 8, kNcclSum = 0;

int load_nccl() {
  if (g_nccl.lib) return 0;
  // prefer a copy that is already mapped (torch bundles one), then the system library
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return set_error(CDAE_E_NCCL, "cannot load libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    return set_error(CDAE_E_NCCL, "libnccl.so.2 lacks the expected symbols");
  g_nccl.lib = lib;
  return 0;
}
#define NC(call)                                                                               \
  do {                                                                                         \
    int r__ = (call);                                                                          \
    if (r__ != 0)                                                                              \
      return set_error(CDAE_E_NCCL, "%s failed: %s", #call,                                    \
                       g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?");              \
  } while (0)
}  // namespace

// ------------------------------------------------------------------------------------------
// handle helpers
static int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

static float* param_ptr(cdae_handle* h, int which, int64_t* rows, int64_t* cols, int* ldp) {
  ModelDev& m = h->m;
  float* p = nullptr;
  int64_t r = 0, c = h->K;
  int ld = h->ld;
  switch (which) {
    case CDAE_P_W: p = m.W; r = h->I; break;
    case CDAE_P_W_AG: p = m.W_ag; r = h->I; break;
    case CDAE_P_V: p = m.V; r = m.V ? h->I : 0; break;
    case CDAE_P_V_AG: p = m.V_ag; r = m.V_ag ? h->I : 0; break;
    case CDAE_P_WU: p = m.Wu; r = m.Wu ? h->U : 0; break;
    case CDAE_P_WU_AG: p = m.Wu_ag; r = m.Wu_ag ? h->U : 0; break;
    case CDAE_P_UU: p = m.Uu; r = m.Uu ? h->U : 0; break;
    case CDAE_P_UU_AG: p = m.Uu_ag; r = m.Uu_ag ? h->U : 0; break;
    // vectors are stored as ONE padded row (b: [ld]) or a plain array (b': [I])
    case CDAE_P_B: p = m.b; r = 1; c = h->K; break;
    case CDAE_P_B_AG: p = m.b_ag; r = 1; c = h->K; break;
    case CDAE_P_BPRIME: p = m.bp; r = 1; c = h->I; ld = (int)h->I; break;
    case CDAE_P_BPRIME_AG: p = m.bp_ag; r = 1; c = h->I; ld = (int)h->I; break;
    default: return nullptr;
  }
  if (r == 0) c = 0;
  *rows = r; *cols = c; *ldp = ld;
  return p;
}

static int fill_table(cdae_handle* h, float* p, int64_t rows, int K, int ld, float fill) {
  if (rows * ld > 0) {
    fill_kernel<<<cdiv(rows * ld, 256), 256, 0, h->stream>>>(p, rows, K, ld, fill);
    KERNEL_OK(h);
  }
  return 0;
}
static int alloc_table(cdae_handle* h, float** p, int64_t rows, int K, int ld, float fill) {
  CU(cudaMalloc(p, sizeof(float) * (size_t)std::max<int64_t>(rows * ld, 4)));
  h->dev_bytes += sizeof(float) * (size_t)(rows * ld);
  return fill_table(h, *p, rows, K, ld, fill);
}

// The item side lives in THREE flat buffers of one layout, [W | V (asymmetric) | b' | counts x2 | b | steps]:
// parameters, AdaGrad state, and the per-minibatch gradients (the counts / steps slots exist only in the
// gradient buffer; they are padding in the other two).  Element i of the gradient buffer is the
// gradient of element i of the parameter buffer, so the optimiser step — and its slice-per-rank form
// over NVLink peer memory (p2p_allreduce.cuh) — is one flat elementwise pass.
static void point_item_side(cdae_handle* h, float* grad_base) {
  ModelDev& m = h->m;
  const int64_t I = h->I;
  const bool asym = h->cfg.asymmetric != 0;
  float *p = h->item_params.p, *a = h->item_acc.p, *g = grad_base;
  size_t off = 0;
  m.W = p + off; m.W_ag = a + off; m.gW = g + off; off += (size_t)(I * h->ld);
  if (asym) { m.V = p + off; m.V_ag = a + off; m.gV = g + off; off += (size_t)(I * h->ld); }
  else { m.V = nullptr; m.V_ag = nullptr; m.gV = nullptr; }
  m.bp = p + off; m.bp_ag = a + off; m.gbp = g + off; off += (size_t)h->I4;
  m.gcnt = g + off; off += 2 * (size_t)h->I4;
  m.b = p + off; m.b_ag = a + off; m.gb = g + off; off += (size_t)h->ld;
  m.g_steps = g + off;
}

template <class T>
static int ensure(cdae_handle* h, DevBuf<T>& b, size_t n) {
  if (b.cap >= n && b.p) return 0;
  if (b.p) CU(cudaFree(b.p));
  b.p = nullptr;
  size_t cap = std::max<size_t>(n + n / 4, 256);
  CU(cudaMalloc(&b.p, cap * sizeof(T)));
  b.cap = cap;
  (void)h;
  return 0;
}

// Chunk the rows of `users` (global ids, in the order given) into warp work items.
// aux offsets count slots from the first listed user, u_local counts users from it.
static void build_items(const int64_t* row_ptr, const int64_t* users, int64_t n_users,
                        int64_t uid_base, int ch_in, int ch_out, std::vector<WorkItem>* in,
                        std::vector<WorkItem>* out, int64_t* total_slots) {
  int64_t aux = 0;
  for (int64_t i = 0; i < n_users; ++i) {
    const int64_t uid = users ? users[i] : uid_base + i;
    const int64_t s0 = row_ptr[uid], n = row_ptr[uid + 1] - s0;
    for (int pass = 0; pass < 2; ++pass) {
      const int ch = pass == 0 ? ch_in : ch_out;
      std::vector<WorkItem>* dst = pass == 0 ? in : out;
      if (!dst) continue;
      for (int64_t o = 0; o < n; o += ch) {
        WorkItem w;
        w.uid = (int32_t)uid;
        w.u_local = (int32_t)i;
        w.n = (int32_t)std::min<int64_t>(ch, n - o);
        w.aux0 = (int32_t)(aux + o);
        w.s0 = s0 + o;
        w.first = o == 0;
        w.row_off = (int32_t)o;
        dst->push_back(w);
      }
    }
    aux += n;
  }
  if (total_slots) *total_slots = aux;
}

static int validate_csr(int64_t U, int64_t I, const int64_t* rp, const int32_t* col) {
  if (rp[0] != 0) return set_error(CDAE_E_INVALID, "row_ptr[0] must be 0");
  for (int64_t u = 0; u < U; ++u) {
    if (rp[u + 1] < rp[u]) return set_error(CDAE_E_INVALID, "row_ptr not monotone at user %lld", (long long)u);
    for (int64_t s = rp[u]; s < rp[u + 1]; ++s) {
      if (col[s] < 0 || col[s] >= I)
        return set_error(CDAE_E_INVALID, "item id %d of user %lld outside [0,%lld)", col[s], (long long)u, (long long)I);
      if (s > rp[u] && col[s - 1] >= col[s])
        return set_error(CDAE_E_INVALID, "row of user %lld is not strictly ascending", (long long)u);
    }
  }
  return 0;
}

// (Re)build the epoch plan: this rank's slice of every global minibatch, as contiguous ranges
// of one work-item list.  Depends on row_ptr only.
static int build_plan(cdae_handle* h) {
  const int64_t B = h->batch_users;
  const int64_t U = h->U;
  const int64_t n_mb = (U + B - 1) / B;
  h->plan.clear();
  std::vector<WorkItem> in, out;
  std::vector<int32_t> uids;
  in.reserve((size_t)(U + h->nnz / h->ch_in));
  out.reserve((size_t)(U + h->nnz / h->ch_out));
  int64_t max_slots = 0, max_users = 0;
  for (int64_t mb = 0; mb < n_mb; ++mb) {
    const int64_t lo = mb * B, n = std::min<int64_t>(B, U - lo);
    const int64_t a = lo + (n * h->rank) / h->world, b = lo + (n * (h->rank + 1)) / h->world;
    MiniBatch p;
    p.user0 = (int64_t)uids.size();
    p.uid0 = a;
    p.n_users = b - a;
    p.in0 = (int64_t)in.size();
    p.out0 = (int64_t)out.size();
    int64_t slots = 0;
    build_items(h->row_ptr_h.data(), nullptr, b - a, a, h->ch_in, h->ch_out, &in, &out, &slots);
    for (int64_t u = a; u < b; ++u) {
      if (h->row_ptr_h[u + 1] == h->row_ptr_h[u])
        return set_error(CDAE_E_INVALID, "user %lld has no train item (reference CHECK, cdae.hpp:139)", (long long)u);
      uids.push_back((int32_t)u);
    }
    p.n_in = (int64_t)in.size() - p.in0;
    p.n_out = (int64_t)out.size() - p.out0;
    p.slots = slots;
    max_slots = std::max(max_slots, slots);
    max_users = std::max(max_users, p.n_users);
    h->plan.push_back(p);
  }
  TRY(ensure(h, h->plan_in, in.size()));
  TRY(ensure(h, h->plan_out, out.size()));
  TRY(ensure(h, h->plan_uids, uids.size()));
  CU(cudaMemcpyAsync(h->plan_in.p, in.data(), in.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->plan_out.p, out.data(), out.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->plan_uids.p, uids.data(), uids.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));  // the host vectors go out of scope
  h->plan_max_slots = max_slots;
  h->plan_max_users = max_users;
  h->plan_n_uids = uids.size();
  h->plan_valid = true;
  return 0;
}

static int ensure_scratch(cdae_handle* h, int64_t users, int64_t slots) {
  TRY(ensure(h, h->keep, (size_t)std::max<int64_t>(slots, 1)));
  TRY(ensure(h, h->negs, (size_t)std::max<int64_t>(slots * std::max(h->cfg.num_neg, 1), 1)));
  // H | HG | GU are zeroed together each minibatch; Z and D are fully overwritten
  const size_t per = (size_t)std::max<int64_t>(users, 1) * h->ld;
  TRY(ensure(h, h->acc3, per * 3));
  TRY(ensure(h, h->zd, per * 2));
  h->scratch_users = users;
  return 0;
}

static BatchDev make_batch(cdae_handle* h, const WorkItem* in, int64_t n_in, const WorkItem* out,
                           int64_t n_out, const int32_t* uids, int64_t n_users) {
  BatchDev bt;
  const size_t per = (size_t)std::max<int64_t>(h->scratch_users, 1) * h->ld;
  bt.in_items = in; bt.out_items = out;
  bt.n_in_items = (int)n_in; bt.n_out_items = (int)n_out; bt.n_users = (int)n_users;
  bt.uids = uids;
  bt.row_ptr = h->row_ptr_d.p; bt.col = h->col_d.p;
  bt.keep = h->keep.p; bt.negs = h->negs.p;
  bt.H = h->acc3.p; bt.HG = h->acc3.p + per; bt.GU = h->acc3.p + 2 * per;
  bt.Z = h->zd.p; bt.D = h->zd.p + per;
  bt.flags = 0;
  bt.ch_in = h->ch_in;
  return bt;
}

// Leading dimension for K columns: whole 128-byte lines, rounded up to a size one <G,NV>
// geometry covers exactly (4*G*NV floats), so the kernels need no column bound checks.
static int row_stride(int K) {
  const int lines = (K + 31) / 32;
  const int l = lines <= 4 ? lines : lines <= 6 ? 6 : lines <= 8 ? 8 : lines <= 12 ? 12 : 16;
  return l * 32;
}

// launch one of the <G,NV> row-geometry instantiations by leading dimension (ld is a multiple
// of 32 floats): a warp-level 16-byte access covers one 128-byte line per row with G = 8 lanes,
// two lines with 16, four with 32; NV = vectors per lane.
#define DISPATCH_LD(ld, CALL)                                          \
  do {                                                                 \
    const int lines__ = (ld) / 32;                                     \
    if (lines__ <= 1) { CALL(8, 1); }                                  \
    else if (lines__ <= 2) { CALL(8, 2); }                             \
    else if (lines__ <= 3) { CALL(8, 3); }                             \
    else if (lines__ <= 4) { CALL(8, 4); }                             \
    else if (lines__ <= 6) { CALL(16, 3); }                            \
    else if (lines__ <= 8) { CALL(16, 4); }                            \
    else if (lines__ <= 12) { CALL(32, 3); }                           \
    else { CALL(32, 4); }                                              \
  } while (0)

static int launch_gather(cdae_handle* h, const BatchDev& bt, const SampleArgs* sa, bool count_kept) {
  if (bt.n_in_items == 0) return 0;
  ProfScope ps(h, CDAE_K_GATHER);
  const int grid = cdiv((int64_t)bt.n_in_items * 32, 256);
  StatsDev* st = count_kept ? h->stats_d : nullptr;
  if (sa) {
#define CALL(G, NV) gather_kernel<G, NV, true><<<grid, 256, 0, h->stream>>>(h->m, bt, *sa, st)
    DISPATCH_LD(h->ld, CALL);
#undef CALL
  } else {
    SampleArgs none{};
#define CALL(G, NV) gather_kernel<G, NV, false><<<grid, 256, 0, h->stream>>>(h->m, bt, none, st)
    DISPATCH_LD(h->ld, CALL);
#undef CALL
  }
  KERNEL_OK(h);
  return 0;
}
// H3 + H4 in one kernel for the epoch path (encode_fused_kernel): mode 1 = register-direct row loads,
// 2 = rows staged in shared memory by cp.async.bulk.  Users with more than one input chunk are finished
// by activate_kernel (bt.flags carries BATCH_FUSED_ENCODE so it skips the others).
static int launch_encode_fused(cdae_handle* h, const BatchDev& bt, const SampleArgs& sa, int mode) {
  if (bt.n_in_items == 0) return 0;
  ProfScope ps(h, CDAE_K_GATHER);
  const int grid = cdiv((int64_t)bt.n_in_items * 32, 256);
  if (mode == 2) {
#define CALL(G, NV) encode_fused_kernel<G, NV, true><<<grid, 256, 0, h->stream>>>(h->m, bt, sa, h->stats_d, h->m.scale)
    DISPATCH_LD(h->ld, CALL);
#undef CALL
  } else {
#define CALL(G, NV) encode_fused_kernel<G, NV, false><<<grid, 256, 0, h->stream>>>(h->m, bt, sa, h->stats_d, h->m.scale)
    DISPATCH_LD(h->ld, CALL);
#undef CALL
  }
  KERNEL_OK(h);
  return 0;
}
static int launch_scatter(cdae_handle* h, const BatchDev& bt) {
  if (bt.n_in_items == 0) return 0;
  ProfScope ps(h, CDAE_K_SCATTER);
  const int grid = cdiv((int64_t)bt.n_in_items * 32, 256);
#define CALL(G, NV) scatter_kernel<G, NV><<<grid, 256, 0, h->stream>>>(h->m, bt)
  DISPATCH_LD(h->ld, CALL);
#undef CALL
  KERNEL_OK(h);
  return 0;
}
static int launch_decode(cdae_handle* h, const BatchDev& bt, bool train, const SampleArgs* sa) {
  if (bt.n_out_items == 0) return 0;
  ProfScope ps(h, CDAE_K_DECODE);
  const int grid = cdiv((int64_t)bt.n_out_items * 32, 256);
  SampleArgs none{};
  if (train && sa) {
    // the epoch path: loss fixed at compile time for the two losses CDAE is used with
    if (h->m.loss == LOSS_CE) {
#define CALL(G, NV) decode_kernel<G, NV, true, true, LOSS_CE><<<grid, 256, 0, h->stream>>>(h->m, bt, *sa, h->stats_d)
      DISPATCH_LD(h->ld, CALL);
#undef CALL
    } else if (h->m.loss == LOSS_SQUARE) {
#define CALL(G, NV) decode_kernel<G, NV, true, true, LOSS_SQUARE><<<grid, 256, 0, h->stream>>>(h->m, bt, *sa, h->stats_d)
      DISPATCH_LD(h->ld, CALL);
#undef CALL
    } else {
#define CALL(G, NV) decode_kernel<G, NV, true, true, -1><<<grid, 256, 0, h->stream>>>(h->m, bt, *sa, h->stats_d)
      DISPATCH_LD(h->ld, CALL);
#undef CALL
    }
  } else if (train) {
#define CALL(G, NV) decode_kernel<G, NV, true, false, -1><<<grid, 256, 0, h->stream>>>(h->m, bt, none, h->stats_d)
    DISPATCH_LD(h->ld, CALL);
#undef CALL
  } else {
#define CALL(G, NV) decode_kernel<G, NV, false, false, -1><<<grid, 256, 0, h->stream>>>(h->m, bt, none, h->stats_d)
    DISPATCH_LD(h->ld, CALL);
#undef CALL
  }
  KERNEL_OK(h);
  return 0;
}
static int launch_activate(cdae_handle* h, const BatchDev& bt, float scale) {
  if (bt.n_users == 0) return 0;
  ProfScope ps(h, CDAE_K_ACTIVATE);
  activate_kernel<<<cdiv((int64_t)bt.n_users * (h->ld / 4), 256), 256, 0, h->stream>>>(h->m, bt, scale);
  KERNEL_OK(h);
  return 0;
}
static SampleArgs make_sample_args(cdae_handle* h, uint64_t seed, uint32_t pass) {
  SampleArgs sa;
  sa.seed = seed;
  sa.pass = pass;
  const double q = h->cfg.corruption_ratio;
  sa.keep_mode = 2;
  sa.keep_thr = 0;
  if (q <= 0.) sa.keep_mode = 0;
  else if (q >= 1.) sa.keep_mode = 1;
  else sa.keep_thr = (uint32_t)std::floor(q * 4294967296.0);
  return sa;
}

static int run_fulldec(cdae_handle* h, const BatchDev& bt);   // fulldec_api.inl
static void mc_release(cdae_handle* h);                       // mc_nvls.inl
static int combine_and_apply(cdae_handle* h);

// gather -> activate -> decode -> hidden_backward -> scatter -> [all-reduce] -> apply
static int run_train_minibatch(cdae_handle* h, const BatchDev& bt_in, const SampleArgs* sa) {
  const size_t per = (size_t)std::max<int64_t>(h->scratch_users, 1) * h->ld;
  CU(cudaMemsetAsync(h->acc3.p, 0, sizeof(float) * per * (h->m.linear_function ? 3 : 2), h->stream));
  // CDAE_B200_ENCODE: split (gather_kernel + activate_kernel, default) | fused | tma — the A/B of the
  // north-star encode (profiles/r02_e_*); only the epoch path (device-side sampling) has the fused forms
  static const int enc_mode = [] {
    const char* e = getenv("CDAE_B200_ENCODE");
    return !e ? 0 : strcmp(e, "fused") == 0 ? 1 : strcmp(e, "tma") == 0 ? 2 : 0;
  }();
  BatchDev bt = bt_in;
  if (sa && enc_mode != 0) {
    bt.flags |= BATCH_FUSED_ENCODE;
    TRY(launch_encode_fused(h, bt, *sa, h->ld <= 64 ? enc_mode : 1));   // (the staged form exists for ld <= 64)
  } else {
    TRY(launch_gather(h, bt, sa, true));
  }
  TRY(launch_activate(h, bt, h->m.scale));
  if (h->cfg.full_decode) {
    TRY(run_fulldec(h, bt));
  } else {
    TRY(launch_decode(h, bt, true, sa));
  }
  // hidden_backward (user rows, hidden bias) and scatter (encoder item rows) both consume HG + Z and write
  // disjoint state: outside profiling they run side by side, hidden_backward on a second stream.
  const bool side = !h->profiling && bt.n_users > 0;
  if (side) {
    if (!h->side_stream) {
      CU(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&h->side_fork, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&h->side_join, cudaEventDisableTiming));
    }
    CU(cudaEventRecord(h->side_fork, h->stream));
    CU(cudaStreamWaitEvent(h->side_stream, h->side_fork, 0));
  }
  if (bt.n_users > 0) {
    ProfScope ps(h, CDAE_K_HIDDEN_BWD);
    const int bx = h->ld / 4, by = std::max(1, 256 / bx);
    hidden_backward_kernel<<<cdiv(bt.n_users, by), dim3(bx, by), sizeof(float4) * bx * by, side ? h->side_stream : h->stream>>>(h->m, bt, h->stats_d);
    KERNEL_OK(h);
  }
  if (side) CU(cudaEventRecord(h->side_join, h->side_stream));
  if (h->cfg.full_decode && !h->m.asym) {
    // tied weights: every item is an output of every user, so fd_gemm_kernel<.., true> already added
    // n*lambda*W[i]; an input item's occurrence merges into that one (cdae.hpp:249-250,342-343)
    // and the scatter must not add a second lambda term
    ModelDev keep_m = h->m;
    h->m.lambda = 0.f;
    const int rc = launch_scatter(h, bt);
    h->m = keep_m;
    TRY(rc);
  } else {
    TRY(launch_scatter(h, bt));
  }
  if (side) CU(cudaStreamWaitEvent(h->stream, h->side_join, 0));
  if (h->m.linear_function && bt.n_users > 0) {
    uu_update_kernel<<<cdiv((int64_t)bt.n_users * h->ld, 256), 256, 0, h->stream>>>(h->m, bt, h->stats_d);
    KERNEL_OK(h);
  }
  TRY(combine_and_apply(h));
  return 0;
}

// The end of a minibatch: item-side gradients of all ranks are combined and upd() is applied once.
//   single GPU                : apply_kernel
//   process group, peer memory: p2p::fused_step_kernel — reduce-scatter by peer loads, the rank's slice of
//                               the optimiser step, all-gather of the updated parameters by peer stores
//   process group, NCCL       : ncclAllReduce of the gradient buffer, then apply_kernel on every rank
static int combine_and_apply(cdae_handle* h) {
  // CDAE_B200_DEBUG_SKIP_ALLREDUCE=1: measurement aid only (ranks diverge) — isolates the cost of
  // the collective in a scaling run
  static const bool skip_allreduce = getenv("CDAE_B200_DEBUG_SKIP_ALLREDUCE") != nullptr;
  if (h->world > 1 && !skip_allreduce && h->p2p_on && h->mc_active) {
    // NVLS: the same fused step through the switch's multicast engine (p2p::mc_step_kernel)
    ProfScope ps(h, CDAE_K_ALLREDUCE);
    p2p::McArgs ma;
    const size_t cur = (size_t)h->p2p_parity * h->grad_floats, nxt = (size_t)(h->p2p_parity ^ 1) * h->grad_floats;
    float* mc_params = reinterpret_cast<float*>(h->mc_mc + 4096);
    ma.mc_grad = mc_params + h->grad_floats + cur;
    ma.mc_params = mc_params;
    ma.mc_flags = reinterpret_cast<uint32_t*>(h->mc_mc);
    ma.flags = reinterpret_cast<const uint32_t*>(h->mc_uc);
    ma.params = h->item_params.p;
    ma.acc = h->item_acc.p;
    ma.grad_next = getenv("CDAE_B200_DEBUG_NOZERO") ? nullptr : h->grad.p + nxt;   // (debug: timing without the zeroing pass)
    ma.done = h->p2p_done;
    ma.bad_csr_out = &h->stats_d->bad_csr;
    ma.rank = h->rank; ma.world = h->world;
    ma.target = (uint32_t)h->world * (++h->p2p_epoch);
    ma.n4 = (int64_t)(h->grad_floats / 4);
    const int64_t ld4 = h->ld / 4, rows4 = h->I * ld4;
    ma.w_rows_end = rows4;
    ma.w_end = rows4 * (h->m.asym ? 2 : 1);
    ma.bp_end = ma.w_end + h->I4 / 4;
    ma.b_lo = ma.bp_end + 2 * (h->I4 / 4);
    ma.b_hi = ma.b_lo + ld4;
    ma.ld4 = (int)ld4;
    ma.cnt_off = (int64_t)(h->m.gcnt - h->m.gW);
    ma.steps_off = (int64_t)(h->m.g_steps - h->m.gW);
    ma.lr = h->m.lr; ma.beta = h->m.beta; ma.lambda = h->m.lambda; ma.adagrad = h->m.adagrad;
    ma.ts = h->p2p_ts;
    p2p::mc_step_kernel<<<h->sm_count * 2, 512, 0, h->stream>>>(ma);
    KERNEL_OK(h);
    h->p2p_parity ^= 1;
    point_item_side(h, h->grad.p + (size_t)h->p2p_parity * h->grad_floats);
    return 0;
  }
  if (h->world > 1 && !skip_allreduce && h->p2p_on && h->p2p_fused) {
    ProfScope ps(h, CDAE_K_ALLREDUCE);
    p2p::FusedArgs fa;
    const size_t cur = (size_t)h->p2p_parity * h->grad_floats, nxt = (size_t)(h->p2p_parity ^ 1) * h->grad_floats;
    for (int r = 0; r < p2p::MAX_RANKS; ++r) {
      fa.grads[r] = h->p2p_bufs[r] ? h->p2p_bufs[r] + cur : nullptr;
      fa.params[r] = h->p2p_params[r];
      fa.flags[r] = h->p2p_flags[r];
    }
    fa.acc = h->item_acc.p;
    fa.grad_next = getenv("CDAE_B200_DEBUG_NOZERO") ? nullptr : h->grad.p + nxt;   // (debug: timing without the zeroing pass)
    fa.done = h->p2p_done;
    fa.bad_csr_out = &h->stats_d->bad_csr;
    fa.rank = h->rank; fa.world = h->world;
    h->p2p_epoch += 2;
    fa.epoch = h->p2p_epoch - 1;                    // the kernel announces epoch (gradients complete) and epoch + 1 (stores out)
    fa.n4 = (int64_t)(h->grad_floats / 4);
    const int64_t ld4 = h->ld / 4, rows4 = h->I * ld4;
    fa.w_rows_end = rows4;
    fa.w_end = rows4 * (h->m.asym ? 2 : 1);
    fa.bp_end = fa.w_end + h->I4 / 4;
    fa.b_lo = fa.bp_end + 2 * (h->I4 / 4);
    fa.b_hi = fa.b_lo + ld4;
    fa.ld4 = (int)ld4;
    fa.cnt_off = (int64_t)(h->m.gcnt - h->m.gW);     // slot 0 (the whole buffer is cleared every other minibatch)
    fa.steps_off = (int64_t)(h->m.g_steps - h->m.gW);
    fa.steps_slot = 0;
    fa.lr = h->m.lr; fa.beta = h->m.beta; fa.lambda = h->m.lambda; fa.adagrad = h->m.adagrad;
    fa.ts = h->p2p_ts;
    p2p::fused_step_kernel<<<h->sm_count * 2, 512, 0, h->stream>>>(fa);   // 2 resident CTAs per SM: more NVLink loads in flight
    KERNEL_OK(h);
    h->p2p_parity ^= 1;
    point_item_side(h, h->grad.p + (size_t)h->p2p_parity * h->grad_floats);
    return 0;
  }
  if (h->world > 1 && !skip_allreduce && h->p2p_on) {
    // two-shot all-reduce over NVLink peer memory + the barrier before the replicated apply
    ProfScope ps(h, CDAE_K_ALLREDUCE);
    p2p::Args pa;
    for (int r = 0; r < p2p::MAX_RANKS; ++r) { pa.bufs[r] = h->p2p_bufs[r]; pa.flags[r] = h->p2p_flags[r]; }
    pa.rank = h->rank; pa.world = h->world; pa.n4 = (int64_t)(h->grad_floats / 4);
    pa.epoch = ++h->p2p_epoch;
    p2p::reduce_kernel<<<h->sm_count, 512, 0, h->stream>>>(pa);
    KERNEL_OK(h);
    pa.epoch = ++h->p2p_epoch;
    p2p::barrier_kernel<<<1, 32, 0, h->stream>>>(pa);
    KERNEL_OK(h);
  } else if (h->world > 1 && !skip_allreduce) {
    ProfScope ps(h, CDAE_K_ALLREDUCE);
    NC(g_nccl.AllReduce(h->grad.p, h->grad.p, h->grad_floats, kNcclFloat, kNcclSum,
                        (ncclComm_t)h->comm, h->stream));
  }
  ProfScope ps_apply(h, CDAE_K_APPLY);
  ApplyArgs a;
  a.nseg = 0;
  a.lr = h->m.lr; a.beta = h->m.beta; a.adagrad = h->m.adagrad; a.g_steps = h->m.g_steps;
  a.steps_slot = h->m.steps_slot;
  // the encoder rows' lambda terms: lambda * (occurrences as a kept input) * W[j], counted by scatter_kernel
  // (tied full decode: fd_gemm_kernel already added n*lambda*W[j] for every item, the counters stay 0)
  a.seg[a.nseg++] = ApplySeg{h->m.W, h->m.W_ag, h->m.gW, h->I * h->ld / 4, 0.f,
                             h->m.gcnt + (int64_t)h->m.steps_slot * h->I4, h->m.lambda, h->ld / 4};
  if (h->m.asym) a.seg[a.nseg++] = ApplySeg{h->m.V, h->m.V_ag, h->m.gV, h->I * h->ld / 4, 0.f, nullptr, 0.f, 1};
  a.seg[a.nseg++] = ApplySeg{h->m.bp, h->m.bp_ag, h->m.gbp, h->I4 / 4, 0.f, nullptr, 0.f, 1};
  a.seg[a.nseg++] = ApplySeg{h->m.b, h->m.b_ag, h->m.gb, h->ld / 4, h->m.lambda, nullptr, 0.f, 1};
  a.cnt_clear = h->m.gcnt + (int64_t)(h->m.steps_slot ^ 1) * h->I4;
  a.n_cnt = h->I4;
  a.bad_csr_out = &h->stats_d->bad_csr;
  apply_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(a);
  KERNEL_OK(h);
  h->m.steps_slot ^= 1;  // apply cleared the other slot for the next minibatch
  return 0;
}

static int begin_call(cdae_handle* h) {
  CU(cudaSetDevice(h->cfg.device));
  h->launches = 0;
  h->h2d = h->d2h = 0;
  CU(cudaMemsetAsync(h->stats_d, 0, sizeof(StatsDev), h->stream));
  CU(cudaEventRecord(h->ev0, h->stream));
  return 0;
}
static int end_call(cdae_handle* h, cdae_epoch_stats_t* stats) {
  CU(cudaEventRecord(h->ev1, h->stream));
  CU(cudaMemcpyAsync(h->stats_h, h->stats_d, sizeof(StatsDev), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->d2h += sizeof(StatsDev);
  if (h->profiling) h->prof_collect();
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (stats) {
    double loss = 0.;
    unsigned long long outs = 0, kept = 0;
    for (int i = 0; i < STAT_STRIPES; ++i) {
      loss += h->stats_h->loss_sum[i];
      outs += h->stats_h->outputs[i];
      kept += h->stats_h->inputs_kept[i];
    }
    stats->user_steps = (int64_t)h->stats_h->user_steps;
    stats->outputs = (int64_t)outs;
    stats->inputs_kept = (int64_t)kept;
    stats->loss_sum = loss;
    stats->device_ms = ms;
    stats->kernel_launches = h->launches;
    stats->h2d_bytes = h->h2d;
    stats->d2h_bytes = h->d2h;
  }
  if (h->stats_h->bad_csr == 2)
    return set_error(CDAE_E_STATE, "a peer GPU did not reach the combine step within %d s (its process or thread failed?); "
                     "the parameters of this call are undefined", (int)(p2p::SPIN_LIMIT_NS / 1000000000ull));
  if (h->stats_h->bad_csr) h->csr_bad = true;
  if (h->stats_h->bad_csr)
    return set_error(CDAE_E_INVALID, "the CSR passed to cdae_train_epoch_csr has an item id outside [0,%lld) or a row that is "
                     "not strictly ascending; the offending minibatch and all later ones were not applied", (long long)h->I);
  if (h->stats_h->bad_loss)
    return set_error(CDAE_E_NUMERIC, "LOGISTIC loss received a score outside (0,1) "
                     "(the reference CHECK-aborts here, loss.hpp:96; use CROSS_ENTROPY)");
  return 0;
}

// ------------------------------------------------------------------------------------------
extern "C" {

int cdae_abi_version(void) { return CDAE_B200_ABI_VERSION; }
const char* cdae_last_error(void) { return g_last_error.c_str(); }

int cdae_config_default(cdae_config_t* c) {
  if (!c) return set_error(CDAE_E_INVALID, "cfg is NULL");
  memset(c, 0, sizeof(*c));
  c->lambda = 0.01; c->learn_rate = 0.1; c->corruption_ratio = 0.5; c->beta = 0.;
  c->loss_type = CDAE_LOSS_LOGISTIC; c->num_dim = 10; c->num_neg = 5; c->num_corruptions = 1;
  c->using_adagrad = 1; c->asymmetric = 0; c->user_factor = 1; c->linear = 0; c->scaled = 1;
  c->linear_function = 0; c->tanh_act = 0;
  c->batch_users = 0; c->device = 0;
  return 0;
}

int cdae_create(const cdae_config_t* cfg, int64_t U, int64_t I, const int64_t* row_ptr,
                const int32_t* col, cdae_handle** out) {
  if (!cfg || !row_ptr || !out || (!col && row_ptr[U] > 0)) return set_error(CDAE_E_INVALID, "NULL argument");
  if (U <= 0 || I <= 0 || U > 0x7fffffff || I > 0x7fffffff) return set_error(CDAE_E_INVALID, "need 0 < U, I < 2^31");
  if (cfg->num_dim < 1 || cfg->num_dim > 512) return set_error(CDAE_E_INVALID, "num_dim must be in [1,512]");
  if (cfg->num_neg < 0 || cfg->num_neg > DECODE_MAX_NEGS || cfg->num_corruptions < 1)
    return set_error(CDAE_E_INVALID, "0 <= num_neg <= %d and num_corruptions >= 1 required", DECODE_MAX_NEGS);
  if (cfg->loss_type < 0 || cfg->loss_type > CDAE_LOSS_LOGM) return set_error(CDAE_E_INVALID, "unknown loss_type %d", cfg->loss_type);
  if (!std::isfinite(cfg->corruption_ratio) || cfg->corruption_ratio < 0. || cfg->corruption_ratio > 1.)
    return set_error(CDAE_E_INVALID, "corruption_ratio must be in [0,1], got %g", cfg->corruption_ratio);
  if (!std::isfinite(cfg->lambda) || cfg->lambda < 0. || !std::isfinite(cfg->learn_rate) || !std::isfinite(cfg->beta) || cfg->beta < 0.)
    return set_error(CDAE_E_INVALID, "lambda >= 0, beta >= 0 and a finite learn_rate are required");
  if (row_ptr[U] >= 0x7fffffff) return set_error(CDAE_E_INVALID, "nnz must be < 2^31");
  if (cfg->full_decode) {
    if (cfg->loss_type != CDAE_LOSS_CROSS_ENTROPY && cfg->loss_type != CDAE_LOSS_SQUARE)
      return set_error(CDAE_E_INVALID, "full_decode needs CROSS_ENTROPY or SQUARE loss");
    if (cfg->num_dim > fd::MAX_KB * tc::KBLK)
      return set_error(CDAE_E_INVALID, "full_decode needs num_dim <= %d", fd::MAX_KB * tc::KBLK);
  }
  TRY(validate_csr(U, I, row_ptr, col));
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return set_error(CDAE_E_INVALID, "device %d of %d", cfg->device, ndev);
  CU(cudaSetDevice(cfg->device));

  cdae_handle* h = new cdae_handle();
  h->cfg = *cfg;
  h->U = U; h->I = I; h->K = cfg->num_dim;
  h->ld = row_stride(cfg->num_dim);
  h->I4 = round_up(I, 4);
  h->nnz = row_ptr[U];
  h->batch_users = cfg->batch_users > 0 ? cfg->batch_users : 16384;   // default: profiles/r02_h_batch_size_sweep.json + r02_d_* (trajectory)
  h->ch_in = 64;
  h->ch_out = std::max(1, 96 / (1 + cfg->num_neg));
  h->rank = 0; h->world = 1;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, cfg->device));
  h->sm_count = prop.multiProcessorCount;
  // full decode: one 128-user tile per SM makes every tensor kernel exactly one wave
  if (cfg->full_decode && cfg->batch_users <= 0) h->batch_users = (int64_t)128 * h->sm_count;
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&h->ev0));
  CU(cudaEventCreate(&h->ev1));
  CU(cudaMalloc(&h->stats_d, sizeof(StatsDev)));
  CU(cudaMallocHost(&h->stats_h, sizeof(StatsDev)));

  ModelDev& m = h->m;
  memset(&m, 0, sizeof(m));
  m.I = I; m.U = U; m.K = h->K; m.ld = h->ld;
  m.lambda = (float)cfg->lambda; m.lr = (float)cfg->learn_rate; m.beta = (float)cfg->beta;
  // cdae.hpp:202-205.  q == 1: the reference's scale is inf but its input set is always empty, so the
  // factor is never multiplied (:366, :377-380) — scale 1 gives the same hidden values without inf * 0.
  m.scale = (cfg->scaled && cfg->corruption_ratio < 1.) ? (float)(1. / (1. - cfg->corruption_ratio)) : 1.f;
  m.loss = cfg->loss_type; m.nu = cfg->num_neg;
  m.adagrad = cfg->using_adagrad; m.asym = cfg->asymmetric; m.user_factor = cfg->user_factor;
  m.linear = cfg->linear; m.tanh_act = cfg->tanh_act; m.linear_function = cfg->linear_function;

  int rc = 0;
  do {
    // cdae.hpp:109-134: accumulators 1e-4, b = b' = 0, Uu = 1
    // the item side: flat parameter / accumulator / gradient buffers of one layout (point_item_side)
    h->grad_floats = (size_t)(I * h->ld) * (cfg->asymmetric ? 2 : 1) + 3 * (size_t)h->I4 + (size_t)h->ld + 4;
    if ((rc = ensure(h, h->item_params, h->grad_floats))) break;
    if ((rc = ensure(h, h->item_acc, h->grad_floats))) break;
    if ((rc = ensure(h, h->grad, h->grad_floats))) break;
    h->dev_bytes += 3 * sizeof(float) * h->grad_floats;
    if (cudaMemsetAsync(h->item_params.p, 0, sizeof(float) * h->grad_floats, h->stream) != cudaSuccess ||
        cudaMemsetAsync(h->item_acc.p, 0, sizeof(float) * h->grad_floats, h->stream) != cudaSuccess ||
        cudaMemsetAsync(h->grad.p, 0, sizeof(float) * h->grad_floats, h->stream) != cudaSuccess) { rc = set_error(CDAE_E_CUDA, "memset"); break; }
    m.I4 = h->I4;
    point_item_side(h, h->grad.p);
    if ((rc = fill_table(h, m.W_ag, I, h->K, h->ld, 1e-4f))) break;
    if (cfg->asymmetric && (rc = fill_table(h, m.V_ag, I, h->K, h->ld, 1e-4f))) break;
    if ((rc = fill_table(h, m.b_ag, 1, h->K, h->ld, 1e-4f))) break;
    if ((rc = fill_table(h, m.bp_ag, 1, (int)I, (int)h->I4, 1e-4f))) break;
    if (cfg->user_factor) {
      if ((rc = alloc_table(h, &m.Wu, U, h->K, h->ld, 0.f))) break;
      if ((rc = alloc_table(h, &m.Wu_ag, U, h->K, h->ld, 1e-4f))) break;
    }
    if (cfg->linear_function) {
      if ((rc = alloc_table(h, &m.Uu, U, h->K, h->ld, 1.f))) break;
      if ((rc = alloc_table(h, &m.Uu_ag, U, h->K, h->ld, 1e-4f))) break;
    }
    // CSR
    h->row_ptr_h.assign(row_ptr, row_ptr + U + 1);
    if ((rc = ensure(h, h->row_ptr_d, (size_t)U + 1))) break;
    if ((rc = ensure(h, h->col_d, (size_t)std::max<int64_t>(h->nnz, 1)))) break;
    if (cudaMemcpyAsync(h->row_ptr_d.p, row_ptr, sizeof(int64_t) * (U + 1), cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
        cudaMemcpyAsync(h->col_d.p, col, sizeof(int32_t) * h->nnz, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) {
      rc = set_error(CDAE_E_CUDA, "CSR upload failed: %s", cudaGetErrorString(cudaGetLastError()));
      break;
    }
  } while (0);
  if (rc) {
    cdae_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

int cdae_destroy(cdae_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)h->comm);
  for (int r = 0; r < 8 && h->p2p_ipc; ++r) {
    if (r == h->rank) continue;
    if (h->p2p_bufs[r]) cudaIpcCloseMemHandle(h->p2p_bufs[r]);
    if (h->p2p_params[r]) cudaIpcCloseMemHandle(h->p2p_params[r]);
    if (h->p2p_flags[r]) cudaIpcCloseMemHandle(h->p2p_flags[r]);
  }
  if (h->p2p_my_flags) cudaFree(h->p2p_my_flags);
  if (h->p2p_ts) cudaFree(h->p2p_ts);
  mc_release(h);   // NVLS mode: the item-side buffers belong to a VMM block, not to cudaMalloc
  ModelDev& m = h->m;
  float* tabs[] = {m.Wu, m.Uu, m.Wu_ag, m.Uu_ag};
  for (float* p : tabs) if (p) cudaFree(p);
  h->item_params.release(); h->item_acc.release();
  h->grad.release(); h->row_ptr_d.release(); h->col_d.release();
  h->plan_in.release(); h->plan_out.release(); h->plan_uids.release();
  h->tmp_in.release(); h->tmp_out.release(); h->tmp_uids.release();
  h->keep.release(); h->negs.release(); h->acc3.release(); h->zd.release();
  h->stage_d.release(); h->stage_f.release();
  h->topn_ids.release(); h->topn_scores.release(); h->topn_z.release();
  h->cand_id.release(); h->cand_cnt.release(); h->cand_s.release(); h->flag_d.release();
  h->test_rp_d.release(); h->test_col_d.release();
  h->tc_zb.release(); h->tc_wb.release(); h->tc_wmax.release(); h->tc_eps.release();
  h->tc_thr.release(); h->tc_redo.release(); h->tc_redo_thr.release();
  h->tc_probe_wb.release(); h->tc_probe_keys.release(); h->tc_probe_ids.release(); h->tc_probe_pos.release();
  h->tc_probe_bits.release(); h->tc_probe_thr.release(); h->tc_probe_zsum.release(); h->tc_probe_tmp.release();
  h->fd_zb.release(); h->fd_wb.release(); h->fd_g.release(); h->fd_bits.release(); h->fd_bias.release();
  if (h->stats_d) cudaFree(h->stats_d);
  if (h->stats_h) cudaFreeHost(h->stats_h);
  if (h->side_fork) cudaEventDestroy(h->side_fork);
  if (h->side_join) cudaEventDestroy(h->side_join);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  for (cudaEvent_t e : h->copy_ev) cudaEventDestroy(e);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int cdae_init_params(cdae_handle* h, uint64_t seed) {
  if (!h) return set_error(CDAE_E_INVALID, "handle is NULL");
  CU(cudaSetDevice(h->cfg.device));
  h->topn_k = 0;  // stored recommendation lists are stale
  const double scale = 4. * std::sqrt(6. / (double)(h->I + h->K));
  struct { float* p; int64_t rows; uint32_t which; } blocks[3] = {
      {h->m.W, h->I, CDAE_P_W}, {h->m.V, h->I, CDAE_P_V}, {h->m.Wu, h->U, CDAE_P_WU}};
  for (auto& b : blocks) {
    if (!b.p) continue;
    init_uniform_kernel<<<cdiv(b.rows * h->K, 256), 256, 0, h->stream>>>(b.p, b.rows, h->K, h->ld, scale, seed, b.which);
    KERNEL_OK(h);
  }
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cdae_param_shape(cdae_handle* h, int which, int64_t* rows, int64_t* cols) {
  if (!h) return set_error(CDAE_E_INVALID, "handle is NULL");
  int64_t r, c; int ld;
  if (!param_ptr(h, which, &r, &c, &ld) && (which < 0 || which >= CDAE_P_COUNT))
    return set_error(CDAE_E_INVALID, "unknown parameter block %d", which);
  // report vectors the way the reference declares them: b is K x 1, b' is I x 1
  if (which == CDAE_P_B || which == CDAE_P_B_AG || which == CDAE_P_BPRIME || which == CDAE_P_BPRIME_AG) { r = c; c = 1; }
  if (rows) *rows = r;
  if (cols) *cols = c;
  return 0;
}

int cdae_set_param(cdae_handle* h, int which, const double* src, int64_t n) {
  if (!h || (!src && n > 0)) return set_error(CDAE_E_INVALID, "NULL argument");
  CU(cudaSetDevice(h->cfg.device));
  int64_t r, c; int ld;
  float* p = param_ptr(h, which, &r, &c, &ld);
  if (which < 0 || which >= CDAE_P_COUNT) return set_error(CDAE_E_INVALID, "unknown parameter block %d", which);
  if (n != r * c) return set_error(CDAE_E_INVALID, "block %d holds %lld values, got %lld", which, (long long)(r * c), (long long)n);
  if (n == 0) return 0;
  h->topn_k = 0;  // stored recommendation lists are stale
  const bool vec_bp = which == CDAE_P_BPRIME || which == CDAE_P_BPRIME_AG;
  TRY(ensure(h, h->stage_d, (size_t)n));
  CU(cudaMemcpyAsync(h->stage_d.p, src, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  const int K = (int)c, L = vec_bp ? (int)c : ld;
  pack_from_double_kernel<<<cdiv(r * L, 256), 256, 0, h->stream>>>(p, h->stage_d.p, r, K, L);
  KERNEL_OK(h);
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cdae_get_param(cdae_handle* h, int which, double* dst, int64_t n) {
  if (!h || (!dst && n > 0)) return set_error(CDAE_E_INVALID, "NULL argument");
  CU(cudaSetDevice(h->cfg.device));
  int64_t r, c; int ld;
  float* p = param_ptr(h, which, &r, &c, &ld);
  if (which < 0 || which >= CDAE_P_COUNT) return set_error(CDAE_E_INVALID, "unknown parameter block %d", which);
  if (n != r * c) return set_error(CDAE_E_INVALID, "block %d holds %lld values, got %lld", which, (long long)(r * c), (long long)n);
  if (n == 0) return 0;
  const bool vec_bp = which == CDAE_P_BPRIME || which == CDAE_P_BPRIME_AG;
  const bool user_block = which == CDAE_P_WU || which == CDAE_P_WU_AG || which == CDAE_P_UU || which == CDAE_P_UU_AG;
  TRY(ensure(h, h->stage_d, (size_t)n));
  const int K = (int)c, L = vec_bp ? (int)c : ld;
  const float* src = p;
  if (h->world > 1 && user_block) {
    // user rows are only ever updated by their owning rank: keep owned rows, zero the rest,
    // and sum across ranks (rows nobody trained are identical everywhere -> divide is avoided
    // by letting rank ownership cover every row exactly once).
    TRY(ensure(h, h->stage_f, (size_t)(r * ld)));
    owned_rows_kernel<<<cdiv(r * ld, 256), 256, 0, h->stream>>>(h->stage_f.p, p, r, ld, h->batch_users, h->U, h->rank, h->world);
    KERNEL_OK(h);
    NC(g_nccl.AllReduce(h->stage_f.p, h->stage_f.p, (size_t)(r * ld), kNcclFloat, kNcclSum, (ncclComm_t)h->comm, h->stream));
    src = h->stage_f.p;
  }
  const bool item_acc_block = which == CDAE_P_W_AG || which == CDAE_P_V_AG || which == CDAE_P_B_AG || which == CDAE_P_BPRIME_AG;
  if (h->world > 1 && h->p2p_on && h->p2p_fused && item_acc_block) {
    // fused peer-memory mode shards the AdaGrad state: rank q keeps elements [n*q/world, n*(q+1)/world) of the
    // flat item-side buffer current.  Keep the owned part of the block, zero the rest, sum across ranks.
    const int64_t nblk = vec_bp ? h->I4 : r * (int64_t)ld;
    const int64_t n4 = (int64_t)(h->grad_floats / 4);
    const int64_t lo = 4 * (n4 * h->rank / h->world), hi = 4 * (n4 * (h->rank + 1) / h->world);
    TRY(ensure(h, h->stage_f, (size_t)nblk));
    owned_flat_kernel<<<cdiv(nblk, 256), 256, 0, h->stream>>>(h->stage_f.p, p, nblk, (int64_t)(p - h->item_acc.p), lo, hi);
    KERNEL_OK(h);
    NC(g_nccl.AllReduce(h->stage_f.p, h->stage_f.p, (size_t)nblk, kNcclFloat, kNcclSum, (ncclComm_t)h->comm, h->stream));
    src = h->stage_f.p;
  }
  unpack_to_double_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(h->stage_d.p, src, r, K, L);
  KERNEL_OK(h);
  CU(cudaMemcpyAsync(dst, h->stage_d.p, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

// selected rows of a table (or selected entries of b') as doubles
// process group: keep the gathered rows whose user this rank trains (ownership rule of build_plan), zero the rest
__global__ void zero_unowned_rows_kernel(double* dst, const int64_t* rows, int64_t n, int K, int64_t B, int64_t U,
                                         int rank, int world) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * K) return;
  const int64_t uid = rows[i / K];
  const int64_t lo = (uid / B) * B;
  const int64_t nb = min(B, U - lo);
  const int64_t a = lo + (nb * rank) / world, b = lo + (nb * (rank + 1)) / world;
  if (uid < a || uid >= b) dst[i] = 0.;
}

__global__ void gather_rows_to_double_kernel(double* dst, const float* src, const int64_t* rows,
                                             int64_t n, int K, int ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * K) return;
  dst[i] = (double)src[rows[i / K] * ld + (i % K)];
}

int cdae_get_param_rows(cdae_handle* h, int which, const int64_t* rows, int64_t n, double* dst) {
  if (!h || !rows || !dst || n <= 0) return set_error(CDAE_E_INVALID, "NULL / empty argument");
  CU(cudaSetDevice(h->cfg.device));
  int64_t r, c; int ld;
  float* p = param_ptr(h, which, &r, &c, &ld);
  if (which < 0 || which >= CDAE_P_COUNT || !p || r * c == 0) return set_error(CDAE_E_INVALID, "parameter block %d is absent", which);
  const bool vec = which == CDAE_P_B || which == CDAE_P_B_AG || which == CDAE_P_BPRIME || which == CDAE_P_BPRIME_AG;
  const bool user_block = which == CDAE_P_WU || which == CDAE_P_WU_AG || which == CDAE_P_UU || which == CDAE_P_UU_AG;
  const bool item_acc_block = which == CDAE_P_W_AG || which == CDAE_P_V_AG || which == CDAE_P_B_AG || which == CDAE_P_BPRIME_AG;
  if (h->world > 1 && item_acc_block && h->p2p_on && h->p2p_fused)
    return set_error(CDAE_E_STATE, "accumulators are sharded across ranks in peer-memory mode; use cdae_get_param");
  const int64_t limit = vec ? c : r;   // vectors: `rows` index the entries
  for (int64_t i = 0; i < n; ++i)
    if (rows[i] < 0 || rows[i] >= limit) return set_error(CDAE_E_INVALID, "row %lld outside [0,%lld)", (long long)rows[i], (long long)limit);
  const int K = vec ? 1 : (int)c, L = vec ? 1 : ld;
  // staging: [n*K doubles | n int64 row ids]
  TRY(ensure(h, h->stage_d, (size_t)(n * K + n)));
  int64_t* rows_d = reinterpret_cast<int64_t*>(h->stage_d.p + n * K);
  CU(cudaMemcpyAsync(rows_d, rows, sizeof(int64_t) * n, cudaMemcpyHostToDevice, h->stream));
  gather_rows_to_double_kernel<<<cdiv(n * K, 256), 256, 0, h->stream>>>(h->stage_d.p, p, rows_d, n, K, L);
  KERNEL_OK(h);
  if (h->world > 1 && user_block) {
    // user rows are only current on the rank that trains them: zero the others, sum across ranks (collective)
    zero_unowned_rows_kernel<<<cdiv(n * K, 256), 256, 0, h->stream>>>(h->stage_d.p, rows_d, n, K, h->batch_users, h->U, h->rank, h->world);
    KERNEL_OK(h);
    NC(g_nccl.AllReduce(h->stage_d.p, h->stage_d.p, (size_t)(n * K), kNcclDouble, kNcclSum, (ncclComm_t)h->comm, h->stream));
  }
  CU(cudaMemcpyAsync(dst, h->stage_d.p, sizeof(double) * n * K, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

// rp_host / col_host != nullptr: cdae_train_epoch_csr — the pieces of the caller's CSR are put on the copy
// stream minibatch by minibatch, each right before that minibatch's kernels are queued, so the GPU starts on
// minibatch 0 while the host is still issuing the later pieces.
static int train_epoch_impl(cdae_handle* h, uint64_t seed, int64_t epoch, cdae_epoch_stats_t* stats,
                            const int64_t* rp_host = nullptr, const int32_t* col_host = nullptr) {
  h->topn_k = 0;  // stored recommendation lists are stale
  if (!h->plan_valid) TRY(build_plan(h));
  if (h->csr_unchecked) {
    // cdae_train_epoch_csr: the caller's CSR was uploaded without a host pass over it.  One warp per
    // trained user checks it on the device (item ids in range — out-of-range ids are clamped in the
    // device copy so no kernel can index outside a table — and rows strictly ascending); a violation
    // raises stats->bad_csr and the shared flag g_steps[2], which make hidden_backward / uu_update /
    // apply skip their updates from that minibatch on; the call returns CDAE_E_INVALID.
    // The upload runs on a second stream, one piece per minibatch (cdae_train_epoch_csr), and each
    // minibatch only waits for — and checks — its own rows: PCIe transfer of minibatch k + 1 overlaps the
    // kernels of minibatch k.
    CU(cudaMemsetAsync(h->m.g_steps + 2, 0, sizeof(float), h->stream));
  }
  const bool staged = h->csr_unchecked;
  h->csr_unchecked = false;
  TRY(ensure_scratch(h, h->plan_max_users, h->plan_max_slots));
  const int cnum = h->cfg.num_corruptions;
  size_t mb_index = 0;
  for (const MiniBatch& p : h->plan) {
    ++mb_index;
    if (staged && p.n_users > 0) {
      {
        const int64_t s0 = rp_host[p.uid0], s1 = rp_host[p.uid0 + p.n_users];
        CU(cudaMemcpyAsync(h->row_ptr_d.p + p.uid0, rp_host + p.uid0, sizeof(int64_t) * (p.n_users + 1), cudaMemcpyHostToDevice, h->copy_stream));
        if (s1 > s0) CU(cudaMemcpyAsync(h->col_d.p + s0, col_host + s0, sizeof(int32_t) * (s1 - s0), cudaMemcpyHostToDevice, h->copy_stream));
        h->h2d += sizeof(int64_t) * (p.n_users + 1) + sizeof(int32_t) * (s1 - s0);
        CU(cudaEventRecord(h->copy_ev[mb_index], h->copy_stream));
      }
      CU(cudaStreamWaitEvent(h->stream, h->copy_ev[mb_index], 0));
      validate_rows_kernel<<<cdiv(p.n_users * 32, 256), 256, 0, h->stream>>>(h->plan_uids.p + p.user0, p.n_users, h->row_ptr_d.p,
                                                                          h->col_d.p, h->I, h->stats_d, h->m.g_steps + 2);
      KERNEL_OK(h);
    }
    for (int c = 0; c < cnum; ++c) {
      BatchDev bt = make_batch(h, h->plan_in.p + p.in0, p.n_in, h->plan_out.p + p.out0, p.n_out,
                               h->plan_uids.p + p.user0, p.n_users);
      const SampleArgs sa = make_sample_args(h, seed, (uint32_t)(epoch * cnum + c));
      TRY(run_train_minibatch(h, bt, &sa));
    }
  }
  return end_call(h, stats);
}

int cdae_train_epoch(cdae_handle* h, uint64_t seed, int64_t epoch, cdae_epoch_stats_t* stats) {
  if (!h) return set_error(CDAE_E_INVALID, "handle is NULL");
  if (h->csr_bad) return set_error(CDAE_E_STATE, "the resident CSR failed validation in the last cdae_train_epoch_csr; pass a valid one");
  TRY(begin_call(h));
  return train_epoch_impl(h, seed, epoch, stats);
}

int cdae_train_epoch_csr(cdae_handle* h, const int64_t* row_ptr, const int32_t* col, uint64_t seed,
                         int64_t epoch, cdae_epoch_stats_t* stats) {
  if (!h || !row_ptr || !col) return set_error(CDAE_E_INVALID, "NULL argument");
  TRY(begin_call(h));
  const int64_t nnz = row_ptr[h->U];
  if (nnz >= 0x7fffffff) return set_error(CDAE_E_INVALID, "nnz must be < 2^31");
  // the work lists depend on row_ptr only: rebuild them only when the row structure changed.  In a
  // process group a rank only ever looks at the rows of the users it trains (one contiguous range per
  // minibatch), so only those ranges are compared and uploaded — at 8 ranks x 100K users that is 1/8 of
  // the host work and of the PCIe bytes per call.
  bool same_rows = nnz == h->nnz;
  const bool sliced = h->world > 1 && h->plan_valid;
  if (same_rows) {
    if (!sliced) {
      same_rows = memcmp(row_ptr, h->row_ptr_h.data(), sizeof(int64_t) * (h->U + 1)) == 0;
    } else {
      for (const MiniBatch& p : h->plan)
        if (p.n_users > 0 && memcmp(row_ptr + p.uid0, h->row_ptr_h.data() + p.uid0, sizeof(int64_t) * (p.n_users + 1)) != 0) {
          same_rows = false;
          break;
        }
    }
  }
  if (!same_rows) {
    // row_ptr is small (U + 1 entries) and the work lists are cut from it on the host: check it here;
    // col (nnz entries) is checked on the device after the upload (validate_rows_kernel)
    if (row_ptr[0] != 0) return set_error(CDAE_E_INVALID, "row_ptr[0] must be 0");
    for (int64_t u = 0; u < h->U; ++u)
      if (row_ptr[u + 1] < row_ptr[u]) return set_error(CDAE_E_INVALID, "row_ptr not monotone at user %lld", (long long)u);
    h->row_ptr_h.assign(row_ptr, row_ptr + h->U + 1);
    h->nnz = nnz;
    h->plan_valid = false;
    TRY(ensure(h, h->col_d, (size_t)std::max<int64_t>(nnz, 1)));
  }
  // The upload runs on a second stream, one piece per minibatch this rank trains (its row_ptr range and its col
  // range), an event after each piece; train_epoch_impl issues piece k right before it queues minibatch k's
  // kernels, which wait for that event — the transfer of minibatch k + 1 rides under the kernels of minibatch
  // k.  (The previous call ended with a synchronise of the compute stream, so nothing still reads the old copy.)
  if (!h->plan_valid) TRY(build_plan(h));
  if (!h->copy_stream) CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  while (h->copy_ev.size() < h->plan.size() + 1) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->copy_ev.push_back(e);
  }
  h->csr_unchecked = true;
  h->csr_bad = false;
  const int rc = train_epoch_impl(h, seed, epoch, stats, row_ptr, col);
  if (rc != 0) cudaStreamSynchronize(h->copy_stream);   // the caller's buffers must not be read after an error return
  return rc;
}

// Upload the work lists of an explicit user list into the tmp_* buffers.
static int stage_users(cdae_handle* h, const int64_t* uids, int64_t n, bool want_out, int64_t* slots,
                       int64_t* n_in, int64_t* n_out) {
  std::vector<WorkItem> in, out;
  std::vector<int32_t> u32((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    if (uids[i] < 0 || uids[i] >= h->U) return set_error(CDAE_E_INVALID, "uid %lld outside [0,%lld)", (long long)uids[i], (long long)h->U);
    u32[(size_t)i] = (int32_t)uids[i];
  }
  build_items(h->row_ptr_h.data(), uids, n, 0, h->ch_in, h->ch_out, &in, want_out ? &out : nullptr, slots);
  TRY(ensure(h, h->tmp_in, std::max<size_t>(in.size(), 1)));
  TRY(ensure(h, h->tmp_out, std::max<size_t>(out.size(), 1)));
  TRY(ensure(h, h->tmp_uids, std::max<size_t>(u32.size(), 1)));
  if (!in.empty()) CU(cudaMemcpyAsync(h->tmp_in.p, in.data(), in.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, h->stream));
  if (!out.empty()) CU(cudaMemcpyAsync(h->tmp_out.p, out.data(), out.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, h->stream));
  if (n) CU(cudaMemcpyAsync(h->tmp_uids.p, u32.data(), u32.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *n_in = (int64_t)in.size();
  *n_out = (int64_t)out.size();
  return 0;
}

int cdae_train_users(cdae_handle* h, const int64_t* uids, int64_t n, const uint8_t* keep_mask,
                     const int32_t* negatives, cdae_epoch_stats_t* stats) {
  if (!h || !uids || n <= 0) return set_error(CDAE_E_INVALID, "need a non-empty uid list");
  TRY(begin_call(h));
  h->topn_k = 0;  // stored recommendation lists are stale
  std::vector<int64_t> own_uids;
  std::vector<uint8_t> own_keep;
  std::vector<int32_t> own_negs;
  if (h->world > 1) {
    // process group (collective; every rank passes the SAME arguments): a rank trains the listed users it
    // OWNS — the rule of build_plan; their Wu / Uu rows are current only there — and the minibatch is
    // combined like any other.  stats are this rank's part.
    {
      std::vector<int64_t> sorted(uids, uids + n);
      std::sort(sorted.begin(), sorted.end());
      if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end()) return set_error(CDAE_E_INVALID, "uids must be distinct inside one frozen minibatch");
    }
    const int nu = h->cfg.num_neg;
    int64_t off = 0;
    for (int64_t i = 0; i < n; ++i) {
      const int64_t u = uids[i];
      if (u < 0 || u >= h->U) return set_error(CDAE_E_INVALID, "uid %lld outside [0,%lld)", (long long)u, (long long)h->U);
      const int64_t len = h->row_ptr_h[u + 1] - h->row_ptr_h[u];
      const int64_t lo = (u / h->batch_users) * h->batch_users, nb = std::min<int64_t>(h->batch_users, h->U - lo);
      const int64_t a = lo + nb * h->rank / h->world, b = lo + nb * (h->rank + 1) / h->world;
      if (u >= a && u < b) {
        own_uids.push_back(u);
        if (keep_mask) own_keep.insert(own_keep.end(), keep_mask + off, keep_mask + off + len);
        if (negatives) own_negs.insert(own_negs.end(), negatives + off * nu, negatives + (off + len) * nu);
      }
      off += len;
    }
    // a rank that owns none of the listed users still takes part in the combine step
    uids = own_uids.data();
    n = (int64_t)own_uids.size();
    if (keep_mask) { own_keep.push_back(0); keep_mask = own_keep.data(); }
    if (negatives) { own_negs.push_back(0); negatives = own_negs.data(); }
  }
  int64_t slots = 0, n_in = 0, n_out = 0;
  TRY(stage_users(h, uids, n, true, &slots, &n_in, &n_out));
  if (slots > 0 && !keep_mask) return set_error(CDAE_E_INVALID, "keep_mask is NULL");
  if (h->cfg.full_decode) negatives = nullptr;   // the output set is every item
  if (slots * h->cfg.num_neg > 0 && !negatives && !h->cfg.full_decode) return set_error(CDAE_E_INVALID, "negatives is NULL");
  {  // distinct users, non-empty rows, negatives outside the user's row
    std::vector<int64_t> s(uids, uids + n);
    std::sort(s.begin(), s.end());
    if (std::adjacent_find(s.begin(), s.end()) != s.end()) return set_error(CDAE_E_INVALID, "uids must be distinct inside one frozen minibatch");
    int64_t off = 0;
    for (int64_t i = 0; i < n; ++i) {
      const int64_t nu_ = h->row_ptr_h[uids[i] + 1] - h->row_ptr_h[uids[i]];
      if (nu_ == 0) return set_error(CDAE_E_INVALID, "user %lld has no train item (reference CHECK, cdae.hpp:139)", (long long)uids[i]);
      for (int64_t j = 0; negatives && j < nu_ * h->cfg.num_neg; ++j)
        if (negatives[off * h->cfg.num_neg + j] < 0 || negatives[off * h->cfg.num_neg + j] >= h->I)
          return set_error(CDAE_E_INVALID, "negative id out of range");
      off += nu_;
    }
  }
  TRY(ensure_scratch(h, n, slots));
  CU(cudaMemcpyAsync(h->keep.p, keep_mask, (size_t)slots, cudaMemcpyHostToDevice, h->stream));
  if (slots * h->cfg.num_neg > 0 && negatives)
    CU(cudaMemcpyAsync(h->negs.p, negatives, sizeof(int32_t) * (size_t)(slots * h->cfg.num_neg), cudaMemcpyHostToDevice, h->stream));
  h->h2d += (size_t)slots + sizeof(int32_t) * (size_t)(slots * h->cfg.num_neg);
  BatchDev bt = make_batch(h, h->tmp_in.p, n_in, h->tmp_out.p, n_out, h->tmp_uids.p, n);
  TRY(run_train_minibatch(h, bt, nullptr));
  return end_call(h, stats);
}

int cdae_encode(cdae_handle* h, const int64_t* uids, int64_t n, const uint8_t* keep_mask, double scale,
                float* z_out) {
  if (!h || !uids || !z_out || n <= 0) return set_error(CDAE_E_INVALID, "NULL / empty argument");
  TRY(begin_call(h));
  int64_t slots = 0, n_in = 0, n_out = 0;
  TRY(stage_users(h, uids, n, false, &slots, &n_in, &n_out));
  TRY(ensure_scratch(h, n, slots));
  if (keep_mask) {
    CU(cudaMemcpyAsync(h->keep.p, keep_mask, (size_t)slots, cudaMemcpyHostToDevice, h->stream));
  } else {
    CU(cudaMemsetAsync(h->keep.p, 1, (size_t)std::max<int64_t>(slots, 1), h->stream));
  }
  BatchDev bt = make_batch(h, h->tmp_in.p, n_in, h->tmp_out.p, 0, h->tmp_uids.p, n);
  const size_t per = (size_t)std::max<int64_t>(h->scratch_users, 1) * h->ld;
  CU(cudaMemsetAsync(h->acc3.p, 0, sizeof(float) * per, h->stream));
  TRY(launch_gather(h, bt, nullptr, false));
  TRY(launch_activate(h, bt, (float)scale));
  TRY(ensure(h, h->stage_f, (size_t)(n * h->K)));
  unpack_to_float_kernel<<<cdiv(n * h->K, 256), 256, 0, h->stream>>>(h->stage_f.p, bt.Z, n, h->K, h->ld);
  KERNEL_OK(h);
  CU(cudaMemcpyAsync(z_out, h->stage_f.p, sizeof(float) * n * h->K, cudaMemcpyDeviceToHost, h->stream));
  h->d2h += sizeof(float) * n * h->K;
  return end_call(h, nullptr);
}

int cdae_data_loss(cdae_handle* h, uint64_t seed, double* out) {
  if (!h || !out) return set_error(CDAE_E_INVALID, "NULL argument");
  TRY(begin_call(h));
  if (!h->plan_valid) TRY(build_plan(h));
  TRY(ensure_scratch(h, h->plan_max_users, h->plan_max_slots));
  const int cnum = h->cfg.num_corruptions;
  const size_t per = (size_t)std::max<int64_t>(h->scratch_users, 1) * h->ld;
  for (const MiniBatch& p : h->plan) {
    for (int c = 0; c < cnum; ++c) {
      BatchDev bt = make_batch(h, h->plan_in.p + p.in0, p.n_in, h->plan_out.p + p.out0, p.n_out,
                               h->plan_uids.p + p.user0, p.n_users);
      const SampleArgs sa = make_sample_args(h, seed, 0x80000000u + (uint32_t)c);
      CU(cudaMemsetAsync(h->acc3.p, 0, sizeof(float) * per, h->stream));
      TRY(launch_gather(h, bt, &sa, false));
      TRY(launch_activate(h, bt, h->m.scale));
      TRY(launch_decode(h, bt, false, nullptr));
    }
  }
  cdae_epoch_stats_t st;
  TRY(end_call(h, &st));
  double v = st.loss_sum / (double)cnum;  // user_rets / num_corruptions_, cdae.hpp:98
  if (h->world > 1) {
    // sum of the per-rank partial losses (tiny): reuse the stats staging as a 2-float all-reduce
    // (hi/lo split keeps double-ish precision)
    float hl[2] = {(float)v, (float)(v - (double)(float)v)};
    TRY(ensure(h, h->stage_f, 2));
    CU(cudaMemcpyAsync(h->stage_f.p, hl, sizeof(hl), cudaMemcpyHostToDevice, h->stream));
    NC(g_nccl.AllReduce(h->stage_f.p, h->stage_f.p, 2, kNcclFloat, kNcclSum, (ncclComm_t)h->comm, h->stream));
    CU(cudaMemcpyAsync(hl, h->stage_f.p, sizeof(hl), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    v = (double)hl[0] + (double)hl[1];
  }
  *out = v;
  return 0;
}

int cdae_penalty_loss(cdae_handle* h, double* out) {
  if (!h || !out) return set_error(CDAE_E_INVALID, "NULL argument");
  CU(cudaSetDevice(h->cfg.device));
  struct { const float* p; int64_t n; } blocks[5] = {
      {h->m.W, h->I * h->ld}, {h->m.V, h->m.V ? h->I * h->ld : 0}, {h->m.Wu, h->m.Wu ? h->U * h->ld : 0},
      {h->m.b, h->ld}, {h->m.bp, h->I4}};
  TRY(ensure(h, h->stage_d, 2));
  CU(cudaMemsetAsync(h->stage_d.p, 0, 2 * sizeof(double), h->stream));
  for (auto& b : blocks) {
    if (!b.p || b.n == 0) continue;  // pad entries are exactly 0 and do not contribute
    if (h->world > 1 && b.p == h->m.Wu) continue;
    sumsq_kernel<<<std::min(cdiv(b.n, 256), h->sm_count * 8), 256, 0, h->stream>>>(b.p, b.n, h->stage_d.p);
    KERNEL_OK(h);
  }
  if (h->world > 1 && h->m.Wu) {
    // Wu rows are current only on their owning rank: each rank sums the squares of the rows it trains into
    // a second slot, the slots are summed across ranks (collective), the replicated item side counts once
    sumsq_owned_rows_kernel<<<std::min(cdiv(h->U * h->ld, 256), h->sm_count * 8), 256, 0, h->stream>>>(
        h->m.Wu, h->U, h->ld, h->batch_users, h->rank, h->world, h->stage_d.p + 1);
    KERNEL_OK(h);
    NC(g_nccl.AllReduce(h->stage_d.p + 1, h->stage_d.p + 1, 1, kNcclDouble, kNcclSum, (ncclComm_t)h->comm, h->stream));
    add_double_kernel<<<1, 1, 0, h->stream>>>(h->stage_d.p, h->stage_d.p + 1);
    KERNEL_OK(h);
  }
  double s = 0.;
  CU(cudaMemcpyAsync(&s, h->stage_d.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *out = 0.5 * h->cfg.lambda * s;  // cdae.hpp:104
  return 0;
}

int cdae_dist_unique_id(void* id128_out) {
  if (!id128_out) return set_error(CDAE_E_INVALID, "NULL argument");
  TRY(load_nccl());
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  memcpy(id128_out, &id, sizeof(id));
  return 0;
}

int cdae_dist_init(cdae_handle* h, int32_t rank, int32_t world, const void* nccl_unique_id) {
  if (!h || !nccl_unique_id || world < 1 || rank < 0 || rank >= world) return set_error(CDAE_E_INVALID, "bad rank/world");
  if (h->comm) return set_error(CDAE_E_STATE, "process group already initialised");
  TRY(load_nccl());
  CU(cudaSetDevice(h->cfg.device));
  ncclUniqueId id;
  memcpy(&id, nccl_unique_id, sizeof(id));
  ncclComm_t comm = nullptr;
  NC(g_nccl.CommInitRank(&comm, world, id, rank));
  h->comm = comm;
  h->rank = rank; h->world = world;
  h->plan_valid = false;
  return 0;
}

// Peer-memory mode, step 1: switch the gradients to two ping-pong buffers (the fused step zeroes the idle
// one), allocate the flag array and export CUDA IPC handles of {gradient buffers, item-side parameter
// buffer, flags} — 3 x 64 bytes in a 256-byte record.
static int p2p_prepare(cdae_handle* h) {
  if (h->p2p_my_flags) return 0;
  CU(cudaStreamSynchronize(h->stream));
  static const bool unfused = getenv("CDAE_B200_P2P_FUSED") && atoi(getenv("CDAE_B200_P2P_FUSED")) == 0;
  h->p2p_fused = !unfused;
  if (h->p2p_fused) {
    float* two = nullptr;
    CU(cudaMalloc(&two, 2 * sizeof(float) * h->grad_floats));
    CU(cudaMemset(two, 0, 2 * sizeof(float) * h->grad_floats));
    h->grad.release();
    h->grad.p = two;
    h->grad.cap = 2 * h->grad_floats;
    h->p2p_parity = 0;
    h->m.steps_slot = 0;
    h->m.direct_lambda = 1;
    point_item_side(h, h->grad.p);
  }
  CU(cudaMalloc(&h->p2p_my_flags, 64 * sizeof(uint32_t)));
  CU(cudaMemset(h->p2p_my_flags, 0, 64 * sizeof(uint32_t)));
  h->p2p_done = reinterpret_cast<unsigned int*>(h->p2p_my_flags + 32);   // block counter, not touched by peers
  return 0;
}

int cdae_dist_p2p_export(cdae_handle* h, void* out256) {
  if (!h || !out256) return set_error(CDAE_E_INVALID, "NULL argument");
  if (h->world < 2 || h->world > p2p::MAX_RANKS) return set_error(CDAE_E_STATE, "needs a process group of 2..%d ranks (cdae_dist_init first)", p2p::MAX_RANKS);
  CU(cudaSetDevice(h->cfg.device));
  TRY(p2p_prepare(h));
  cudaIpcMemHandle_t hg, hp, hf;
  CU(cudaIpcGetMemHandle(&hg, h->grad.p));
  CU(cudaIpcGetMemHandle(&hp, h->item_params.p));
  CU(cudaIpcGetMemHandle(&hf, h->p2p_my_flags));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memset(out256, 0, 256);
  memcpy(out256, &hg, 64);
  memcpy((char*)out256 + 64, &hp, 64);
  memcpy((char*)out256 + 128, &hf, 64);
  return 0;
}

int cdae_dist_p2p_open(cdae_handle* h, const void* handles) {
  if (!h || !handles) return set_error(CDAE_E_INVALID, "NULL argument");
  if (h->world < 2 || h->world > p2p::MAX_RANKS || !h->p2p_my_flags) return set_error(CDAE_E_STATE, "cdae_dist_p2p_export has not run on this rank");
  CU(cudaSetDevice(h->cfg.device));
  for (int r = 0; r < h->world; ++r) {
    if (r == h->rank) {
      h->p2p_bufs[r] = h->grad.p;
      h->p2p_params[r] = h->item_params.p;
      h->p2p_flags[r] = h->p2p_my_flags;
      continue;
    }
    cudaIpcMemHandle_t hg, hp, hf;
    memcpy(&hg, (const char*)handles + (size_t)r * 256, 64);
    memcpy(&hp, (const char*)handles + (size_t)r * 256 + 64, 64);
    memcpy(&hf, (const char*)handles + (size_t)r * 256 + 128, 64);
    void *pg = nullptr, *pp = nullptr, *pf = nullptr;
    CU(cudaIpcOpenMemHandle(&pg, hg, cudaIpcMemLazyEnablePeerAccess));
    CU(cudaIpcOpenMemHandle(&pp, hp, cudaIpcMemLazyEnablePeerAccess));
    CU(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
    h->p2p_bufs[r] = (float*)pg;
    h->p2p_params[r] = (float*)pp;
    h->p2p_flags[r] = (uint32_t*)pf;
    h->p2p_ipc = true;
  }
  h->p2p_epoch = 0;
  h->p2p_on = true;
  return 0;
}

int cdae_profile(cdae_handle* h, int32_t enable) {
  if (!h) return set_error(CDAE_E_INVALID, "handle is NULL");
  h->profiling = enable != 0;
  for (int c = 0; c < CDAE_K_COUNT; ++c) { h->prof_ms[c] = 0.; h->prof_n[c] = 0; }
  h->prof_used = 0;
  return 0;
}
int cdae_profile_get(cdae_handle* h, double* ms_out, int64_t* launches_out) {
  if (!h || !ms_out || !launches_out) return set_error(CDAE_E_INVALID, "NULL argument");
  for (int c = 0; c < CDAE_K_COUNT; ++c) { ms_out[c] = h->prof_ms[c]; launches_out[c] = h->prof_n[c]; }
  return 0;
}

// The L2 roofline of the sampled decode's access pattern (probe_kernels.cuh): `row_visits` uniformly
// random rows of a rows x ld table are read with 16-byte vector loads (mode & 1) and / or receive a
// 16-byte vector reduction into a second table (mode & 2), in the handle's own <G,NV> row geometry.
int cdae_probe_l2(cdae_handle* h, int64_t rows, int32_t mode, int64_t row_visits, int32_t reps,
                  double* gbs_out, double* ms_out) {
  if (!h || !gbs_out || rows <= 0 || row_visits <= 0 || reps <= 0 || mode < 1 || mode > 5)
    return set_error(CDAE_E_INVALID, "bad argument");
  CU(cudaSetDevice(h->cfg.device));
  const int ld = h->ld;
  struct Scratch {   // released on every return path (the CU macro returns early on a CUDA error)
    float *src = nullptr, *dst = nullptr, *sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Scratch() {
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
      cudaFree(src); cudaFree(dst); cudaFree(sink);
    }
  } sc;
  float *&src = sc.src, *&dst = sc.dst, *&sink = sc.sink;
  const size_t bytes = sizeof(float) * (size_t)rows * ld;
  CU(cudaMalloc(&src, bytes));
  CU(cudaMalloc(&dst, bytes));
  CU(cudaMalloc(&sink, 16));
  CU(cudaMemsetAsync(src, 0, bytes, h->stream));
  CU(cudaMemsetAsync(dst, 0, bytes, h->stream));
  const int rows_per_warp = 192;                    // one output chunk of the decode: <= 96 rows, two chunks' worth
  const uint32_t live_bytes = (uint32_t)((h->K * 4 + 15) / 16 * 16);   // modes 4, 5: bulk reductions of the real columns only
  const int n_warps = (int)((row_visits + rows_per_warp - 1) / rows_per_warp);
  const int grid = cdiv((int64_t)n_warps * 32, 256);
  cudaEvent_t &e0 = sc.e0, &e1 = sc.e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  int rc = 0;
  for (int it = 0; it < reps + 2 && rc == 0; ++it) {  // two warm-up launches bring both tables into L2
    if (it == 2) CU(cudaEventRecord(e0, h->stream));
#define CALL(G, NV)                                                                                              \
    do {                                                                                                           \
      if (mode == 1) l2_probe_kernel<G, NV, 1><<<grid, 256, 0, h->stream>>>(src, dst, rows, ld, rows_per_warp, n_warps, (uint32_t)it * 7919u, sink); \
      else if (mode == 2) l2_probe_kernel<G, NV, 2><<<grid, 256, 0, h->stream>>>(src, dst, rows, ld, rows_per_warp, n_warps, (uint32_t)it * 7919u, sink); \
      else if (mode == 3) l2_probe_kernel<G, NV, 3><<<grid, 256, 0, h->stream>>>(src, dst, rows, ld, rows_per_warp, n_warps, (uint32_t)it * 7919u, sink); \
      else if (mode == 4) l2_probe_bulk_kernel<G, NV, false><<<grid, 256, 0, h->stream>>>(src, dst, rows, rows_per_warp, n_warps, (uint32_t)it * 7919u, live_bytes, sink); \
      else l2_probe_bulk_kernel<G, NV, true><<<grid, 256, 0, h->stream>>>(src, dst, rows, rows_per_warp, n_warps, (uint32_t)it * 7919u, live_bytes, sink); \
    } while (0)
    DISPATCH_LD(ld, CALL);
#undef CALL
    if (cudaGetLastError() != cudaSuccess) rc = set_error(CDAE_E_CUDA, "probe launch failed");
  }
  float ms = 0.f;
  if (rc == 0) {
    CU(cudaEventRecord(e1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaEventElapsedTime(&ms, e0, e1));
  }
  if (rc) return rc;
  const double per = ms / reps;
  // modes 4 / 5 (bulk reductions, without / with the row loads) are reported in the same units as 2 / 3:
  // padded row bytes per visit, so the numbers compare as "rows per second"
  const int eq = mode == 4 ? 2 : mode == 5 ? 3 : mode;
  const double moved = (double)n_warps * rows_per_warp * ld * 4.0 * ((eq & 1) + ((eq >> 1) & 1));
  *gbs_out = moved / (per * 1e-3) / 1e9;
  if (ms_out) *ms_out = per;
  return 0;
}

// Measurement aid: the combine step alone (no user work in front of it, so no rank skew): `reps` times
// "fill the current gradient buffer with a non-zero pattern, combine_and_apply".  With cdae_profile on,
// cdae_profile_get's allreduce / apply classes then hold the pure cost of the step.  Collective.
int cdae_debug_combine(cdae_handle* h, int32_t reps) {
  if (!h || reps <= 0) return set_error(CDAE_E_INVALID, "bad argument");
  TRY(begin_call(h));
  if (!h->p2p_ts) {
    CU(cudaMalloc(&h->p2p_ts, 8 * sizeof(unsigned long long)));
    CU(cudaMemset(h->p2p_ts, 0, 8 * sizeof(unsigned long long)));
  }
  for (int r = 0; r < reps; ++r) {
    // 0x2f2f2f2f = 1.59e-10f: every element has a (tiny) gradient, so every parameter is rewritten and published
    CU(cudaMemsetAsync(h->m.gW, 0x2f, sizeof(float) * (size_t)(h->m.gcnt - h->m.gW), h->stream));
    TRY(combine_and_apply(h));
  }
  return end_call(h, nullptr);
}

// %globaltimer (ns) of the LAST fused combine kernel: [0] entry, [1] every peer's gradients complete, [2] block 0
// finished its slice (loads, optimiser step, stores, zeroing), [3] all blocks done, [4] every peer's stores are in
int cdae_debug_combine_times(cdae_handle* h, uint64_t* out5) {
  if (!h || !out5) return set_error(CDAE_E_INVALID, "NULL argument");
  if (!h->p2p_ts) return set_error(CDAE_E_STATE, "cdae_debug_combine has not run");
  CU(cudaStreamSynchronize(h->stream));
  unsigned long long t[8];
  CU(cudaMemcpy(t, h->p2p_ts, sizeof(t), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 5; ++i) out5[i] = t[i];
  return 0;
}

int cdae_host_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes < 0) return set_error(CDAE_E_INVALID, "bad argument");
  CU(cudaMallocHost(ptr, (size_t)std::max<int64_t>(bytes, 1)));
  return 0;
}
int cdae_host_free(void* ptr) {
  if (ptr) CU(cudaFreeHost(ptr));
  return 0;
}
int cdae_synchronize(cdae_handle* h) {
  if (!h) return set_error(CDAE_E_INVALID, "handle is NULL");
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}
int cdae_stream(cdae_handle* h, void** stream_out) {
  if (!h || !stream_out) return set_error(CDAE_E_INVALID, "NULL argument");
  *stream_out = (void*)h->stream;
  return 0;
}

}  // extern "C"

// Phase A of recommend: per-user candidate lists (TOPN_M best by score).
static int topn_candidates(cdae_handle* h, const float* Wd, const int32_t* users, int64_t n_users) {
  const size_t dyn = (size_t)TT_U * TOPN_M * (sizeof(float) + sizeof(int));
  {  // per launch: the attribute belongs to the CURRENT device (several handles / devices per process)
    CU(cudaFuncSetAttribute(topn_tile_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  }
  ProfScope ps(h, CDAE_K_TOPN_EXACT);
  topn_tile_fp32_kernel<<<cdiv(n_users, TT_U), 256, dyn, h->stream>>>(
      h->topn_z.p, Wd, h->m.bp, h->I, h->ld, users, (int)n_users, h->row_ptr_d.p, h->col_d.p,
      h->cand_id.p, h->cand_s.p, h->cand_cnt.p);
  KERNEL_OK(h);
  return 0;
}

#include "topn_api.inl"
#include "fulldec_api.inl"
#include "dataset.inl"
#include "mc_nvls.inl"
#include "group.inl"
