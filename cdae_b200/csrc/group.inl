// cdae_b200/csrc/group.inl — single-process multi-GPU mode, included at the end of api.cu.
//
// SURVEY.md §5 / VERDICT r1: the reference's app is ONE process (Solver<CDAE>::train, solver-inl.hpp:19,53,55);
// to let it use every GPU of the box, a cdae_group owns one engine handle per device, drives them from
// one worker thread per GPU per call, and wires them exactly like a process group of `n` ranks: NCCL
// communicator (ncclCommInitRank from n threads), and the fused combine step over peer memory — here
// the peers' buffers are plain device pointers of the same process (cudaDeviceEnablePeerAccess) or, where
// the switch supports it, one multicast object shared without any file-descriptor passing.
// Every cdae_group_* call is the collective form of the cdae_* call of the same name.
#include <functional>
#include <thread>

struct cdae_group {
  std::vector<cdae_handle*> h;
  int n = 0;
};

namespace {
// run fn(rank) on one thread per GPU; first failure wins (its message becomes the caller's last error)
static int group_run(cdae_group* g, const std::function<int(int)>& fn) {
  std::vector<int> rc((size_t)g->n, 0);
  std::vector<std::string> msg((size_t)g->n);
  std::vector<std::thread> th;
  for (int r = 0; r < g->n; ++r)
    th.emplace_back([&, r]() {
      rc[(size_t)r] = fn(r);
      if (rc[(size_t)r] != 0) msg[(size_t)r] = cdae_last_error();
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < g->n; ++r)
    if (rc[(size_t)r] != 0) return set_error(rc[(size_t)r], "GPU %d of the group: %s", r, msg[(size_t)r].c_str());
  return 0;
}
static bool group_owns(const cdae_handle* h, int64_t u) {
  const int64_t lo = (u / h->batch_users) * h->batch_users, nb = std::min<int64_t>(h->batch_users, h->U - lo);
  return u >= lo + nb * h->rank / h->world && u < lo + nb * (h->rank + 1) / h->world;
}
}  // namespace

extern "C" {

int cdae_group_destroy(cdae_group* g) {
  if (!g) return 0;
  for (cdae_handle* h : g->h) cdae_destroy(h);
  delete g;
  return 0;
}

int cdae_group_create(const cdae_config_t* cfg, int64_t U, int64_t I, const int64_t* row_ptr, const int32_t* col,
                      const int32_t* devices, int32_t n, cdae_group** out) {
  if (!cfg || !out || n < 1 || n > p2p::MAX_RANKS) return set_error(CDAE_E_INVALID, "need 1..%d devices", p2p::MAX_RANKS);
  cdae_group* g = new cdae_group();
  g->n = n;
  for (int r = 0; r < n; ++r) {
    cdae_config_t c = *cfg;
    c.device = devices ? devices[r] : r;
    // the minibatch of the config is the GLOBAL one; by default every GPU keeps the single-GPU share
    if (c.batch_users <= 0 && !c.full_decode) c.batch_users = 16384 * n;
    cdae_handle* h = nullptr;
    const int rc = cdae_create(&c, U, I, row_ptr, col, &h);
    if (rc != 0) { cdae_group_destroy(g); return rc; }
    if (c.full_decode && cfg->batch_users <= 0) h->batch_users *= n;
    g->h.push_back(h);
  }
  if (n > 1) {
    char id[128];
    int rc = cdae_dist_unique_id(id);
    if (rc == 0) rc = group_run(g, [&](int r) { return cdae_dist_init(g->h[(size_t)r], r, n, id); });
    // the fused combine step over peer memory: NVLS where the switch offers multicast, else direct peer pointers
    static const bool want_p2p = !(getenv("CDAE_B200_P2P") && atoi(getenv("CDAE_B200_P2P")) == 0);
    static const bool want_mc = !(getenv("CDAE_B200_NVLS") && atoi(getenv("CDAE_B200_NVLS")) == 0);
    bool mc_ok = false;
    // (NVLS pays off where the peer-memory kernel is NVLink-bound: 8 GPUs; on 2-4 GPUs the peer kernel is faster)
    const bool force_mc = getenv("CDAE_B200_NVLS") && atoi(getenv("CDAE_B200_NVLS")) == 1;
    if (rc == 0 && want_p2p && want_mc && (n >= 8 || force_mc)) {
      int32_t fd = -1;
      if (cdae_dist_mc_create(g->h[0], &fd) == 0) {
        if (fd >= 0) close(fd);
        mc_ok = true;
        for (int r = 0; r < n && mc_ok; ++r) {
          if (r > 0) {                                  // same process: the object handle is shared directly
            g->h[(size_t)r]->mc_handle = g->h[0]->mc_handle;
            g->h[(size_t)r]->mc_size = g->h[0]->mc_size;
            g->h[(size_t)r]->mc_creator = true;         // "already holds the handle": attach only adds the device
            g->h[(size_t)r]->mc_shared = true;          // ... but does not own it
          }
          mc_ok = cdae_dist_mc_attach(g->h[(size_t)r], -1) == 0;
        }
        for (int r = 0; r < n && mc_ok; ++r) mc_ok = cdae_dist_mc_bind(g->h[(size_t)r]) == 0;
        if (!mc_ok) rc = set_error(CDAE_E_STATE, "multicast set-up failed half way: %s", cdae_last_error());
      }
    }
    if (rc == 0 && want_p2p && !mc_ok) {
      for (int a = 0; a < n && rc == 0; ++a) {
        if (cudaSetDevice(g->h[(size_t)a]->cfg.device) != cudaSuccess) rc = set_error(CDAE_E_CUDA, "cudaSetDevice");
        for (int b = 0; b < n && rc == 0; ++b) {
          if (a == b) continue;
          int can = 0;
          cudaDeviceCanAccessPeer(&can, g->h[(size_t)a]->cfg.device, g->h[(size_t)b]->cfg.device);
          if (!can) { rc = set_error(CDAE_E_STATE, "no peer access between devices %d and %d", g->h[(size_t)a]->cfg.device, g->h[(size_t)b]->cfg.device); break; }
          const cudaError_t e = cudaDeviceEnablePeerAccess(g->h[(size_t)b]->cfg.device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = set_error(CDAE_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
          cudaGetLastError();
        }
      }
      if (rc != 0 && rc == CDAE_E_STATE) rc = 0;          // no peer access: stay on NCCL all-reduce + replicated apply
      else if (rc == 0) {
        for (int r = 0; r < n && rc == 0; ++r) {
          cudaSetDevice(g->h[(size_t)r]->cfg.device);
          rc = p2p_prepare(g->h[(size_t)r]);
        }
        for (int r = 0; r < n && rc == 0; ++r) {
          cdae_handle* h = g->h[(size_t)r];
          for (int q = 0; q < n; ++q) {
            h->p2p_bufs[q] = g->h[(size_t)q]->grad.p;
            h->p2p_params[q] = g->h[(size_t)q]->item_params.p;
            h->p2p_flags[q] = g->h[(size_t)q]->p2p_my_flags;
          }
          h->p2p_epoch = 0;
          h->p2p_on = true;
        }
      }
    }
    if (rc != 0) { cdae_group_destroy(g); return rc; }
  }
  *out = g;
  return 0;
}

int cdae_group_size(cdae_group* g, int32_t* n) {
  if (!g || !n) return set_error(CDAE_E_INVALID, "NULL argument");
  *n = g->n;
  return 0;
}
/* the engine handle of one GPU, e.g. for replicated (item-side) reads from GPU 0 */
int cdae_group_handle(cdae_group* g, int32_t rank, cdae_handle** out) {
  if (!g || !out || rank < 0 || rank >= g->n) return set_error(CDAE_E_INVALID, "bad rank");
  *out = g->h[(size_t)rank];
  return 0;
}

int cdae_group_init_params(cdae_group* g, uint64_t seed) {
  if (!g) return set_error(CDAE_E_INVALID, "group is NULL");
  return group_run(g, [&](int r) { return cdae_init_params(g->h[(size_t)r], seed); });
}

static void group_sum_stats(const std::vector<cdae_epoch_stats_t>& s, cdae_epoch_stats_t* out) {
  if (!out) return;
  memset(out, 0, sizeof(*out));
  for (const auto& x : s) {
    out->user_steps += x.user_steps; out->outputs += x.outputs; out->inputs_kept += x.inputs_kept;
    out->loss_sum += x.loss_sum; out->kernel_launches += x.kernel_launches;
    out->h2d_bytes += x.h2d_bytes; out->d2h_bytes += x.d2h_bytes;
    out->device_ms = std::max(out->device_ms, x.device_ms);       // time of the slowest GPU
  }
}

int cdae_group_train_epoch(cdae_group* g, uint64_t seed, int64_t epoch, cdae_epoch_stats_t* stats) {
  if (!g) return set_error(CDAE_E_INVALID, "group is NULL");
  std::vector<cdae_epoch_stats_t> s((size_t)g->n);
  TRY(group_run(g, [&](int r) { return cdae_train_epoch(g->h[(size_t)r], seed, epoch, &s[(size_t)r]); }));
  group_sum_stats(s, stats);
  return 0;
}
int cdae_group_train_epoch_csr(cdae_group* g, const int64_t* row_ptr, const int32_t* col, uint64_t seed, int64_t epoch,
                               cdae_epoch_stats_t* stats) {
  if (!g) return set_error(CDAE_E_INVALID, "group is NULL");
  std::vector<cdae_epoch_stats_t> s((size_t)g->n);
  TRY(group_run(g, [&](int r) { return cdae_train_epoch_csr(g->h[(size_t)r], row_ptr, col, seed, epoch, &s[(size_t)r]); }));
  group_sum_stats(s, stats);
  return 0;
}
int cdae_group_train_users(cdae_group* g, const int64_t* uids, int64_t n, const uint8_t* keep_mask, const int32_t* negatives,
                           cdae_epoch_stats_t* stats) {
  if (!g) return set_error(CDAE_E_INVALID, "group is NULL");
  std::vector<cdae_epoch_stats_t> s((size_t)g->n);
  TRY(group_run(g, [&](int r) { return cdae_train_users(g->h[(size_t)r], uids, n, keep_mask, negatives, &s[(size_t)r]); }));
  group_sum_stats(s, stats);
  return 0;
}
int cdae_group_data_loss(cdae_group* g, uint64_t seed, double* out) {
  if (!g || !out) return set_error(CDAE_E_INVALID, "NULL argument");
  std::vector<double> v((size_t)g->n, 0.);
  TRY(group_run(g, [&](int r) { return cdae_data_loss(g->h[(size_t)r], seed, &v[(size_t)r]); }));
  *out = v[0];       // every rank holds the all-reduced value
  return 0;
}
int cdae_group_penalty_loss(cdae_group* g, double* out) {
  if (!g || !out) return set_error(CDAE_E_INVALID, "NULL argument");
  std::vector<double> v((size_t)g->n, 0.);
  TRY(group_run(g, [&](int r) { return cdae_penalty_loss(g->h[(size_t)r], &v[(size_t)r]); }));
  *out = v[0];
  return 0;
}
int cdae_group_topn_build(cdae_group* g, int32_t topk) {
  if (!g) return set_error(CDAE_E_INVALID, "group is NULL");
  return group_run(g, [&](int r) { return cdae_topn_build(g->h[(size_t)r], topk); });
}
/* thread-safe like cdae_topn_lookup: the list of `uid` comes from the GPU that trains that user (its Wu / Uu rows are
 * current only there) */
int cdae_group_topn_lookup(cdae_group* g, int64_t uid, int64_t* ids_out, float* scores_out) {
  if (!g) return set_error(CDAE_E_INVALID, "group is NULL");
  if (uid < 0 || uid >= g->h[0]->U) return set_error(CDAE_E_INVALID, "uid out of range");
  for (cdae_handle* h : g->h)
    if (group_owns(h, uid)) return cdae_topn_lookup(h, uid, ids_out, scores_out);
  return set_error(CDAE_E_STATE, "no owner for user %lld", (long long)uid);
}
/* cdae_encode for users owned by different GPUs: each GPU encodes the listed users it owns */
int cdae_group_encode(cdae_group* g, const int64_t* uids, int64_t n, const uint8_t* keep_mask, double scale, float* z_out) {
  if (!g || !uids || !z_out || n <= 0) return set_error(CDAE_E_INVALID, "NULL / empty argument");
  const int K = g->h[0]->K;
  std::vector<int64_t> off((size_t)n + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    if (uids[i] < 0 || uids[i] >= g->h[0]->U) return set_error(CDAE_E_INVALID, "uid out of range");
    off[(size_t)i + 1] = off[(size_t)i] + (g->h[0]->row_ptr_h[uids[i] + 1] - g->h[0]->row_ptr_h[uids[i]]);
  }
  return group_run(g, [&](int r) {
    cdae_handle* h = g->h[(size_t)r];
    std::vector<int64_t> mine, pos;
    std::vector<uint8_t> keep;
    for (int64_t i = 0; i < n; ++i)
      if (group_owns(h, uids[i])) {
        mine.push_back(uids[i]);
        pos.push_back(i);
        if (keep_mask) keep.insert(keep.end(), keep_mask + off[(size_t)i], keep_mask + off[(size_t)i + 1]);
      }
    if (mine.empty()) return 0;
    keep.push_back(0);
    std::vector<float> z(mine.size() * (size_t)K);
    const int rc = cdae_encode(h, mine.data(), (int64_t)mine.size(), keep_mask ? keep.data() : nullptr, scale, z.data());
    if (rc != 0) return rc;
    for (size_t j = 0; j < mine.size(); ++j) memcpy(z_out + pos[j] * K, z.data() + j * (size_t)K, sizeof(float) * (size_t)K);
    return 0;
  });
}
/* collective read-back: user-private blocks and (in fused peer-memory mode) accumulators are assembled from their owners */
int cdae_group_get_param(cdae_group* g, int which, double* dst, int64_t n) {
  if (!g) return set_error(CDAE_E_INVALID, "group is NULL");
  std::vector<std::vector<double>> scratch((size_t)g->n);
  return group_run(g, [&](int r) {
    if (r == 0) return cdae_get_param(g->h[0], which, dst, n);
    scratch[(size_t)r].resize((size_t)std::max<int64_t>(n, 1));
    return cdae_get_param(g->h[(size_t)r], which, scratch[(size_t)r].data(), n);
  });
}
int cdae_group_get_param_rows(cdae_group* g, int which, const int64_t* rows, int64_t n, double* dst) {
  if (!g) return set_error(CDAE_E_INVALID, "group is NULL");
  std::vector<std::vector<double>> scratch((size_t)g->n);
  return group_run(g, [&](int r) {
    if (r == 0) return cdae_get_param_rows(g->h[0], which, rows, n, dst);
    scratch[(size_t)r].resize((size_t)std::max<int64_t>(n, 1) * (size_t)g->h[0]->K);
    return cdae_get_param_rows(g->h[(size_t)r], which, rows, n, scratch[(size_t)r].data());
  });
}
int cdae_group_set_param(cdae_group* g, int which, const double* src, int64_t n) {
  if (!g) return set_error(CDAE_E_INVALID, "group is NULL");
  return group_run(g, [&](int r) { return cdae_set_param(g->h[(size_t)r], which, src, n); });
}
int cdae_group_save(cdae_group* g, const char* path) {
  if (!g) return set_error(CDAE_E_INVALID, "group is NULL");
  return group_run(g, [&](int r) { return cdae_save(g->h[(size_t)r], path); });
}
int cdae_group_load(cdae_group* g, const char* path) {
  if (!g) return set_error(CDAE_E_INVALID, "group is NULL");
  return group_run(g, [&](int r) { return cdae_load(g->h[(size_t)r], path); });
}

}  // extern "C"
