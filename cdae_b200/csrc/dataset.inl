// cdae_b200/csrc/dataset.inl — host side of the DATA path in front of the hot path (SURVEY.md §8f
// N2, N3), included at the end of api.cu.  Pure C++ (no device work): the text file the reference
// reads with Data::load(..., RECSYS, ...) goes straight to the CSR cdae_create() takes, skipping the
// reference's vector<Instance> of nested vectors and its hash of hashes (data-inl.hpp:45-64,
// instance.hpp:94-225, recsys_model_base.hpp:29-34), a binary cache of the loaded (and split) data set — the
// counterpart of Data::save / Data::load — and a versioned binary checkpoint of the model (the reference has none).
#include <fstream>
#include <memory>
#include <string_view>
#include <unordered_map>

struct cdae_dataset {
  std::vector<std::string> raw[2];       // FeatureGroupInfo::raw_str_map_ of the user / item group
  std::vector<int32_t> pu, pi;           // instances in file order (duplicates kept, like instances_)
  std::vector<int64_t> rp[3];            // CSR row pointers: 0 = all, 1 = train, 2 = test
  std::vector<int32_t> col[3];
  bool split_done = false;
};

namespace {
// rows = sorted unique items per user (the hash of hashes collapses duplicate pairs the same way)
void build_csr(int64_t U, const std::vector<int32_t>& u, const std::vector<int32_t>& it, const std::vector<uint8_t>* sel,
               uint8_t want, std::vector<int64_t>* rp, std::vector<int32_t>* col) {
  rp->assign((size_t)U + 1, 0);
  for (size_t k = 0; k < u.size(); ++k)
    if (!sel || (*sel)[k] == want) ++(*rp)[(size_t)u[k] + 1];
  for (int64_t x = 0; x < U; ++x) (*rp)[(size_t)x + 1] += (*rp)[(size_t)x];
  std::vector<int32_t> tmp((size_t)(*rp)[(size_t)U]);
  std::vector<int64_t> cur(rp->begin(), rp->end() - 1);
  for (size_t k = 0; k < u.size(); ++k)
    if (!sel || (*sel)[k] == want) tmp[(size_t)cur[(size_t)u[k]]++] = it[k];
  col->clear();
  col->reserve(tmp.size());
  std::vector<int64_t> out_rp((size_t)U + 1, 0);
  for (int64_t x = 0; x < U; ++x) {
    int32_t* b = tmp.data() + (*rp)[(size_t)x];
    int32_t* e = tmp.data() + (*rp)[(size_t)x + 1];
    std::sort(b, e);
    e = std::unique(b, e);
    col->insert(col->end(), b, e);
    out_rp[(size_t)x + 1] = (int64_t)col->size();
  }
  rp->swap(out_rp);
}
}  // namespace

extern "C" {

static int dataset_load_pairs_impl(const char* path, const char* delimiters, int32_t skip_header, cdae_dataset** out) {
  const std::string delims = delimiters && *delimiters ? delimiters : " ";
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return set_error(CDAE_E_INVALID, "cannot open %s", path);
  const std::streamsize sz = f.tellg();
  f.seekg(0);
  std::string buf((size_t)sz, '\0');
  if (sz > 0 && !f.read(&buf[0], sz)) return set_error(CDAE_E_INVALID, "cannot read %s", path);
  bool is_delim[256] = {false};
  for (unsigned char c : delims) is_delim[c] = true;

  std::unique_ptr<cdae_dataset> holder(new cdae_dataset());
  cdae_dataset* d = holder.get();
  std::unordered_map<std::string_view, int32_t> idx[2];   // FeatureGroupInfo::idx_map_: ids in first-seen order
  idx[0].reserve(1 << 16);
  idx[1].reserve(1 << 16);
  std::vector<std::string_view> first[2];
  size_t line_num = 0;                                     // counts NON-EMPTY lines (file_line_reader-inl.hpp:12-19)
  const char* p = buf.data();
  const char* end = p + buf.size();
  while (p < end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* le = nl ? nl : end;
    const char* next = nl ? nl + 1 : end;
    if (le > p && le[-1] == '\r') --le;                    // getline on a CRLF file would keep it; a token never wants it
    if (le == p) { p = next; continue; }                   // empty line: skipped, not counted
    const bool header = skip_header && line_num == 0;
    ++line_num;
    if (!header) {
      // split_line (file_utils.hpp:15-25): any delimiter character separates, empty tokens are dropped
      std::string_view tok[3];
      int nt = 0;
      const char* q = p;
      while (q < le) {
        while (q < le && is_delim[(unsigned char)*q]) ++q;
        if (q == le) break;
        const char* t0 = q;
        while (q < le && !is_delim[(unsigned char)*q]) ++q;
        if (nt < 3) tok[nt] = std::string_view(t0, (size_t)(q - t0));
        ++nt;
      }
      if (nt != 2) {                                       // the app's parser CHECK_EQ(rets.size(), 2), yelp.cpp:62
        return set_error(CDAE_E_INVALID, "%s: line %zu has %d fields, expected 2 (the reference CHECK-aborts, yelp.cpp:62)",
                         path, line_num, nt);
      }
      int32_t id[2];
      for (int g = 0; g < 2; ++g) {
        auto it = idx[g].find(tok[g]);
        if (it == idx[g].end()) {
          id[g] = (int32_t)idx[g].size();                  // FeatureGroupInfo::get_index, instance-inl.hpp:22-37
          idx[g].emplace(tok[g], id[g]);
          first[g].push_back(tok[g]);
        } else {
          id[g] = it->second;
        }
      }
      d->pu.push_back(id[0]);
      d->pi.push_back(id[1]);
    }
    p = next;
  }
  for (int g = 0; g < 2; ++g) {
    d->raw[g].reserve(first[g].size());
    for (auto sv : first[g]) d->raw[g].emplace_back(sv);
  }
  build_csr((int64_t)d->raw[0].size(), d->pu, d->pi, nullptr, 0, &d->rp[0], &d->col[0]);
  *out = holder.release();
  return 0;
}

int cdae_dataset_load_pairs(const char* path, const char* delimiters, int32_t skip_header, cdae_dataset** out) {
  if (!path || !out) return set_error(CDAE_E_INVALID, "NULL argument");
  try {   // no exception may cross the C ABI
    return dataset_load_pairs_impl(path, delimiters, skip_header, out);
  } catch (const std::exception& e) {
    return set_error(CDAE_E_INVALID, "%s: %s", path, e.what());
  }
}

int cdae_dataset_info(const cdae_dataset* d, int64_t* users, int64_t* items, int64_t* instances) {
  if (!d) return set_error(CDAE_E_INVALID, "dataset is NULL");
  if (users) *users = (int64_t)d->raw[0].size();
  if (items) *items = (int64_t)d->raw[1].size();
  if (instances) *instances = (int64_t)d->pu.size();
  return 0;
}

static int dataset_split_impl(cdae_dataset* d, double test_ratio, uint64_t seed) {
  // Data::random_split_by_feature_group(train, test, 0, ratio), data-inl.hpp:231-272: every user's
  // INSTANCES are shuffled and the first floor(n * ratio) go to test.  The shuffle is a Fisher-Yates
  // over the user's instances in file order driven by Philox4x32(seed, {uid, k, 0, 0x5B117}) — the
  // reference uses its time-seeded mt19937_64, which no one can reproduce either.
  const int64_t U = (int64_t)d->raw[0].size();
  std::vector<int64_t> rp((size_t)U + 1, 0);
  for (int32_t u : d->pu) ++rp[(size_t)u + 1];
  for (int64_t u = 0; u < U; ++u) rp[(size_t)u + 1] += rp[(size_t)u];
  std::vector<int64_t> order(d->pu.size()), cur(rp.begin(), rp.end() - 1);
  for (size_t k = 0; k < d->pu.size(); ++k) order[(size_t)cur[(size_t)d->pu[k]]++] = (int64_t)k;
  std::vector<uint8_t> sel(d->pu.size(), 1);               // 1 = train, 2 = test
  for (int64_t u = 0; u < U; ++u) {
    int64_t* b = order.data() + rp[(size_t)u];
    const int64_t n = rp[(size_t)u + 1] - rp[(size_t)u];
    for (int64_t k = n - 1; k > 0; --k) {
      const Philox4 r = philox4x32(seed, (uint32_t)u, (uint32_t)k, (uint32_t)((uint64_t)k >> 32), 0x5B117u);
      const uint64_t w = ((uint64_t)r.x << 32) | r.y;
      const int64_t j = (int64_t)(((unsigned __int128)w * (unsigned __int128)(k + 1)) >> 64);
      std::swap(b[k], b[j]);
    }
    const int64_t n_test = (int64_t)((double)n * test_ratio);   // static_cast<size_t>(size * ratio), :252
    for (int64_t k = 0; k < n_test; ++k) sel[(size_t)b[k]] = 2;
  }
  build_csr(U, d->pu, d->pi, &sel, 1, &d->rp[1], &d->col[1]);
  build_csr(U, d->pu, d->pi, &sel, 2, &d->rp[2], &d->col[2]);
  d->split_done = true;
  return 0;
}

int cdae_dataset_split(cdae_dataset* d, double test_ratio, uint64_t seed) {
  if (!d || !(test_ratio >= 0. && test_ratio <= 1.)) return set_error(CDAE_E_INVALID, "bad argument");
  try {   // no exception may cross the C ABI
    return dataset_split_impl(d, test_ratio, seed);
  } catch (const std::exception& e) {
    d->split_done = false;
    return set_error(CDAE_E_INVALID, "split: %s", e.what());
  }
}

int cdae_dataset_nnz(const cdae_dataset* d, int32_t which, int64_t* nnz) {
  if (!d || !nnz || which < 0 || which > 2) return set_error(CDAE_E_INVALID, "bad argument");
  if (which > 0 && !d->split_done) return set_error(CDAE_E_STATE, "cdae_dataset_split has not run");
  *nnz = (int64_t)d->col[which].size();
  return 0;
}

int cdae_dataset_csr(const cdae_dataset* d, int32_t which, int64_t* row_ptr, int32_t* col) {
  if (!d || !row_ptr || which < 0 || which > 2) return set_error(CDAE_E_INVALID, "bad argument");
  if (which > 0 && !d->split_done) return set_error(CDAE_E_STATE, "cdae_dataset_split has not run");
  memcpy(row_ptr, d->rp[which].data(), sizeof(int64_t) * d->rp[which].size());
  if (col && !d->col[which].empty()) memcpy(col, d->col[which].data(), sizeof(int32_t) * d->col[which].size());
  return 0;
}

int cdae_dataset_raw_id(const cdae_dataset* d, int32_t group, int64_t idx, const char** out) {
  if (!d || !out || group < 0 || group > 1 || idx < 0 || idx >= (int64_t)d->raw[group].size())
    return set_error(CDAE_E_INVALID, "bad argument");
  *out = d->raw[group][(size_t)idx].c_str();
  return 0;
}

int cdae_dataset_free(cdae_dataset* d) {
  delete d;
  return 0;
}

// ---- data set cache (version 1) — the counterpart of Data::save / Data::load (data.hpp:25-33, 52-60;
// io/serialize.hpp:16-46: gzip over a boost binary archive of the instance vector and the feature-group
// tables).  That byte format cannot be reproduced without Boost, so the cache has its own:
//   "CDAEDS01" | u32 version | per group (users, items): u64 count, then u32 length + bytes per raw id |
//   u64 instances, i32 user ids, i32 item ids (file order, duplicates kept) | u8 split_done |
//   if split: per part (train, test): u64 nnz, i64 row_ptr[U + 1], i32 col[nnz]
// The CSR of all pairs is rebuilt on load (it is a function of the instances); the split is stored, so a CPU
// baseline run and a GPU run can share one split without sharing a seed convention.
static const char kDsMagic[8] = {'C', 'D', 'A', 'E', 'D', 'S', '0', '1'};
static const uint32_t kDsVersion = 1;

int cdae_dataset_save(const cdae_dataset* d, const char* path) {
  if (!d || !path) return set_error(CDAE_E_INVALID, "NULL argument");
  FILE* f = fopen(path, "wb");
  if (!f) return set_error(CDAE_E_INVALID, "cannot open %s for writing", path);
  bool ok = fwrite(kDsMagic, 1, 8, f) == 8 && fwrite(&kDsVersion, sizeof(kDsVersion), 1, f) == 1;
  for (int g = 0; g < 2 && ok; ++g) {
    const uint64_t n = d->raw[g].size();
    ok = fwrite(&n, sizeof(n), 1, f) == 1;
    for (size_t k = 0; k < d->raw[g].size() && ok; ++k) {
      const uint32_t len = (uint32_t)d->raw[g][k].size();
      ok = fwrite(&len, sizeof(len), 1, f) == 1 && (len == 0 || fwrite(d->raw[g][k].data(), 1, len, f) == len);
    }
  }
  const uint64_t n_inst = d->pu.size();
  ok = ok && fwrite(&n_inst, sizeof(n_inst), 1, f) == 1;
  if (n_inst) ok = ok && fwrite(d->pu.data(), sizeof(int32_t), n_inst, f) == n_inst && fwrite(d->pi.data(), sizeof(int32_t), n_inst, f) == n_inst;
  const uint8_t split = d->split_done ? 1 : 0;
  ok = ok && fwrite(&split, 1, 1, f) == 1;
  for (int w = 1; w <= 2 && ok && split; ++w) {
    const uint64_t nnz = d->col[w].size();
    ok = fwrite(&nnz, sizeof(nnz), 1, f) == 1 && fwrite(d->rp[w].data(), sizeof(int64_t), d->rp[w].size(), f) == d->rp[w].size() &&
         (nnz == 0 || fwrite(d->col[w].data(), sizeof(int32_t), nnz, f) == nnz);
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) return set_error(CDAE_E_INVALID, "short write to %s", path);
  return 0;
}

static int dataset_load_impl(const char* path, FILE* f, cdae_dataset* d) {
  auto fail = [&](const char* what) { return set_error(CDAE_E_INVALID, "%s: %s", path, what); };
  // sizes in the file are checked against what is left of it before anything is allocated
  fseek(f, 0, SEEK_END);
  const long file_size = ftell(f);
  fseek(f, 0, SEEK_SET);
  auto left = [&]() -> uint64_t { return (uint64_t)(file_size - ftell(f)); };
  char magic[8];
  uint32_t version = 0;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kDsMagic, 8) != 0) return fail("not a cdae_b200 data set cache");
  if (fread(&version, sizeof(version), 1, f) != 1 || version != kDsVersion) return fail("unsupported cache version");
  for (int g = 0; g < 2; ++g) {
    uint64_t n = 0;
    if (fread(&n, sizeof(n), 1, f) != 1 || n > left() / sizeof(uint32_t) || n >= 0x7fffffffull) return fail("truncated (id table)");
    d->raw[g].resize((size_t)n);
    for (uint64_t k = 0; k < n; ++k) {
      uint32_t len = 0;
      if (fread(&len, sizeof(len), 1, f) != 1 || len > left()) return fail("truncated (raw id)");
      d->raw[g][(size_t)k].resize(len);
      if (len && fread(&d->raw[g][(size_t)k][0], 1, len, f) != len) return fail("truncated (raw id)");
    }
  }
  const int64_t U = (int64_t)d->raw[0].size(), I = (int64_t)d->raw[1].size();
  uint64_t n_inst = 0;
  if (fread(&n_inst, sizeof(n_inst), 1, f) != 1 || n_inst > left() / (2 * sizeof(int32_t))) return fail("truncated (instances)");
  d->pu.resize((size_t)n_inst);
  d->pi.resize((size_t)n_inst);
  if (n_inst && (fread(d->pu.data(), sizeof(int32_t), n_inst, f) != n_inst || fread(d->pi.data(), sizeof(int32_t), n_inst, f) != n_inst))
    return fail("truncated (instances)");
  for (uint64_t k = 0; k < n_inst; ++k)
    if (d->pu[(size_t)k] < 0 || d->pu[(size_t)k] >= U || d->pi[(size_t)k] < 0 || d->pi[(size_t)k] >= I) return fail("instance id out of range");
  uint8_t split = 0;
  if (fread(&split, 1, 1, f) != 1 || split > 1) return fail("truncated (split flag)");
  for (int w = 1; w <= 2 && split; ++w) {
    uint64_t nnz = 0;
    if (fread(&nnz, sizeof(nnz), 1, f) != 1 || nnz > left() / sizeof(int32_t) || (uint64_t)(U + 1) > left() / sizeof(int64_t))
      return fail("truncated (split part)");
    d->rp[w].resize((size_t)U + 1);
    d->col[w].resize((size_t)nnz);
    if (fread(d->rp[w].data(), sizeof(int64_t), (size_t)U + 1, f) != (size_t)U + 1 ||
        (nnz && fread(d->col[w].data(), sizeof(int32_t), nnz, f) != nnz))
      return fail("truncated (split part)");
    // the same invariants cdae_create checks: monotone row_ptr ending at nnz, ids in range, rows strictly ascending
    if (d->rp[w][0] != 0 || d->rp[w][(size_t)U] != (int64_t)nnz) return fail("split part: row_ptr does not match nnz");
    for (int64_t u = 0; u < U; ++u) {
      if (d->rp[w][(size_t)u + 1] < d->rp[w][(size_t)u]) return fail("split part: row_ptr not monotone");
      for (int64_t s0 = d->rp[w][(size_t)u]; s0 < d->rp[w][(size_t)u + 1]; ++s0) {
        const int32_t c = d->col[w][(size_t)s0];
        if (c < 0 || c >= I || (s0 > d->rp[w][(size_t)u] && d->col[w][(size_t)s0 - 1] >= c)) return fail("split part: bad row");
      }
    }
  }
  d->split_done = split != 0;
  build_csr(U, d->pu, d->pi, nullptr, 0, &d->rp[0], &d->col[0]);
  return 0;
}

int cdae_dataset_load(const char* path, cdae_dataset** out) {
  if (!path || !out) return set_error(CDAE_E_INVALID, "NULL argument");
  FILE* f = fopen(path, "rb");
  if (!f) return set_error(CDAE_E_INVALID, "cannot open %s", path);
  cdae_dataset* d = nullptr;
  int rc;
  try {   // no exception may cross the C ABI (a damaged file can still ask for more memory than there is)
    d = new cdae_dataset();
    rc = dataset_load_impl(path, f, d);
  } catch (const std::exception& e) {
    rc = set_error(CDAE_E_INVALID, "%s: %s", path, e.what());
  }
  fclose(f);
  if (rc != 0) {
    delete d;
    return rc;
  }
  *out = d;
  return 0;
}

// ---- model checkpoint (version 2):
//   "CDAEB200" | u32 version | config, field by field at fixed widths (4 x f64, 15 x i32) | i64 U, I, K |
//   per present block: i64 id, rows, cols, rows*cols doubles | i64 -1, 0, 0
// (version 1 wrote the raw cdae_config_t, whose layout depends on the compiler's padding.)
static const char kCkptMagic[8] = {'C', 'D', 'A', 'E', 'B', '2', '0', '0'};
static const uint32_t kCkptVersion = 2;

static void ckpt_cfg_pack(const cdae_config_t& c, double d[4], int32_t i[15]) {
  d[0] = c.lambda; d[1] = c.learn_rate; d[2] = c.corruption_ratio; d[3] = c.beta;
  const int32_t v[15] = {c.loss_type, c.num_dim, c.num_neg, c.num_corruptions, c.using_adagrad, c.asymmetric,
                         c.user_factor, c.linear, c.scaled, c.linear_function, c.tanh_act, c.batch_users,
                         c.full_decode, 0, 0};
  memcpy(i, v, sizeof(v));
}

// In a process group cdae_save is COLLECTIVE: the user-private blocks (Wu, Uu and their accumulators)
// are assembled from their owners (cdae_get_param all-reduces them), so every rank must call it; only
// rank 0 touches `path`.
int cdae_save(cdae_handle* h, const char* path) {
  if (!h || !path) return set_error(CDAE_E_INVALID, "NULL argument");
  const bool writer = h->rank == 0;
  FILE* f = writer ? fopen(path, "wb") : nullptr;
  int open_failed = (writer && !f) ? 1 : 0;
  // every rank must take the same number of collective steps: the writer reports "cannot open" after them
  double cd[4];
  int32_t ci[15];
  ckpt_cfg_pack(h->cfg, cd, ci);
  const int64_t dims[3] = {h->U, h->I, h->K};
  bool ok = true;
  if (f)
    ok = fwrite(kCkptMagic, 1, 8, f) == 8 && fwrite(&kCkptVersion, 4, 1, f) == 1 && fwrite(cd, sizeof(cd), 1, f) == 1 &&
         fwrite(ci, sizeof(ci), 1, f) == 1 && fwrite(dims, sizeof(dims), 1, f) == 1;
  std::vector<double> buf;
  int rc = 0;
  for (int which = 0; which < CDAE_P_COUNT; ++which) {
    int64_t r = 0, c = 0;
    if (cdae_param_shape(h, which, &r, &c) != 0) { ok = false; break; }
    const int64_t n = r * c;
    if (n == 0) continue;
    buf.resize((size_t)n);
    rc = cdae_get_param(h, which, buf.data(), n);
    if (rc != 0) break;
    if (f && ok) {
      const int64_t hdr[3] = {which, r, c};
      ok = fwrite(hdr, sizeof(hdr), 1, f) == 1 && fwrite(buf.data(), sizeof(double), (size_t)n, f) == (size_t)n;
    }
  }
  if (f) {
    const int64_t tail[3] = {-1, 0, 0};
    ok = ok && fwrite(tail, sizeof(tail), 1, f) == 1;
    if (fclose(f) != 0) ok = false;
  }
  if (rc != 0) return rc;
  if (open_failed) return set_error(CDAE_E_INVALID, "cannot open %s for writing", path);
  return ok ? 0 : set_error(CDAE_E_INVALID, "short write to %s", path);
}

// Restores every parameter block (incl. AdaGrad state).  The hyper-parameters of the HANDLE stay in
// force: the structural ones (shape, asymmetric, user_factor, linear_function) must match the file, a
// difference in the others (lambda, learn_rate, ...) is legal — resuming with a new learning rate is
// the usual reason — and is reported through cdae_last_error() with return code 0.  In a process
// group every rank loads the same file (each keeps the rows it owns up to date).
int cdae_load(cdae_handle* h, const char* path) {
  if (!h || !path) return set_error(CDAE_E_INVALID, "NULL argument");
  FILE* f = fopen(path, "rb");
  if (!f) return set_error(CDAE_E_INVALID, "cannot open %s", path);
  char magic[8];
  uint32_t version = 0;
  double cd[4], hd[4];
  int32_t ci[15], hi[15];
  int64_t dims[3];
  int rc = 0;
  ckpt_cfg_pack(h->cfg, hd, hi);
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kCkptMagic, 8) != 0 || fread(&version, 4, 1, f) != 1 || version != kCkptVersion ||
      fread(cd, sizeof(cd), 1, f) != 1 || fread(ci, sizeof(ci), 1, f) != 1 || fread(dims, sizeof(dims), 1, f) != 1)
    rc = set_error(CDAE_E_INVALID, "%s is not a cdae_b200 checkpoint (version %u)", path, kCkptVersion);
  else if (dims[0] != h->U || dims[1] != h->I || dims[2] != h->K || ci[5] != hi[5] || ci[6] != hi[6] || ci[9] != hi[9])
    rc = set_error(CDAE_E_INVALID, "checkpoint shape %lld x %lld, K=%lld does not match the model (%lld x %lld, K=%d)",
                   (long long)dims[0], (long long)dims[1], (long long)dims[2], (long long)h->U, (long long)h->I, h->K);
  std::vector<double> buf;
  while (rc == 0) {
    int64_t hdr[3];
    if (fread(hdr, sizeof(hdr), 1, f) != 1) { rc = set_error(CDAE_E_INVALID, "%s is truncated", path); break; }
    if (hdr[0] < 0) break;
    const int64_t n = hdr[1] * hdr[2];
    int64_t r = 0, c = 0;
    if (hdr[0] >= CDAE_P_COUNT || cdae_param_shape(h, (int)hdr[0], &r, &c) != 0 || r != hdr[1] || c != hdr[2]) {
      rc = set_error(CDAE_E_INVALID, "%s: block %lld has an unexpected shape", path, (long long)hdr[0]);
      break;
    }
    buf.resize((size_t)n);
    if (fread(buf.data(), sizeof(double), (size_t)n, f) != (size_t)n) { rc = set_error(CDAE_E_INVALID, "%s is truncated", path); break; }
    rc = cdae_set_param(h, (int)hdr[0], buf.data(), n);
  }
  fclose(f);
  if (rc == 0 && (memcmp(cd, hd, sizeof(cd)) != 0 || memcmp(ci, hi, sizeof(ci)) != 0))
    set_error(0, "note: %s was written with different hyper-parameters (lambda %g lr %g q %g beta %g loss %d); the handle's stay in force",
              path, cd[0], cd[1], cd[2], cd[3], ci[0]);
  return rc;
}

}  // extern "C"
