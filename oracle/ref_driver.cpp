// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, not product code.
//
// Compiles the reference's OWN sources (included where they lie under
// /root/reference/src, never copied) into oracle/_ref/libcdae_ref.so behind a
// small C ABI, so that (a) the CPU restatement in oracle/cdae_oracle.c can be
// checked against the real thing, (b) golden vectors can be generated from
// it (tests/golden/make_golden.py), and (c) bench.py --impl reference can time
// the reference's own single-threaded train_one_iteration.
//
// What is verbatim: src/model/recsys/cdae.hpp, recsys_model_base.hpp,
// model_base.hpp, loss.hpp, penalty.hpp, evaluation.hpp, base/data*.hpp,
// base/instance*.hpp, base/heap.hpp, base/random.hpp, base/parallel*.hpp,
// base/io/*.  What is NOT the reference: Eigen, Boost, glog and gflags are
// absent from this image, so the third-party names resolve to the stand-ins
// under /root/repo/compat (eager double loops; see compat/Eigen/Dense for the
// one numerical caveat: summation order inside .dot()).
//
// Access: CDAE keeps its parameters private with no accessor
// (cdae.hpp:428-453).  To inject / read them without editing the reference,
// this translation unit re-defines `private`/`protected` while including it.
// Negatives: RecsysModelBase::sample_negative_item is virtual
// (recsys_model_base.hpp:46); RefCDAE overrides it to replay an explicit
// queue, so RNG is out of the comparison.  With an empty queue it falls
// through to the reference's rand()%I rejection sampler.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <future>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <glog/logging.h>
#include <gflags/gflags.h>
#include <Eigen/Dense>
#include <Eigen/Sparse>
#include <boost/archive/binary_iarchive.hpp>
#include <boost/iostreams/filtering_stream.hpp>
#include <boost/tokenizer.hpp>

#define private public
#define protected public
#include <model/recsys/cdae.hpp>
#include <model/evaluation.hpp>
#undef private
#undef protected

namespace {

using libcf::CDAE;
using libcf::CDAEConfig;
using libcf::Data;
using libcf::DMatrix;
using libcf::DVector;

typedef std::unordered_map<size_t, double> ItemMap;

// TOPN_Evaluation's members are private by default-class access (no
// `public:` label, evaluation.hpp:95), which the `#define private` above
// cannot reach.  Explicit template instantiation may name private members, so
// this (standard-conforming) shim hands out the member pointer.
template <class Tag, typename Tag::type M>
struct PrivateMember {
  friend typename Tag::type get_member(Tag) { return M; }
};
struct EvalRecListTag {
  typedef std::vector<double> (libcf::TOPN_Evaluation<CDAE>::*type)(const std::vector<size_t>&,
                                                                  const ItemMap&) const;
  friend type get_member(EvalRecListTag);
};
template struct PrivateMember<EvalRecListTag, &libcf::TOPN_Evaluation<CDAE>::evaluate_rec_list>;

class RefCDAE : public CDAE {
 public:
  explicit RefCDAE(const CDAEConfig& c) : CDAE(c) {}
  size_t sample_negative_item(const ItemMap& user_map) const override {
    if (neg_queue.empty()) return CDAE::sample_negative_item(user_map);
    size_t j = neg_queue.front();
    neg_queue.pop_front();
    CHECK(!user_map.count(j)) << "explicit negative " << j << " is a positive of this user";
    CHECK_LT(j, num_items_);
    return j;
  }
  mutable std::deque<size_t> neg_queue;
};

struct RefHandle {
  Data train;
  std::unique_ptr<RefCDAE> model;
};

enum Which { P_W = 0, P_V, P_WU, P_B, P_BPRIME, P_UU, P_W_AG, P_V_AG, P_WU_AG, P_B_AG, P_BPRIME_AG, P_UU_AG };

// Returns a flat view (pointer, rows, cols) of one parameter block.
bool param_view(RefCDAE& m, int which, double** p, int64_t* rows, int64_t* cols) {
  DMatrix* M = nullptr;
  DVector* v = nullptr;
  switch (which) {
    case P_W: M = &m.W; break;
    case P_V: M = &m.V; break;
    case P_WU: M = &m.Wu; break;
    case P_UU: M = &m.Uu; break;
    case P_W_AG: M = &m.W_ag; break;
    case P_V_AG: M = &m.V_ag; break;
    case P_WU_AG: M = &m.Wu_ag; break;
    case P_UU_AG: M = &m.Uu_ag; break;
    case P_B: v = &m.b; break;
    case P_BPRIME: v = &m.b_prime; break;
    case P_B_AG: v = &m.b_ag; break;
    case P_BPRIME_AG: v = &m.b_prime_ag; break;
    default: return false;
  }
  if (M) {
    *p = M->data();
    *rows = M->rows();
    *cols = M->cols();
  } else {
    *p = v->data();
    *rows = v->size();
    *cols = 1;
  }
  return true;
}

ItemMap make_map(const int64_t* items, int64_t n) {
  ItemMap m;
  for (int64_t i = 0; i < n; ++i) m.emplace(static_cast<size_t>(items[i]), 1.0);
  return m;
}

}  // namespace

extern "C" {

// cfg_d = {lambda, learn_rate, corruption_ratio, beta}
// cfg_i = {loss_type, num_dim, num_neg, num_corruptions, using_adagrad, asymmetric,
//          user_factor, linear, scaled, linear_function, tanh}
// train_file: text, one "user item" pair per line (no header), loaded through
// the reference's own Data::load(RECSYS) with its own split_line parser.
void* ref_create(const double* cfg_d, const int32_t* cfg_i, const char* train_file) {
  CDAEConfig c;
  c.lambda = cfg_d[0];
  c.learn_rate = cfg_d[1];
  c.corruption_ratio = cfg_d[2];
  c.beta = cfg_d[3];
  c.lt = static_cast<libcf::LossType>(cfg_i[0]);
  c.num_dim = static_cast<size_t>(cfg_i[1]);
  c.num_neg = static_cast<size_t>(cfg_i[2]);
  c.num_corruptions = static_cast<size_t>(cfg_i[3]);
  c.using_adagrad = cfg_i[4] != 0;
  c.asymmetric = cfg_i[5] != 0;
  c.user_factor = cfg_i[6] != 0;
  c.linear = cfg_i[7] != 0;
  c.scaled = cfg_i[8] != 0;
  c.linear_function = cfg_i[9] != 0;
  c.tanh = cfg_i[10] != 0;

  auto* h = new RefHandle();
  auto parser = [](const std::string& line) {
    auto rets = libcf::split_line(line, " ");
    CHECK_EQ(rets.size(), 2u);
    return std::vector<std::string>{rets[0], rets[1], "1"};
  };
  std::string fname(train_file);
  h->train.load(fname, libcf::RECSYS, parser, false);
  h->model.reset(new RefCDAE(c));
  h->model->reset(h->train);
  return h;
}

void ref_destroy(void* hp) { delete static_cast<RefHandle*>(hp); }

int64_t ref_num_users(void* hp) { return static_cast<int64_t>(static_cast<RefHandle*>(hp)->model->num_users_); }
int64_t ref_num_items(void* hp) { return static_cast<int64_t>(static_cast<RefHandle*>(hp)->model->num_items_); }

void ref_seed(uint64_t mt_seed, uint32_t c_seed) {
  libcf::Random::seed(static_cast<size_t>(mt_seed));
  std::srand(c_seed);
}
void ref_set_num_thread(int32_t n) { FLAGS_num_thread = n; }

int ref_param_shape(void* hp, int which, int64_t* rows, int64_t* cols) {
  double* p;
  return param_view(*static_cast<RefHandle*>(hp)->model, which, &p, rows, cols) ? 0 : -1;
}
int ref_set_param(void* hp, int which, const double* src, int64_t n) {
  double* p;
  int64_t r, c;
  if (!param_view(*static_cast<RefHandle*>(hp)->model, which, &p, &r, &c)) return -1;
  if (r * c != n) return -2;
  std::memcpy(p, src, sizeof(double) * static_cast<size_t>(n));
  return 0;
}
int ref_get_param(void* hp, int which, double* dst, int64_t n) {
  double* p;
  int64_t r, c;
  if (!param_view(*static_cast<RefHandle*>(hp)->model, which, &p, &r, &c)) return -1;
  if (r * c != n) return -2;
  std::memcpy(dst, p, sizeof(double) * static_cast<size_t>(n));
  return 0;
}

// Number of train items of a user, and the order in which the reference's
// unordered_map iterates them (= the order the positives loop visits them).
int64_t ref_user_items(void* hp, int64_t uid, int64_t* out, int64_t cap) {
  auto& m = *static_cast<RefHandle*>(hp)->model;
  auto it = m.user_rated_items_.find(static_cast<size_t>(uid));
  if (it == m.user_rated_items_.end()) return -1;
  int64_t n = 0;
  for (auto& p : it->second) {
    if (out && n < cap) out[n] = static_cast<int64_t>(p.first);
    ++n;
  }
  return n;
}

void ref_push_negatives(void* hp, const int64_t* negs, int64_t n) {
  auto& q = static_cast<RefHandle*>(hp)->model->neg_queue;
  for (int64_t i = 0; i < n; ++i) q.push_back(static_cast<size_t>(negs[i]));
}
int64_t ref_pending_negatives(void* hp) {
  return static_cast<int64_t>(static_cast<RefHandle*>(hp)->model->neg_queue.size());
}

// cdae.hpp:198 with an explicit corrupted input set; output set = the user's
// train items exactly as train_one_iteration passes it (cdae.hpp:143).
void ref_train_one_user(void* hp, int64_t uid, const int64_t* in_items, int64_t n_in) {
  auto& m = *static_cast<RefHandle*>(hp)->model;
  auto it = m.user_rated_items_.find(static_cast<size_t>(uid));
  CHECK(it != m.user_rated_items_.end());
  ItemMap in = make_map(in_items, n_in);
  m.train_one_user_corruption(static_cast<size_t>(uid), in, it->second);
}

// cdae.hpp:136 verbatim (its own mt19937_64 masks; rand() negatives unless a
// queue was pushed).  Returns wall seconds of the call.
double ref_train_one_iteration(void* hp) {
  auto* h = static_cast<RefHandle*>(hp);
  auto t0 = std::chrono::steady_clock::now();
  h->model->train_one_iteration(h->train);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Same loop body as cdae.hpp:137-145 restricted to users [u0, u1) — used by
// bench.py --impl reference to time a bounded sample of the epoch.
double ref_train_user_range(void* hp, int64_t u0, int64_t u1) {
  auto& m = *static_cast<RefHandle*>(hp)->model;
  auto t0 = std::chrono::steady_clock::now();
  for (size_t uid = static_cast<size_t>(u0); uid < static_cast<size_t>(u1); ++uid) {
    auto fit = m.user_rated_items_.find(uid);
    CHECK(fit != m.user_rated_items_.end());
    auto& item_set = fit->second;
    for (size_t idx = 0; idx < m.num_corruptions_; ++idx) {
      auto corrupted = m.get_corrputed_input(item_set, m.corruption_ratio_);
      m.train_one_user_corruption(uid, corrupted, item_set);
    }
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// cdae.hpp:361 — returns the kept items (uses Random::uniform()).
int64_t ref_corrupt(void* hp, int64_t uid, double ratio, int64_t* out, int64_t cap) {
  auto& m = *static_cast<RefHandle*>(hp)->model;
  auto it = m.user_rated_items_.find(static_cast<size_t>(uid));
  CHECK(it != m.user_rated_items_.end());
  auto kept = m.get_corrputed_input(it->second, ratio);
  int64_t n = 0;
  for (auto& p : kept) {
    if (n < cap) out[n] = static_cast<int64_t>(p.first);
    ++n;
  }
  return n;
}

void ref_hidden(void* hp, int64_t uid, const int64_t* items, int64_t n, double scale, double* z_out) {
  auto& m = *static_cast<RefHandle*>(hp)->model;
  DVector z = m.get_hidden_values(static_cast<size_t>(uid), make_map(items, n), scale);
  for (Eigen::Index k = 0; k < z.size(); ++k) z_out[k] = z(k);
}

double ref_output(void* hp, const double* z, int64_t item) {
  auto& m = *static_cast<RefHandle*>(hp)->model;
  DVector zz(static_cast<Eigen::Index>(m.num_dim_));
  for (size_t k = 0; k < m.num_dim_; ++k) zz(static_cast<Eigen::Index>(k)) = z[k];
  return m.get_output_values(zz, static_cast<size_t>(item));
}

// cdae.hpp:162 with rated set = the user's train items (what TOPN_Evaluation passes).
void ref_recommend(void* hp, int64_t uid, int64_t topk, int64_t* out) {
  auto& m = *static_cast<RefHandle*>(hp)->model;
  auto it = m.user_rated_items_.find(static_cast<size_t>(uid));
  CHECK(it != m.user_rated_items_.end());
  auto ids = m.recommend(static_cast<size_t>(uid), static_cast<size_t>(topk), it->second);
  for (size_t i = 0; i < ids.size(); ++i) out[i] = static_cast<int64_t>(ids[i]);
}

void ref_user_representations(void* hp, double* out) {
  auto& m = *static_cast<RefHandle*>(hp)->model;
  DMatrix r = m.get_user_representations();
  std::memcpy(out, r.data(), sizeof(double) * static_cast<size_t>(r.size()));
}

double ref_data_loss(void* hp) {
  auto* h = static_cast<RefHandle*>(hp);
  return h->model->data_loss(h->train);
}
double ref_penalty_loss(void* hp) { return static_cast<RefHandle*>(hp)->model->penalty_loss(); }

double ref_loss_gradient(int32_t lt, double pred, double truth) {
  return libcf::Loss::create(static_cast<libcf::LossType>(lt))->gradient(pred, truth);
}
double ref_loss_evaluate(int32_t lt, double pred, double truth) {
  return libcf::Loss::create(static_cast<libcf::LossType>(lt))->evaluate(pred, truth);
}

// evaluation.hpp:183 — the 8 top-N metrics of one recommendation list.
void ref_evaluate_rec_list(const int64_t* list, int64_t n_list, const int64_t* test_items,
                           int64_t n_test, double* out8) {
  libcf::TOPN_Evaluation<CDAE> ev;
  std::vector<size_t> l(list, list + n_list);
  auto r = (ev.*get_member(EvalRecListTag()))(l, make_map(test_items, n_test));
  for (int i = 0; i < 8; ++i) out8[i] = r[static_cast<size_t>(i)];
}

// evaluation.hpp:113 — whole TOPN evaluation against a test file (same text
// format, loaded with the SAME DataInfo so indices agree); returns the 8
// averaged metrics parsed back from the reference's formatted string.
int ref_topn_evaluate(void* hp, const char* test_file, double* out8) {
  auto* h = static_cast<RefHandle*>(hp);
  Data test(h->train.get_data_info());
  // Re-use the train DataInfo: load() would append feature groups, so add the
  // instances by hand through the public line hook instead.
  std::ifstream f(test_file);
  if (!f) return -1;
  std::string line;
  auto info = h->train.get_data_info();
  while (std::getline(f, line)) {
    if (line.empty()) continue;
    auto rets = libcf::split_line(line, " ");
    CHECK_EQ(rets.size(), 2u);
    test.add_line_to_instance(line, [&](const std::string&) {
      libcf::Instance ins;
      ins.add_feat_group(info->feature_group_infos_[0], rets[0]);
      ins.add_feat_group(info->feature_group_infos_[1], rets[1]);
      ins.set_label(1.);
      return ins;
    });
  }
  auto ev = libcf::Evaluation<CDAE>::create(libcf::TOPN);  // public virtual in the base
  std::string s = ev->evaluate(*h->model, test, h->train);
  std::replace(s.begin(), s.end(), '|', ' ');
  std::istringstream is(s);
  for (int i = 0; i < 8; ++i) is >> out8[i];
  return 0;
}

}  // extern "C"
