// cdae_b200/host/model/recsys/cdae.hpp — libcf::CDAE on the B200 engine.
//
// Drop-in replacement for the reference's src/model/recsys/cdae.hpp: same include path, same
// include guard, same `CDAEConfig` fields and the same public member signatures of `class CDAE`
// (reference cdae.hpp:13-31,39,76,78,103,109,136,148,162,198,361,373,418), so that
// Solver<CDAE> (solver-inl.hpp:19,53,55), TOPN_Evaluation<CDAE> (evaluation.hpp:135,145) and
// apps/yelp/yelp.cpp compile UNCHANGED when this directory precedes the reference's src/ on the
// include path.  Every method is a thin caller of the C ABI in
// include/cdae_b200.h; no arithmetic of the path runs on the host, and there is no CPU fallback:
// a failing call is LOG(FATAL), the reference's own error convention (glog CHECK -> abort).
//
// What differs from the reference, by construction:
//  * train_one_iteration runs frozen minibatches of `batch_users` users (CDAEConfig::batch_users,
//    env CDAE_B200_BATCH_USERS; 1 = the reference's per-user online step);
//  * corruption masks / negatives inside train_one_iteration and data_loss come from the engine's
//    counter-based Philox streams (seeded once from rand(), so srand() controls them) instead of
//    Random::uniform() / rand() — the explicit-input members (train_one_user_corruption,
//    get_corrputed_input) still draw from the reference's generators on the host;
//  * parameters live on the device in fp32;
//  * recommend() serves lists from a table built for ALL users at once (pre_recommend(), the
//    hook TOPN_Evaluation calls first); `rated_item_set` must be the user's train row, which is
//    what every caller in the reference passes (evaluation.hpp:145).
#ifndef _LIBCF_CDAE_HPP_
#define _LIBCF_CDAE_HPP_

#include <base/random.hpp>
#include <base/mat.hpp>
#include <base/instance.hpp>
#include <base/data.hpp>
#include <base/parallel.hpp>
#include <model/recsys/recsys_model_base.hpp>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include <cdae_b200.h>

namespace libcf {

struct CDAEConfig {
  CDAEConfig() = default;
  // reference fields, same names / defaults (cdae.hpp:13-31)
  double lambda = 0.01;
  double learn_rate = 0.1;
  LossType lt = LOGISTIC;
  PenaltyType pt = L2;
  size_t num_dim = 10;
  bool using_adagrad = true;
  double corruption_ratio = 0.5;
  size_t num_corruptions = 1;
  bool asymmetric = false;
  bool user_factor = true;
  bool linear = false;
  size_t num_neg = 5;
  bool scaled = true;
  double beta = 0.;
  bool linear_function = false;
  bool tanh = false;
  // device options (no reference counterpart); 0 = engine default / environment
  size_t batch_users = 0;
  int device = -1;
  bool full_decode = false;   // decode against ALL items on tcgen05 (num_neg ignored); env CDAE_B200_FULL_DECODE=1
  int num_gpus = 1;           // > 1: one process drives that many GPUs (cdae_group_*, users sharded data-parallel,
                              // batch_users is then the GLOBAL minibatch); env CDAE_B200_GPUS
};

class CDAE : public RecsysModelBase {
 public:
  CDAE(const CDAEConfig& mcfg) : mcfg_(mcfg) {
    loss_ = Loss::create(mcfg.lt);
    penalty_ = Penalty::create(mcfg.pt);
    if (const char* e = std::getenv("CDAE_B200_BATCH_USERS")) mcfg_.batch_users = std::strtoull(e, nullptr, 10);
    if (const char* e = std::getenv("CDAE_B200_DEVICE")) mcfg_.device = std::atoi(e);
    if (const char* e = std::getenv("CDAE_B200_FULL_DECODE")) mcfg_.full_decode = std::atoi(e) != 0;
    if (const char* e = std::getenv("CDAE_B200_GPUS")) mcfg_.num_gpus = std::max(1, std::atoi(e));
    LOG(INFO) << "CDAE (cdae_b200, ABI " << cdae_abi_version() << ") Configure: \n"
        << "\t{lambda: " << mcfg_.lambda << "}, "
        << "{Loss: " << loss_->loss_type() << "}, "
        << "{Penalty: " << penalty_->penalty_type() << "}\n"
        << "\t{Dim: " << mcfg_.num_dim << "}, "
        << "{LearnRate: " << mcfg_.learn_rate << "}, "
        << "{Using AdaGrad: " << mcfg_.using_adagrad << "}\n"
        << "\t{Corruption Ratio: " << mcfg_.corruption_ratio << "}, "
        << "{Num Corruptions: " << mcfg_.num_corruptions << "}, "
        << "{Asymmetric: " << mcfg_.asymmetric << "}\n"
        << "\t{UserFactor: " << mcfg_.user_factor << "}, "
        << "{Linear: " << mcfg_.linear << "}, "
        << "{Num Negative: " << mcfg_.num_neg << "}, "
        << "{Scaled: " << mcfg_.scaled << "}\n"
        << "\t{Beta: " << mcfg_.beta << "}, "
        << "{LinearFunction: " << mcfg_.linear_function << "}, "
        << "{tanh: " << mcfg_.tanh << "}, "
        << "{BatchUsers: " << mcfg_.batch_users << "}, "
        << "{FullDecode: " << mcfg_.full_decode << "}";
  }

  CDAE() : CDAE(CDAEConfig()) {}

  // cdae.hpp:78-101 — sum over users of the loss of their positives under a fresh corruption
  double data_loss(const Data& data_set, size_t sample_size = 0) const {
    double v = 0.;
    const uint64_t s = seed_ + 0x9E3779B97F4A7C15ull * (uint64_t)(++st_->loss_calls);
    check(st_->g ? cdae_group_data_loss(st_->g, s, &v) : cdae_data_loss(handle(), s, &v));
    return v;
  }

  // cdae.hpp:103-107
  double penalty_loss() const {
    double v = 0.;
    check(st_->g ? cdae_group_penalty_loss(st_->g, &v) : cdae_penalty_loss(handle(), &v));
    return v;
  }

  // cdae.hpp:109-134 — Solver copy-constructs the model BEFORE calling reset (solver.hpp:17), so
  // the device state is created here and held by a ref-counted handle.
  void reset(const Data& data_set) {
    RecsysModelBase::reset(data_set);
    // user -> items hash of hashes (recsys_model_base.hpp:31) -> CSR with ascending rows
    std::vector<int64_t> row_ptr(num_users_ + 1, 0);
    std::vector<int32_t> col;
    for (size_t uid = 0; uid < num_users_; ++uid) {
      auto fit = user_rated_items_.find(uid);
      const size_t n = fit == user_rated_items_.end() ? 0 : fit->second.size();
      row_ptr[uid + 1] = row_ptr[uid] + (int64_t)n;
    }
    col.resize((size_t)row_ptr[num_users_]);
    for (size_t uid = 0; uid < num_users_; ++uid) {
      auto fit = user_rated_items_.find(uid);
      if (fit == user_rated_items_.end()) continue;
      int32_t* dst = col.data() + row_ptr[uid];
      size_t k = 0;
      for (auto& p : fit->second) dst[k++] = (int32_t)p.first;
      std::sort(dst, dst + k);
    }
    cdae_config_t c;
    check(cdae_config_default(&c));
    c.lambda = mcfg_.lambda; c.learn_rate = mcfg_.learn_rate;
    c.corruption_ratio = mcfg_.corruption_ratio; c.beta = mcfg_.beta;
    c.loss_type = (int32_t)mcfg_.lt;  // LossType numbering == enum cdae_loss (loss.hpp:10-18)
    c.num_dim = (int32_t)mcfg_.num_dim; c.num_neg = (int32_t)mcfg_.num_neg;
    c.num_corruptions = (int32_t)mcfg_.num_corruptions;
    c.using_adagrad = mcfg_.using_adagrad; c.asymmetric = mcfg_.asymmetric;
    c.user_factor = mcfg_.user_factor; c.linear = mcfg_.linear; c.scaled = mcfg_.scaled;
    c.linear_function = mcfg_.linear_function; c.tanh_act = mcfg_.tanh;
    c.batch_users = (int32_t)mcfg_.batch_users;
    c.device = mcfg_.device < 0 ? 0 : mcfg_.device;
    c.full_decode = mcfg_.full_decode ? 1 : 0;
    cdae_handle* h = nullptr;
    cdae_group* g = nullptr;
    if (mcfg_.num_gpus > 1) {
      // one process, several GPUs: devices c.device .. c.device + num_gpus - 1
      std::vector<int32_t> devs((size_t)mcfg_.num_gpus);
      for (int d = 0; d < mcfg_.num_gpus; ++d) devs[(size_t)d] = c.device + d;
      check(cdae_group_create(&c, (int64_t)num_users_, (int64_t)num_items_, row_ptr.data(), col.data(), devs.data(),
                              (int32_t)devs.size(), &g));
      check(cdae_group_handle(g, 0, &h));
    } else {
      check(cdae_create(&c, (int64_t)num_users_, (int64_t)num_items_, row_ptr.data(), col.data(), &h));
    }
    st_ = std::make_shared<State>();
    st_->h = h;
    st_->g = g;
    st_->row_ptr = std::move(row_ptr);
    st_->col = std::move(col);
    // the reference draws W, V, Wu with Eigen's Random(), i.e. from rand() (cdae.hpp:112-121):
    // take the engine's stream seed from rand() too, so srand() still decides the run
    seed_ = ((uint64_t)(uint32_t)rand() << 32) ^ (uint64_t)(uint32_t)rand();
    if (const char* e = std::getenv("CDAE_B200_SEED")) seed_ = std::strtoull(e, nullptr, 10);
    check(g ? cdae_group_init_params(g, seed_) : cdae_init_params(h, seed_));
    epoch_ = 0;
  }

  // cdae.hpp:136-146 — one pass over all users (the data is the set given to reset(), exactly
  // as in the reference, which also ignores `train_data` here)
  void train_one_iteration(const Data& train_data) {
    cdae_epoch_stats_t s;
    check(st_->g ? cdae_group_train_epoch(st_->g, seed_, epoch_++, &s) : cdae_train_epoch(handle(), seed_, epoch_++, &s));
    st_->last = s;
    st_->topn_k = 0;  // lists are stale now
  }

  // cdae.hpp:148-159
  DMatrix get_user_representations() {
    DMatrix user_vec(num_users_, mcfg_.num_dim);
    const size_t K = mcfg_.num_dim, step = 1 << 16;
    std::vector<int64_t> ids;
    std::vector<float> z;
    for (size_t a = 0; a < num_users_; a += step) {
      const size_t b = std::min(num_users_, a + step);
      ids.resize(b - a);
      z.resize((b - a) * K);
      for (size_t u = a; u < b; ++u) ids[u - a] = (int64_t)u;
      check(st_->g ? cdae_group_encode(st_->g, ids.data(), (int64_t)ids.size(), nullptr, 1.0, z.data())
                   : cdae_encode(handle(), ids.data(), (int64_t)ids.size(), nullptr, 1.0, z.data()));
      for (size_t u = a; u < b; ++u)
        for (size_t k = 0; k < K; ++k) user_vec(u, k) = z[(u - a) * K + k];
    }
    return user_vec;
  }

  // recsys_model_base.hpp:72 hook, called by TOPN_Evaluation before the per-user fan-out
  // (evaluation.hpp:135): score every user against every item ONCE on the device.
  virtual void pre_recommend() { build_lists(kDefaultTopk); }

  // cdae.hpp:162-196.  const and callable concurrently from ThreadPool workers
  // (evaluation.hpp:137-158): a lookup in the prebuilt table; (re)building is serialised.
  std::vector<size_t> recommend(size_t uid, size_t topk,
                                const std::unordered_map<size_t, double>& rated_item_set) const {
    CHECK_LT(uid, num_users_);
    const int64_t n_u = st_->row_ptr[uid + 1] - st_->row_ptr[uid];
    bool same = (int64_t)rated_item_set.size() == n_u;
    for (int64_t s = st_->row_ptr[uid]; same && s < st_->row_ptr[uid + 1]; ++s)
      same = rated_item_set.count((size_t)st_->col[s]) != 0;
    CHECK(same) << "cdae_b200: recommend() expects the user's train items as rated_item_set "
                   "(the set TOPN_Evaluation passes, evaluation.hpp:145)";
    {
      std::lock_guard<std::mutex> lk(st_->mu);
      if (st_->topn_k != (int)topk) build_lists((int)topk);
    }
    std::vector<int64_t> ids(topk);
    check(st_->g ? cdae_group_topn_lookup(st_->g, (int64_t)uid, ids.data(), nullptr)
                 : cdae_topn_lookup(handle(), (int64_t)uid, ids.data(), nullptr));
    std::vector<size_t> ret(topk);
    for (size_t i = 0; i < topk; ++i) ret[i] = (size_t)ids[i];
    return ret;
  }

  // cdae.hpp:198-358 — one user, explicit corrupted input; negatives drawn the reference's way
  // (sample_negative_item -> rand(), all before the update, cdae.hpp:217-220).
  void train_one_user_corruption(size_t uid,
                                 const std::unordered_map<size_t, double>& input_set,
                                 const std::unordered_map<size_t, double>& output_set) {
    const int64_t r0 = st_->row_ptr[uid], n_u = st_->row_ptr[uid + 1] - r0;
    CHECK_EQ((int64_t)output_set.size(), n_u) << "output_set must be the user's train items";
    std::vector<uint8_t> keep((size_t)n_u);
    size_t kept = 0;
    for (int64_t s = 0; s < n_u; ++s) {
      CHECK(output_set.count((size_t)st_->col[r0 + s])) << "output_set must be the user's train items";
      keep[(size_t)s] = input_set.count((size_t)st_->col[r0 + s]) ? 1 : 0;
      kept += keep[(size_t)s];
    }
    CHECK_EQ(kept, input_set.size()) << "input_set must be a subset of output_set";
    std::vector<int32_t> negs((size_t)n_u * mcfg_.num_neg);
    for (auto& j : negs) j = (int32_t)sample_negative_item(output_set);
    const int64_t u = (int64_t)uid;
    cdae_epoch_stats_t s;
    check(st_->g ? cdae_group_train_users(st_->g, &u, 1, keep.data(), negs.data(), &s)
                 : cdae_train_users(handle(), &u, 1, keep.data(), negs.data(), &s));
    st_->topn_k = 0;
  }

  // cdae.hpp:361-371 — host-side helper with the reference's generator and rule
  std::unordered_map<size_t, double> get_corrputed_input(const std::unordered_map<size_t, double>& input_set,
                                                         double corruption_ratio) const {
    std::unordered_map<size_t, double> rets;
    rets.reserve(static_cast<size_t>(input_set.size() * (1. - corruption_ratio)));
    for (auto& p : input_set)
      if (Random::uniform() > corruption_ratio) rets.insert(p);
    return rets;
  }

  // cdae.hpp:373-416 — item_set must be a subset of the user's train items (every call site in
  // the reference passes the train row or a corruption of it)
  DVector get_hidden_values(size_t uid, const std::unordered_map<size_t, double>& item_set,
                            double scale = 1.0) const {
    const int64_t r0 = st_->row_ptr[uid], n_u = st_->row_ptr[uid + 1] - r0;
    std::vector<uint8_t> keep((size_t)std::max<int64_t>(n_u, 1));
    size_t kept = 0;
    for (int64_t s = 0; s < n_u; ++s) {
      keep[(size_t)s] = item_set.count((size_t)st_->col[r0 + s]) ? 1 : 0;
      kept += keep[(size_t)s];
    }
    CHECK_EQ(kept, item_set.size()) << "cdae_b200: item_set must be a subset of the user's train items";
    std::vector<float> z(mcfg_.num_dim);
    const int64_t u = (int64_t)uid;
    check(st_->g ? cdae_group_encode(st_->g, &u, 1, keep.data(), scale, z.data())
                 : cdae_encode(handle(), &u, 1, keep.data(), scale, z.data()));
    DVector h2(mcfg_.num_dim);
    for (size_t k = 0; k < mcfg_.num_dim; ++k) h2(k) = z[k];
    return h2;
  }

  // cdae.hpp:418-426 — W'.row(iid) . z + b'(iid), W' = V if asymmetric else W
  double get_output_values(const DVector& z, size_t iid) const {
    std::vector<double> row(mcfg_.num_dim);
    double bp = 0.;
    const int64_t i = (int64_t)iid;
    check(cdae_get_param_rows(handle(), mcfg_.asymmetric ? CDAE_P_V : CDAE_P_W, &i, 1, row.data()));
    check(cdae_get_param_rows(handle(), CDAE_P_BPRIME, &i, 1, &bp));
    double ret = 0.;
    for (size_t k = 0; k < mcfg_.num_dim; ++k) ret += row[k] * z(k);
    return ret + bp;
  }

  // ---- additions (no reference counterpart) ----
  cdae_handle* engine() const { return handle(); }       // (GPU 0's handle in multi-GPU mode)
  cdae_group* engine_group() const { return st_ ? st_->g : nullptr; }
  // model checkpoint incl. AdaGrad state (the reference can only save Data, io/serialize.hpp:16-46)
  void save(const std::string& path) const { check(st_->g ? cdae_group_save(st_->g, path.c_str()) : cdae_save(handle(), path.c_str())); }
  void load(const std::string& path) {
    check(st_->g ? cdae_group_load(st_->g, path.c_str()) : cdae_load(handle(), path.c_str()));
    st_->topn_k = 0;
  }
  const cdae_epoch_stats_t& last_epoch_stats() const { return st_->last; }
  // TOPN_Evaluation::evaluate (evaluation.hpp:113-181) on the device, for callers that hold the
  // test set as CSR: out8 = P@1,P@5,P@10,R@1,R@5,R@10,MAP@5,MAP@10
  std::vector<double> evaluate_topn(const std::vector<int64_t>& test_row_ptr,
                                    const std::vector<int32_t>& test_col) {
    {
      std::lock_guard<std::mutex> lk(st_->mu);
      if (st_->topn_k != kDefaultTopk) build_lists(kDefaultTopk);
    }
    std::vector<double> out(8);
    int64_t n = 0;
    CHECK(!st_->g) << "cdae_b200: evaluate_topn is single-GPU; in multi-GPU mode TOPN_Evaluation reads the lists through recommend()";
    check(cdae_topn_evaluate(handle(), test_row_ptr.data(), test_col.data(), out.data(), &n));
    return out;
  }

 private:
  static constexpr int kDefaultTopk = 10;  // the list length TOPN_Evaluation asks for (evaluation.hpp:145)

  struct State {
    cdae_handle* h = nullptr;   // the engine (GPU 0's handle in multi-GPU mode: owned by g)
    cdae_group* g = nullptr;    // multi-GPU mode
    std::vector<int64_t> row_ptr;
    std::vector<int32_t> col;
    std::mutex mu;
    int topn_k = 0;
    uint64_t loss_calls = 0;
    cdae_epoch_stats_t last{};
    ~State() {
      if (g) cdae_group_destroy(g);
      else if (h) cdae_destroy(h);
    }
  };

  cdae_handle* handle() const {
    CHECK(st_ && st_->h) << "cdae_b200: reset(data) has not been called";
    return st_->h;
  }
  static void check(int rc) {
    if (rc != 0) LOG(FATAL) << "cdae_b200 error " << rc << ": " << cdae_last_error();
  }
  void build_lists(int topk) const {
    check(st_->g ? cdae_group_topn_build(st_->g, topk) : cdae_topn_build(handle(), topk));
    st_->topn_k = topk;
  }

  CDAEConfig mcfg_;
  std::shared_ptr<State> st_;  // shared by copies made after reset()
  uint64_t seed_ = 0;
  int64_t epoch_ = 0;
};

}  // namespace libcf

#endif  // _LIBCF_CDAE_HPP_
