// cdae_b200/host/model/recsys/imf.hpp — libcf::IMF (matrix factorisation for implicit feedback) for the
// drop-in host tree.  Reference: src/model/recsys/imf.hpp:12-142; apps/yelp --method=MF (yelp.cpp:122-142).
//
//   prediction   r(u,i) = ub[u] + ib[i] + <uv[u], iv[i]>                                   (:124-126)
//   one step     g = loss'(r(u,i), label); every touched parameter p gets the raw gradient
//                g * (the other factor) + 2*lambda*p, AdaGrad-scaled (acc += grad^2; grad /= beta + sqrt(acc))
//                when enabled, then p -= learn_rate * grad.  All four gradients are formed from the values
//                BEFORE the step (:91-121); the bias pair is touched only with using_bias_term.
//   one epoch    users ascending; for each rated item one positive step, then num_neg steps on items drawn
//                by RecsysModelBase::sample_negative_item with the negative label (:72-86)
//
// Same class surface (IMFConfig, constructors, reset, train_one_iteration, train_one_instance,
// predict_user_item_rating, get_user_vecs / get_item_vecs, the protected members BPR's constructor
// assigns).  Storage is flat row-major std::vector<double>, not Eigen matrices: the step is a dozen fused
// loops over K contiguous doubles with no temporaries.  This model shares the sampled decode's access
// pattern (row gather, K-dot, loss gradient, row update — SURVEY.md 8f N4); it runs on the host here.
#ifndef _LIBCF_IMF_HPP_
#define _LIBCF_IMF_HPP_

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include <base/heap.hpp>
#include <base/utils.hpp>
#include <model/loss.hpp>
#include <model/recsys/recsys_model_base.hpp>

namespace libcf {

struct IMFConfig {
  IMFConfig() = default;
  double learn_rate = 0.1;
  double beta = 1.;
  double lambda = 0.01;
  LossType lt = SQUARE;
  PenaltyType pt = L2;
  size_t num_dim = 10;
  size_t num_neg = 5;
  bool using_bias_term = true;
  bool using_adagrad = true;
};

class IMF : public RecsysModelBase {
 public:
  IMF(const IMFConfig& mcfg) {
    configure("IMF", mcfg.learn_rate, mcfg.beta, mcfg.lambda, mcfg.lt, mcfg.pt, mcfg.num_dim, mcfg.num_neg,
              mcfg.using_bias_term, mcfg.using_adagrad);
  }

  IMF() = default;

  virtual void reset(const Data& data_set) {
    RecsysModelBase::reset(data_set);
    const size_t K = num_dim_;
    auto small_random = [](std::vector<double>& v, size_t n) {   // U[-1, 1] * 0.01 from rand(), as Eigen's Random() does
      v.resize(n);
      for (auto& x : v) x = (2.0 * std::rand() / (double)RAND_MAX - 1.0) * 0.01;
    };
    small_random(uv_, num_users_ * K);
    small_random(iv_, num_items_ * K);
    uv_ag_.assign(num_users_ * K, 1e-4);
    iv_ag_.assign(num_items_ * K, 1e-4);
    ub_.assign(num_users_, 0.);
    ib_.assign(num_items_, 0.);
    ub_ag_.assign(num_users_, 1e-4);
    ib_ag_.assign(num_items_, 1e-4);
    grad_u_.resize(K);
    grad_i_.resize(K);
  }

  virtual void train_one_iteration(const Data&) {
    for (size_t uid = 0; uid < num_users_; ++uid) {
      auto fit = user_rated_items_.find(uid);
      CHECK(fit != user_rated_items_.end());
      const auto& rated = fit->second;
      for (const auto& p : rated) {
        train_one_instance(uid, p.first, loss_->positive_label());
        for (size_t k = 0; k < num_neg_; ++k)
          train_one_instance(uid, sample_negative_item(rated), loss_->negative_label());
      }
    }
  }

  virtual void train_one_instance(size_t uid, size_t iid, double rui) {
    const size_t K = num_dim_;
    double* pu = &uv_[uid * K];
    double* pi = &iv_[iid * K];
    const double g = loss_->gradient(predict_user_item_rating(uid, iid), rui);
    const double reg = 2. * lambda_;
    for (size_t k = 0; k < K; ++k) {            // both raw gradients from the values before the step
      grad_u_[k] = g * pi[k] + reg * pu[k];
      grad_i_[k] = g * pu[k] + reg * pi[k];
    }
    double gub = g + reg * ub_[uid], gib = g + reg * ib_[iid];
    if (using_adagrad_) {
      if (using_bias_term_) {
        gub = scaled_by_history(gub, ub_ag_[uid]);
        gib = scaled_by_history(gib, ib_ag_[iid]);
      }
      double* au = &uv_ag_[uid * K];
      double* ai = &iv_ag_[iid * K];
      for (size_t k = 0; k < K; ++k) {
        grad_u_[k] = scaled_by_history(grad_u_[k], au[k]);
        grad_i_[k] = scaled_by_history(grad_i_[k], ai[k]);
      }
    }
    if (using_bias_term_) {
      ub_[uid] -= learn_rate_ * gub;
      ib_[iid] -= learn_rate_ * gib;
    }
    for (size_t k = 0; k < K; ++k) {
      pu[k] -= learn_rate_ * grad_u_[k];
      pi[k] -= learn_rate_ * grad_i_[k];
    }
  }

  double predict_user_item_rating(size_t uid, size_t iid) const {
    const size_t K = num_dim_;
    const double* pu = &uv_[uid * K];
    const double* pi = &iv_[iid * K];
    double dot = 0.;
    for (size_t k = 0; k < K; ++k) dot += pu[k] * pi[k];
    return ub_[uid] + ib_[iid] + dot;
  }

  DMatrix get_user_vecs() { return as_matrix(uv_, num_users_); }
  DMatrix get_item_vecs() { return as_matrix(iv_, num_items_); }

 protected:
  // shared by IMF and BPR (whose config structs have the same fields): hyper-parameters, loss, penalty, log line
  void configure(const char* name, double learn_rate, double beta, double lambda, LossType lt, PenaltyType pt,
                 size_t num_dim, size_t num_neg, bool bias, bool adagrad) {
    learn_rate_ = learn_rate; beta_ = beta; lambda_ = lambda;
    num_dim_ = num_dim; num_neg_ = num_neg;
    using_bias_term_ = bias; using_adagrad_ = adagrad;
    loss_ = Loss::create(lt);
    penalty_ = Penalty::create(pt);
    LOG(INFO) << name << " Model Configure: {lambda: " << lambda_ << "}, {Learn Rate: " << learn_rate_ << "}, {Beta: " << beta_
              << "}, {Loss: " << loss_->loss_type() << "}, {Penalty: " << penalty_->penalty_type() << "}, {Dim: " << num_dim_
              << "}, {BiasTerm: " << using_bias_term_ << "}, {Using AdaGrad: " << using_adagrad_ << "}, {Num Negative: "
              << num_neg_ << "}";
  }
  // AdaGrad: acc += grad^2; returns grad / (beta + sqrt(acc))
  double scaled_by_history(double grad, double& acc) const {
    acc += grad * grad;
    return grad / (beta_ + std::sqrt(acc));
  }
  DMatrix as_matrix(const std::vector<double>& flat, size_t rows) const {
    DMatrix m(rows, num_dim_);
    for (size_t r = 0; r < rows; ++r)
      for (size_t k = 0; k < num_dim_; ++k) m(r, k) = flat[r * num_dim_ + k];
    return m;
  }

  std::vector<double> uv_, iv_, uv_ag_, iv_ag_;   // [rows][num_dim_] row-major
  std::vector<double> ub_, ib_, ub_ag_, ib_ag_;
  std::vector<double> grad_u_, grad_i_;           // scratch of one step

  double learn_rate_ = 0.1;
  double beta_ = 1.;
  double lambda_ = 0;
  size_t num_dim_ = 0;
  bool using_bias_term_ = true;
  bool using_factor_term_ = true;
  bool using_adagrad_ = true;
  size_t num_neg_ = 0;
};

}  // namespace libcf

#endif  // _LIBCF_IMF_HPP_
