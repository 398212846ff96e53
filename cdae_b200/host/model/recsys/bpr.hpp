// cdae_b200/host/model/recsys/bpr.hpp — libcf::BPR (pairwise ranking on the IMF parameters) for the drop-in
// host tree.  Reference: src/model/recsys/bpr.hpp:12-107; apps/yelp --method=BPR (yelp.cpp:144-165).
//
//   one pair     d = r(u,i) - r(u,j); g = loss'(d, 1); raw gradients (values before the step, :68-77):
//                  uv[u]: g*(iv[i] - iv[j]) + 2*lambda*uv[u]      iv[i]: g*uv[u] + 2*lambda*iv[i]
//                  iv[j]: -g*uv[u] + 2*lambda*iv[j]               ib[i]: g + 2*lambda*ib[i]   ib[j]: -g + 2*lambda*ib[j]
//                AdaGrad / plain step as in IMF; the user bias cancels in d and is not touched.
//   one epoch    users ascending; for each rated item i, num_neg pairs (i, sampled j)           (:52-65)
#ifndef _LIBCF_BPR_HPP_
#define _LIBCF_BPR_HPP_

#include <algorithm>

#include <base/heap.hpp>
#include <base/utils.hpp>
#include <model/loss.hpp>
#include <model/recsys/imf.hpp>

namespace libcf {

struct BPRConfig {
  BPRConfig() = default;
  double learn_rate = 0.1;
  double beta = 1.;
  double lambda = 0.01;
  LossType lt = LOG;
  PenaltyType pt = L2;
  size_t num_dim = 10;
  size_t num_neg = 5;
  bool using_bias_term = true;
  bool using_adagrad = true;
};

class BPR : public IMF {
 public:
  BPR(const BPRConfig& mcfg) {
    configure("BPR", mcfg.learn_rate, mcfg.beta, mcfg.lambda, mcfg.lt, mcfg.pt, mcfg.num_dim, mcfg.num_neg,
              mcfg.using_bias_term, mcfg.using_adagrad);
  }

  void reset(const Data& data_set) {
    IMF::reset(data_set);
    grad_j_.resize(num_dim_);
  }

  virtual void train_one_iteration(const Data&) {
    for (size_t uid = 0; uid < num_users_; ++uid) {
      auto fit = user_rated_items_.find(uid);
      CHECK(fit != user_rated_items_.end());
      const auto& rated = fit->second;
      for (const auto& p : rated)
        for (size_t k = 0; k < num_neg_; ++k) train_one_pair(uid, p.first, sample_negative_item(rated), 1.);
    }
  }

  virtual void train_one_pair(size_t uid, size_t iid, size_t jid, double rui) {
    const size_t K = num_dim_;
    double* pu = &uv_[uid * K];
    double* pi = &iv_[iid * K];
    double* pj = &iv_[jid * K];
    const double g = loss_->gradient(predict_user_item_rating(uid, iid) - predict_user_item_rating(uid, jid), rui);
    const double reg = 2. * lambda_;
    for (size_t k = 0; k < K; ++k) {
      grad_u_[k] = g * (pi[k] - pj[k]) + reg * pu[k];
      grad_i_[k] = g * pu[k] + reg * pi[k];
      grad_j_[k] = -g * pu[k] + reg * pj[k];
    }
    double gib = g + reg * ib_[iid], gjb = -g + reg * ib_[jid];
    if (using_adagrad_) {
      if (using_bias_term_) {
        gib = scaled_by_history(gib, ib_ag_[iid]);
        gjb = scaled_by_history(gjb, ib_ag_[jid]);
      }
      double* au = &uv_ag_[uid * K];
      double* ai = &iv_ag_[iid * K];
      double* aj = &iv_ag_[jid * K];
      for (size_t k = 0; k < K; ++k) {
        grad_u_[k] = scaled_by_history(grad_u_[k], au[k]);
        grad_i_[k] = scaled_by_history(grad_i_[k], ai[k]);
        grad_j_[k] = scaled_by_history(grad_j_[k], aj[k]);
      }
    }
    if (using_bias_term_) {
      ib_[iid] -= learn_rate_ * gib;
      ib_[jid] -= learn_rate_ * gjb;
    }
    for (size_t k = 0; k < K; ++k) {
      pu[k] -= learn_rate_ * grad_u_[k];
      pi[k] -= learn_rate_ * grad_i_[k];
      pj[k] -= learn_rate_ * grad_j_[k];
    }
  }

 protected:
  std::vector<double> grad_j_;
};

}  // namespace libcf

#endif  // _LIBCF_BPR_HPP_
