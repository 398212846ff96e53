// cdae_b200/host/model/recsys/popularity.hpp — libcf::Popularity for the drop-in host tree.
//
// apps/yelp/yelp.cpp always evaluates this baseline first (yelp.cpp:109-113).  Same class surface as the
// reference's src/model/recsys/popularity.hpp:12-65 (default constructor, reset, no-op
// train_one_iteration, recommend), implemented over flat arrays: one counting pass over the data
// set, one ranking of the item ids, and recommend() walks that ranking skipping the user's rated
// items.  Ranking rule = the reference's: item count descending (sort_by_second_desc, utils.hpp:15-19;
// the order among equal counts is unspecified there — here the lower item id goes first).
#ifndef _LIBCF_POPULARITY_HPP_
#define _LIBCF_POPULARITY_HPP_

#include <algorithm>
#include <cstdint>
#include <numeric>
#include <unordered_map>
#include <vector>

#include <base/data.hpp>
#include <base/utils.hpp>
#include <model/recsys/recsys_model_base.hpp>

namespace libcf {

class Popularity : public RecsysModelBase {
 public:
  Popularity() : RecsysModelBase() { LOG(INFO) << "Popularity model"; }

  virtual void train_one_iteration(const Data&) {}   // nothing to learn

  void reset(const Data& data_set) {
    RecsysModelBase::reset(data_set);
    counts_.assign(num_items_, 0u);
    for (auto it = data_set.begin(); it != data_set.end(); ++it) {
      const size_t iid = it->get_feature_group_index(1, 0);
      CHECK_LT(iid, num_items_);
      ++counts_[iid];
    }
    ranking_.resize(num_items_);
    std::iota(ranking_.begin(), ranking_.end(), (size_t)0);
    std::stable_sort(ranking_.begin(), ranking_.end(),
                     [this](size_t a, size_t b) { return counts_[a] > counts_[b]; });
    LOG(INFO) << "Popularity: " << num_items_ << " items ranked, most popular item has "
              << (num_items_ ? counts_[ranking_[0]] : 0u) << " interactions";
  }

  virtual std::vector<size_t> recommend(size_t /*user_id*/, size_t topk,
                                        const std::unordered_map<size_t, double>& rated_items_map) const {
    std::vector<size_t> out;
    out.reserve(topk);
    for (size_t r = 0; r < ranking_.size() && out.size() < topk; ++r)
      if (!rated_items_map.count(ranking_[r])) out.push_back(ranking_[r]);
    CHECK(out.size() == topk || rated_items_map.size() > num_items_ - topk);
    return out;
  }

 protected:
  std::vector<uint32_t> counts_;   // interactions per item
  std::vector<size_t> ranking_;    // item ids, most popular first
};

}  // namespace libcf

#endif  // _LIBCF_POPULARITY_HPP_
