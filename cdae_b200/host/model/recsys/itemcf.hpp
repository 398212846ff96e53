// cdae_b200/host/model/recsys/itemcf.hpp — libcf::ItemCF for the drop-in host tree
// (reference: src/model/recsys/itemcf.hpp:10-53; apps/yelp runs it with Jaccard, 50 neighbours,
// yelp.cpp:115-119).  score(candidate) = sum over the user's rated items of sim(rated, candidate) over
// the rated item's stored neighbours; rated items are excluded; top-k by score, fewer if fewer
// candidates exist.
#ifndef _LIBCF_ITEMCF_HPP_
#define _LIBCF_ITEMCF_HPP_

#include <algorithm>
#include <unordered_map>
#include <utility>
#include <vector>

#include <model/recsys/similarity_base.hpp>

namespace libcf {

class ItemCF : public SimilarityBase {
 public:
  // index entities = items (feature group 1), data entities = users (feature group 0)
  ItemCF(SimilarityType sim_type = Jaccard, size_t topk = 50) : SimilarityBase(1, 0, sim_type, topk) {
    LOG(INFO) << "Item Similarity Model {SimType: " << sim_type_ << "}, {TOPK: " << topk << "}";
  }

  // Candidates are the stored neighbours of the user's rated items that the user has not rated; a candidate's
  // score is the sum of its similarities to those rated items.  Scores are accumulated in one flat vector of
  // (item, score) pairs sorted by item, then the best `topk` are selected (ties: lower item id first).
  virtual std::vector<size_t> recommend(size_t /*uid*/, size_t topk,
                                        const std::unordered_map<size_t, double>& rated_map) const {
    std::vector<std::pair<size_t, double>> hits;
    for (const auto& rated : rated_map) {
      if (rated.first >= topk_neighbors_.size()) continue;
      for (const auto& nb : topk_neighbors_[rated.first])
        if (rated_map.find(nb.first) == rated_map.end()) hits.push_back(nb);
    }
    std::sort(hits.begin(), hits.end(), [](const std::pair<size_t, double>& a, const std::pair<size_t, double>& b) { return a.first < b.first; });
    size_t n = 0;                                       // merge runs of the same item
    for (size_t i = 0; i < hits.size(); ++i) {
      if (n > 0 && hits[n - 1].first == hits[i].first) hits[n - 1].second += hits[i].second;
      else hits[n++] = hits[i];
    }
    hits.resize(n);
    const size_t keep = std::min(topk, hits.size());
    std::partial_sort(hits.begin(), hits.begin() + keep, hits.end(),
                      [](const std::pair<size_t, double>& a, const std::pair<size_t, double>& b) {
                        return a.second > b.second || (a.second == b.second && a.first < b.first);
                      });
    std::vector<size_t> out(keep);
    for (size_t k = 0; k < keep; ++k) out[k] = hits[k].first;
    return out;
  }
};

}  // namespace libcf

#endif  // _LIBCF_ITEMCF_HPP_
