// cdae_b200/host/model/recsys/itemcf.hpp — libcf::ItemCF for the drop-in host tree
// (reference: src/model/recsys/itemcf.hpp:10-53; apps/yelp runs it with Jaccard, 50 neighbours,
// yelp.cpp:115-119).  score(candidate) = sum over the user's rated items of sim(rated, candidate) over
// the rated item's stored neighbours; rated items are excluded; top-k by score, fewer if fewer
// candidates exist.
#ifndef _LIBCF_ITEMCF_HPP_
#define _LIBCF_ITEMCF_HPP_

#include <unordered_map>

#include <model/recsys/similarity_base.hpp>

namespace libcf {

class ItemCF : public SimilarityBase {
 public:
  ItemCF(SimilarityType sim_type = Jaccard, size_t topk = 50) : SimilarityBase(1, 0, sim_type, topk) {
    LOG(INFO) << "Item Similarity Model";
    LOG(INFO) << "\t{SimType: " << sim_type_ << "}, " << "{TOPK: " << topk << "}";
  }

  virtual std::vector<size_t> recommend(size_t /*uid*/, size_t topk,
                                        const std::unordered_map<size_t, double>& rated_map) const {
    std::unordered_map<size_t, double> score;
    for (const auto& rated : rated_map) {
      if (rated.first >= topk_neighbors_.size()) continue;
      for (const auto& nb : topk_neighbors_[rated.first])
        if (!rated_map.count(nb.first)) score[nb.first] += nb.second;
    }
    std::vector<std::pair<size_t, double>> ranked(score.begin(), score.end());
    const size_t keep = std::min(topk, ranked.size());
    std::partial_sort(ranked.begin(), ranked.begin() + keep, ranked.end(),
                      [](const std::pair<size_t, double>& a, const std::pair<size_t, double>& b) {
                        return a.second > b.second || (a.second == b.second && a.first < b.first);
                      });
    std::vector<size_t> out(keep);
    for (size_t k = 0; k < keep; ++k) out[k] = ranked[k].first;
    return out;
  }
};

}  // namespace libcf

#endif  // _LIBCF_ITEMCF_HPP_
