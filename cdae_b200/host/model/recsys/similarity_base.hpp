// cdae_b200/host/model/recsys/similarity_base.hpp — neighbourhood models of the drop-in host tree.
//
// Same surface as the reference's src/model/recsys/similarity_base.hpp:12-126 (SimilarityType,
// SimilarityBase(index group, data group, type, topk), reset, get_neighbors, no-op training), built on
// two CSR incidence structures instead of two hash tables of vectors: for index entity x (an item for
// ItemCF) the co-occurrence count with every other entity is accumulated in a dense per-thread scratch
// array over the entities that share a data entity (a user) with x; Jaccard = c / (n_x + n_y - c),
// Cosine = c / sqrt(n_x n_y) (reference :84-91); the topk most similar are kept, most similar first.
#ifndef _LIBCF_SIMILARITY_BASE_HPP_
#define _LIBCF_SIMILARITY_BASE_HPP_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <utility>
#include <vector>

#include <base/parallel.hpp>
#include <model/recsys/recsys_model_base.hpp>

namespace libcf {

enum SimilarityType { Jaccard, Cosine };

inline std::ostream& operator<<(std::ostream& out, const SimilarityType& st) {
  if (st == Jaccard) out << "Jaccard";
  else if (st == Cosine) out << "Cosine";
  else LOG(FATAL) << "Undefined similarity type!";
  return out;
}

class SimilarityBase : public RecsysModelBase {
 public:
  SimilarityBase(size_t index_feature_group, size_t data_feature_group, SimilarityType sim_type, size_t topk)
      : sim_type_(sim_type), topk_(topk), index_feature_group_(index_feature_group),
        data_feature_group_(data_feature_group) {}

  virtual void reset(const Data& data_set) {
    data_ = &data_set;
    Timer timer;
    CHECK_LT(index_feature_group_, data_set.num_feature_groups());
    CHECK_LT(data_feature_group_, data_set.num_feature_groups());
    const size_t n_index = data_set.feature_group_total_dimension(index_feature_group_);
    const size_t n_data = data_set.feature_group_total_dimension(data_feature_group_);
    // incidence pairs (index entity, data entity), de-duplicated, as two CSRs
    std::vector<std::pair<uint32_t, uint32_t>> pairs;
    for (auto it = data_set.begin(); it != data_set.end(); ++it)
      pairs.emplace_back((uint32_t)it->get_feature_group_index(index_feature_group_, 0),
                         (uint32_t)it->get_feature_group_index(data_feature_group_, 0));
    std::sort(pairs.begin(), pairs.end());
    pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
    std::vector<size_t> ip(n_index + 1, 0), dp(n_data + 1, 0);
    for (auto& p : pairs) { ++ip[p.first + 1]; ++dp[p.second + 1]; }
    for (size_t i = 0; i < n_index; ++i) ip[i + 1] += ip[i];
    for (size_t i = 0; i < n_data; ++i) dp[i + 1] += dp[i];
    std::vector<uint32_t> i2d(pairs.size()), d2i(pairs.size());
    {
      std::vector<size_t> fill(dp.begin(), dp.end() - 1);
      for (size_t k = 0; k < pairs.size(); ++k) {       // pairs are sorted by index entity: i2d fills in order
        i2d[k] = pairs[k].second;
        d2i[fill[pairs[k].second]++] = pairs[k].first;
      }
    }
    topk_neighbors_.assign(n_index, {});
    const size_t n_threads = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), 16));
    auto work = [&](size_t tid) {
      std::vector<float> co(n_index, 0.f);               // co-occurrence counts with the current entity
      std::vector<uint32_t> touched;
      std::vector<std::pair<size_t, double>> cand;
      for (size_t x = tid; x < n_index; x += n_threads) {
        const double nx = (double)(ip[x + 1] - ip[x]);
        if (nx == 0) continue;
        touched.clear();
        for (size_t a = ip[x]; a < ip[x + 1]; ++a) {
          const uint32_t d = i2d[a];
          for (size_t b = dp[d]; b < dp[d + 1]; ++b) {
            const uint32_t y = d2i[b];
            if (y == x) continue;
            if (co[y] == 0.f) touched.push_back(y);
            co[y] += 1.f;
          }
        }
        cand.clear();
        for (uint32_t y : touched) {
          const double c = co[y], ny = (double)(ip[y + 1] - ip[y]);
          co[y] = 0.f;
          cand.emplace_back((size_t)y, sim_type_ == Jaccard ? c / (nx + ny - c) : c / std::sqrt(nx * ny));
        }
        const size_t keep = std::min(topk_, cand.size());
        std::partial_sort(cand.begin(), cand.begin() + keep, cand.end(),
                          [](const std::pair<size_t, double>& a, const std::pair<size_t, double>& b) {
                            return a.second > b.second || (a.second == b.second && a.first < b.first);
                          });
        topk_neighbors_[x].assign(cand.begin(), cand.begin() + keep);
      }
    };
    std::vector<std::thread> pool;
    for (size_t t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    // the reference's protected hash tables (similarity_base.hpp:117-118): subclasses outside this tree read
    // them (UserCF::recommend walks index_data_pair, usercf.hpp:29-31)
    index_data_pair.clear();
    data_index_pair.clear();
    for (size_t x = 0; x < n_index; ++x)
      if (ip[x + 1] > ip[x]) index_data_pair[x].assign(i2d.begin() + ip[x], i2d.begin() + ip[x + 1]);
    for (size_t d = 0; d < n_data; ++d)
      if (dp[d + 1] > dp[d]) data_index_pair[d].assign(d2i.begin() + dp[d], d2i.begin() + dp[d + 1]);
    LOG(INFO) << "Finished getting nearest neighbors in " << timer;
  }

  virtual double data_loss(const Data&, size_t = 0) const { return 0.0; }

  virtual std::vector<size_t> recommend(size_t, size_t, const std::unordered_map<size_t, double>&) const {
    LOG(FATAL) << "UnImplemented!";
    return std::vector<size_t>{};
  }

  virtual void train_one_iteration(const Data&) {}

  std::vector<std::vector<std::pair<size_t, double>>> get_neighbors() const { return topk_neighbors_; }

 protected:
  std::vector<std::vector<std::pair<size_t, double>>> topk_neighbors_;
  std::unordered_map<size_t, std::vector<size_t>> data_index_pair;   // data entity -> index entities
  std::unordered_map<size_t, std::vector<size_t>> index_data_pair;   // index entity -> data entities
  enum SimilarityType sim_type_;
  size_t topk_;
  size_t index_feature_group_;
  size_t data_feature_group_;
};

}  // namespace libcf

#endif  // _LIBCF_SIMILARITY_BASE_HPP_
