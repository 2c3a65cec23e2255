// matcher_cuda_all_in_memory.h -- drop-in CUDA implementation of the reference's collection matcher.
//
//   mvg::feature::MatcherCudaAllInMemory<KeypointSetT>  :  mvg::feature::Matcher   (matcher.h:14-31)
//
// Same constructor / LoadData / Match signatures as MatcherAllInMemory<KeypointSetT, MatcherT>
// (matcher_all_in_memory.h:19-147); replaces `MatcherAllInMemory<KeypointSetT, MatcherT> collectionMatcher(ratio)`
// at apps/compute_matches/compute_matches.cpp:237 one-for-one.  Match() uploads every descriptor array once,
// shards the i<j pair list over `n_gpus` contexts (one host thread per GPU, no collective: pairs are independent),
// runs rows 7-13 of the path (incl. the coordinate de-duplication of IndexedMatchDecorator) on the GPUs,
// and inserts EVERY pair -- empty ones included (matcher_all_in_memory.h:135) -- into the map.
// Documented deviation: images with fewer than 2 descriptors yield empty pairs where the reference has undefined
// behaviour (null Eigen::Map).
#ifndef MVGCUDA_MATCHER_CUDA_ALL_IN_MEMORY_H_
#define MVGCUDA_MATCHER_CUDA_ALL_IN_MEMORY_H_

#include <algorithm>
#include <iostream>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "mvg/feature/features.h"
#include "mvg/feature/matcher.h"
#include "mvg/utils/file_system.h"
#include "mvgcuda.h"

namespace mvg {
namespace feature {

template <typename KeypointSetT>
class MatcherCudaAllInMemory : public Matcher {
  typedef typename KeypointSetT::FeatureT FeatureT;
  typedef typename KeypointSetT::DescriptorT DescriptorT;
  typedef std::vector<DescriptorT> DescsT;
  static_assert(sizeof(DescriptorT) == MVGCUDA_DIM, "MatcherCudaAllInMemory needs Descriptor<unsigned char,128>");

 public:
  explicit MatcherCudaAllInMemory(float distRatio, int n_gpus = 1) : Matcher(), distance_ratio(distRatio), n_gpus_(n_gpus) {}

  bool LoadData(const std::vector<std::string>& file_names, const std::string& match_dir) {
    bool is_ok = true;
    for (size_t j = 0; j < file_names.size(); ++j) {
      const std::string feat_filename =
          mvg::utils::create_filespec(match_dir, mvg::utils::basename_part(file_names[j]), "feat");
      const std::string desc_filename =
          mvg::utils::create_filespec(match_dir, mvg::utils::basename_part(file_names[j]), "desc");
      is_ok &= LoadFeatsFromFile(feat_filename, map_features[j]);
      is_ok &= LoadDescsFromBinFile(desc_filename, map_descriptors[j]);
    }
    return is_ok;
  }

  void Match(const std::vector<std::string>& file_names, PairWiseMatches& map_putatives_matches) const {
    const int n = (int)file_names.size();
    std::vector<const uint8_t*> desc(n, (const uint8_t*)NULL);
    std::vector<const float*> xy(n, (const float*)NULL);
    std::vector<std::vector<float> > xy_store(n);
    std::vector<int32_t> rows(n, 0);
    for (int i = 0; i < n; ++i) {
      const std::vector<FeatureT>& f = map_features.find(i)->second;
      const DescsT& d = map_descriptors.find(i)->second;
      // the reference takes the row count from the features (matcher_all_in_memory.h:80,85,107); never read past
      // the descriptors that were actually loaded
      rows[i] = (int32_t)std::min(f.size(), d.size());
      if (rows[i] > 0) desc[i] = reinterpret_cast<const uint8_t*>(&d[0]);
      xy_store[i].resize(2 * (size_t)rows[i]);
      for (int k = 0; k < rows[i]; ++k) { xy_store[i][2 * k] = f[k].x(); xy_store[i][2 * k + 1] = f[k].y(); }
      xy[i] = xy_store[i].data();
    }
    std::vector<int32_t> pairs;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) { pairs.push_back(i); pairs.push_back(j); }
    const int64_t n_pairs = (int64_t)pairs.size() / 2;
    const float ratio_sq = Square(distance_ratio);  // fp32 product, numeric.h:108-111
    const int gpus = std::max(1, std::min<int>(n_gpus_, (int)std::max<int64_t>(n_pairs, 1)));
    // contiguous cost-balanced shards (cost = rows_i * rows_j)
    std::vector<int64_t> bounds(gpus + 1, n_pairs);
    {
      std::vector<double> csum(n_pairs + 1, 0.0);
      for (int64_t p = 0; p < n_pairs; ++p)
        csum[p + 1] = csum[p] + std::max(1.0, (double)rows[pairs[2 * p]] * (double)rows[pairs[2 * p + 1]]);
      bounds[0] = 0;
      for (int g = 1; g < gpus; ++g)
        bounds[g] = std::lower_bound(csum.begin(), csum.end(), csum[n_pairs] * g / gpus) - csum.begin();
      for (int g = 1; g <= gpus; ++g) bounds[g] = std::max(bounds[g], bounds[g - 1]);
    }
    std::vector<std::vector<std::vector<IndexedMatch> > > shard_out(gpus);
    std::vector<std::string> errors(gpus);
    // GPU 0 reads the collection over PCIe; the other GPUs take a replica of its arena device to device (NVLink)
    std::vector<mvgcuda_ctx*> ctxs(gpus, (mvgcuda_ctx*)NULL);
    // shard g runs on the g-th sm_100 device of the box (CUDA ordinals need not be 0..N-1 on a mixed box)
    if (mvgcuda_create(std::max(0, mvgcuda_device_ordinal(0)), &ctxs[0]) != MVGCUDA_OK) {
      errors[0] = mvgcuda_last_error(NULL);
    } else if (mvgcuda_upload_images(ctxs[0], n, desc.data(), rows.data(), 0) != MVGCUDA_OK ||
               mvgcuda_set_features(ctxs[0], n, xy.data(), rows.data()) != MVGCUDA_OK) {
      errors[0] = mvgcuda_last_error(ctxs[0]);
    }
    auto work = [&](int g) {
      if (g > 0) {
        if (mvgcuda_create(std::max(0, mvgcuda_device_ordinal(g)), &ctxs[g]) != MVGCUDA_OK) { errors[g] = mvgcuda_last_error(NULL); return; }
        if (mvgcuda_clone_images(ctxs[g], ctxs[0]) != MVGCUDA_OK) { errors[g] = mvgcuda_last_error(ctxs[g]); return; }
      }
      mvgcuda_ctx* ctx = ctxs[g];
      const int64_t b = bounds[g], e = bounds[g + 1];
      mvgcuda_pair_matches pm;
      if (mvgcuda_match_collection(ctx, e - b, pairs.data() + 2 * b, ratio_sq, 0, &pm) != MVGCUDA_OK) {
        errors[g] = mvgcuda_last_error(ctx);
        return;
      }
      shard_out[g].resize(e - b);
      for (int64_t p = 0; p < e - b; ++p) {
        std::vector<IndexedMatch>& v = shard_out[g][p];
        v.reserve(pm.counts[p]);
        const int32_t* m = pm.matches + 2 * pm.offsets[p];
        for (int k = 0; k < pm.counts[p]; ++k) v.push_back(IndexedMatch(m[2 * k], m[2 * k + 1]));
      }
    };
    if (errors[0].empty()) {
      std::vector<std::thread> th;
      for (int g = 0; g < gpus; ++g) th.emplace_back(work, g);
      for (auto& t : th) t.join();
    }
    for (int g = 0; g < gpus; ++g)
      if (ctxs[g]) mvgcuda_destroy(ctxs[g]);  // only after every replica has been taken
    for (int g = 0; g < gpus; ++g) {
      if (!errors[g].empty()) {
        std::cerr << "MatcherCudaAllInMemory: GPU " << g << ": " << errors[g] << std::endl;
        return;  // no CPU fallback
      }
      for (int64_t p = bounds[g]; p < bounds[g + 1]; ++p)
        map_putatives_matches.insert(std::make_pair(std::make_pair((size_t)pairs[2 * p], (size_t)pairs[2 * p + 1]),
                                                    shard_out[g][p - bounds[g]]));
    }
  }

 private:
  std::map<size_t, std::vector<FeatureT> > map_features;
  std::map<size_t, DescsT> map_descriptors;
  float distance_ratio;
  int n_gpus_;
};

}  // namespace feature
}  // namespace mvg

#endif  // MVGCUDA_MATCHER_CUDA_ALL_IN_MEMORY_H_
