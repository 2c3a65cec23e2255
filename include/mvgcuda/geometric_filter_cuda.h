// geometric_filter_cuda.h -- drop-in CUDA implementation of the reference's collection geometric filter.
//
//   mvg::feature::ImageCollectionGeometricFilterCuda<FeatureT>   ~   ImageCollectionGeometricFilter<FeatureT>
//                                                                     (geometric_filter.h:17-106)
//
// Same LoadData / Filter signatures; replaces `ImageCollectionGeometricFilter<FeatureT> collection_geom_filter;` at
// apps/compute_matches/compute_matches.cpp:252 one-for-one for the FUNDAMENTAL_MATRIX case (:259-266):
//
//     collection_geom_filter.Filter(GeometricFilter_FMatrix_AC(max_residual_error), map_putatives_matches,
//                                   map_geometric_matches, vec_images_size);
//
// Filter() uploads the feature coordinates, runs the AC-RANSAC of every pair on the GPU (mvgcuda_geometric_filter: the
// reference's sample stream -- glibc rand(), consumed pair after pair in map order --, control flow and double-precision
// solver) and inserts the pairs that keep inliers, in the reference's (residual) order.  GeometricFilter_FMatrix_AC and
// GeometricFilter_HMatrix_AC are accepted (their precision / iteration members are read); the essential functor has no GPU
// implementation and is refused at compile time.
#ifndef MVGCUDA_GEOMETRIC_FILTER_CUDA_H_
#define MVGCUDA_GEOMETRIC_FILTER_CUDA_H_

#include <iostream>
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include "mvg/feature/features.h"
#include "mvg/feature/indexed_match.h"
#include "mvg/multiview/fundamental_acransac.h"
#include "mvg/multiview/homography_acransac.h"
#include "mvg/utils/file_system.h"
#include "mvgcuda.h"

namespace mvg {
namespace feature {

template <typename FeatureT>
class ImageCollectionGeometricFilterCuda {
 public:
  explicit ImageCollectionGeometricFilterCuda(int device_index = 0, unsigned rand_seed = 1)  // 1: the reference never calls srand()
      : device_index_(device_index), seed_(rand_seed) {}

  /// geometric_filter.h:24-34
  bool LoadData(const std::vector<std::string>& file_names, const std::string& match_dir) {
    bool is_ok = true;
    for (size_t j = 0; j < file_names.size(); ++j) {
      const std::string feat_filename = mvg::utils::create_filespec(match_dir, mvg::utils::basename_part(file_names[j]), "feat");
      is_ok &= LoadFeatsFromFile(feat_filename, map_features[j]);
    }
    return is_ok;
  }

  /// geometric_filter.h:37-102
  template <typename GeometricFilterT>
  void Filter(const GeometricFilterT& geometric_filter, PairWiseMatches& map_putatives_matches_pair,
              PairWiseMatches& map_geometric_matches, const std::vector<std::pair<size_t, size_t> >& vec_images_size) const {
    static_assert(std::is_same<GeometricFilterT, mvg::multiview::GeometricFilter_FMatrix_AC>::value ||
                      std::is_same<GeometricFilterT, mvg::multiview::GeometricFilter_HMatrix_AC>::value,
                  "ImageCollectionGeometricFilterCuda: only GeometricFilter_FMatrix_AC and GeometricFilter_HMatrix_AC run on the GPU");
    const int n = (int)map_features.size();
    std::vector<std::vector<float> > xy(n);
    std::vector<const float*> xy_ptr(n, (const float*)NULL);
    std::vector<const uint8_t*> no_desc(n, (const uint8_t*)NULL);
    std::vector<int32_t> rows(n, 0), sizes(2 * (size_t)n, 0);
    {
      int i = 0;
      for (typename std::map<size_t, std::vector<FeatureT> >::const_iterator it = map_features.begin(); it != map_features.end(); ++it, ++i) {
        rows[i] = (int32_t)it->second.size();
        xy[i].resize(2 * it->second.size());
        for (size_t k = 0; k < it->second.size(); ++k) { xy[i][2 * k] = it->second[k].x(); xy[i][2 * k + 1] = it->second[k].y(); }
        xy_ptr[i] = xy[i].data();
        if (i < (int)vec_images_size.size()) { sizes[2 * i] = (int32_t)vec_images_size[i].first; sizes[2 * i + 1] = (int32_t)vec_images_size[i].second; }
      }
    }
    std::vector<int32_t> pairs, counts, matches;
    std::vector<int64_t> offsets(1, 0);
    for (PairWiseMatches::const_iterator it = map_putatives_matches_pair.begin(); it != map_putatives_matches_pair.end(); ++it) {
      pairs.push_back((int32_t)it->first.first);
      pairs.push_back((int32_t)it->first.second);
      counts.push_back((int32_t)it->second.size());
      for (size_t k = 0; k < it->second.size(); ++k) { matches.push_back((int32_t)it->second[k]._i); matches.push_back((int32_t)it->second[k]._j); }
      offsets.push_back(offsets.back() + (int64_t)it->second.size());
    }
    if (matches.empty()) matches.assign(2, 0);
    mvgcuda_ctx* ctx = NULL;
    if (mvgcuda_create(std::max(0, mvgcuda_device_ordinal(device_index_)), &ctx) != MVGCUDA_OK) {
      std::cerr << "ImageCollectionGeometricFilterCuda: " << mvgcuda_last_error(NULL) << std::endl;
      return;  // no CPU fallback
    }
    // only the coordinates are needed on the device: an arena with the right row counts and no descriptors
    mvgcuda_pair_matches gm;
    if (mvgcuda_stream_begin(ctx, n, rows.data()) != MVGCUDA_OK || !stream_features(ctx, n, rows, xy_ptr) || mvgcuda_stream_end(ctx) != MVGCUDA_OK ||
        mvgcuda_geometric_filter(ctx, model_of(geometric_filter), geometric_filter.m_dPrecision, iterations_of(geometric_filter), seed_, (int64_t)counts.size(),
                                 pairs.data(), counts.data(), offsets.data(), matches.data(), sizes.data(), &gm) != MVGCUDA_OK) {
      std::cerr << "ImageCollectionGeometricFilterCuda: " << mvgcuda_last_error(ctx) << std::endl;
      mvgcuda_destroy(ctx);
      return;
    }
    for (int64_t p = 0; p < gm.n_pairs; ++p) {
      if (gm.counts[p] == 0) continue;  // geometric_filter.h:85: only pairs that keep inliers enter the map
      std::vector<IndexedMatch> v;
      v.reserve(gm.counts[p]);
      const int32_t* m = gm.matches + 2 * gm.offsets[p];
      for (int k = 0; k < gm.counts[p]; ++k) v.push_back(IndexedMatch(m[2 * k], m[2 * k + 1]));
      map_geometric_matches[std::make_pair((size_t)pairs[2 * p], (size_t)pairs[2 * p + 1])] = v;
    }
    mvgcuda_destroy(ctx);
  }

 private:
  static char model_of(const mvg::multiview::GeometricFilter_FMatrix_AC&) { return 'f'; }
  static char model_of(const mvg::multiview::GeometricFilter_HMatrix_AC&) { return 'h'; }
  static int iterations_of(const mvg::multiview::GeometricFilter_FMatrix_AC& f) { return (int)f.max_iteration; }
  static int iterations_of(const mvg::multiview::GeometricFilter_HMatrix_AC& f) { return (int)f.m_stIteration; }
  static bool stream_features(mvgcuda_ctx* ctx, int n, const std::vector<int32_t>& rows, const std::vector<const float*>& xy) {
    std::vector<std::vector<uint8_t> > zeros(n);
    for (int i = 0; i < n; ++i) {
      if (rows[i] == 0) continue;
      zeros[i].assign((size_t)rows[i] * MVGCUDA_DIM, 0);  // the arena wants descriptor rows; the filter never reads them
      if (mvgcuda_stream_image(ctx, i, zeros[i].data(), xy[i]) != MVGCUDA_OK) return false;
      if (mvgcuda_stream_end(ctx) != MVGCUDA_OK) return false;  // the pageable staging above is released right away
    }
    return true;
  }

  std::map<size_t, std::vector<FeatureT> > map_features;
  int device_index_;
  unsigned seed_;
};

}  // namespace feature
}  // namespace mvg

#endif  // MVGCUDA_GEOMETRIC_FILTER_CUDA_H_
