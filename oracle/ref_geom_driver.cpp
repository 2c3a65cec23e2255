// ref_geom_driver.cpp -- ORACLE L0 for the step AFTER the hot path: the reference's own
// ImageCollectionGeometricFilter + GeometricFilter_{F,H}Matrix_AC (AC-RANSAC), compiled from where they lie under
// $MVG_REF by oracle/build_ref.sh into oracle/_ref/libmvgref_geom.so.  TEST INFRASTRUCTURE ONLY.
//
//   geometric_filter.h:37-102            collection loop (one AC-RANSAC per pair, pairs in std::map order)
//   fundamental_acransac.h:13-57         GeometricFilter_FMatrix_AC (7-point solver, 4096 iterations, inlier floor 2.5 x 7)
//   homography_acransac.h                GeometricFilter_HMatrix_AC
//   estimator_acransac.h:125-245         ACRANSAC; random_sampling.h:46-60 draws with glibc rand() -- the stream is
//                                        never seeded by the reference (== srand(1)) and is consumed pair after pair,
//                                        so results are defined by the pair order; `seed` pins it here.
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

// Prelude required by the reference headers under g++ (SURVEY.md 8(c) recipe): they rely on MSVC's lax two-phase lookup
// for ControlProgressDisplay / make_pair, so the using-directives come BEFORE the headers.
#include <algorithm>
#include <map>
#include "mvg/utils/progress.h"
using namespace std;
using namespace mvg::utils;

#include "mvg/feature/features.h"
#include "mvg/feature/indexed_match_utils.h"
#include "mvg/feature/geometric_filter.h"
#include "mvg/multiview/fundamental_acransac.h"
#include "mvg/multiview/homography_acransac.h"

using namespace mvg;
using namespace mvg::feature;
using namespace mvg::multiview;

typedef ScalePointFeature FeatureT;

extern "C" {

// names: '\n'-separated image file names (lists.txt order); sizes: [n][2] = width, height (lists.txt columns 2, 3);
// model: 'f' or 'h'; max_residual: compute_matches.cpp:254 uses 4.0.  Imports putative_path with the reference's own
// reader, filters, exports with the reference's own writer.  Returns the number of pairs kept, -1 on error.
int ref_geometric_filter(const char* match_dir, const char* names, const int* sizes, const char* putative_path, char model,
                         double max_residual, unsigned seed, const char* out_path) {
  std::vector<std::string> file_names;
  {
    std::istringstream ss(names);
    std::string line;
    while (std::getline(ss, line))
      if (!line.empty()) file_names.push_back(line);
  }
  std::vector<std::pair<size_t, size_t> > vec_images_size;
  for (size_t k = 0; k < file_names.size(); ++k) vec_images_size.push_back(std::make_pair((size_t)sizes[2 * k], (size_t)sizes[2 * k + 1]));
  PairWiseMatches putatives, geometric;
  std::streambuf* old = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());
  const bool ok = pairedIndexedMatchImport(putative_path, putatives);
  ImageCollectionGeometricFilter<FeatureT> filter;
  const bool loaded = ok && filter.LoadData(file_names, match_dir);
  if (loaded) {
    srand(seed);
    if (model == 'f') filter.Filter(GeometricFilter_FMatrix_AC(max_residual), putatives, geometric, vec_images_size);
    else if (model == 'h') filter.Filter(GeometricFilter_HMatrix_AC(max_residual), putatives, geometric, vec_images_size);
  }
  std::cout.rdbuf(old);
  if (!loaded || (model != 'f' && model != 'h')) return -1;
  std::ofstream file(out_path);
  if (!file.is_open()) return -1;
  PairedIndexedMatchToStream(geometric, file);
  return (int)geometric.size();
}

}  // extern "C"
