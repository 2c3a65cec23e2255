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

// ---- entry points for the unit tests of the restatement (tests/native/test_acransac_core.cpp, tests/test_oracle.py) ----

// SevenPointSolver::Solve on 7 correspondences ([7][2] each); F: up to 3 models, row-major 3x3.  Returns the model count.
int ref_seven_point(const double* x1, const double* x2, double* F) {
  Mat a(2, 7), b(2, 7);
  for (int i = 0; i < 7; ++i) { a(0, i) = x1[2 * i]; a(1, i) = x1[2 * i + 1]; b(0, i) = x2[2 * i]; b(1, i) = x2[2 * i + 1]; }
  std::vector<Mat3> models;
  fundamental::SevenPointSolver::Solve(a, b, &models);
  for (size_t k = 0; k < models.size(); ++k)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) F[9 * k + 3 * r + c] = models[k](r, c);
  return (int)models.size();
}

// homography::FourPointSolver::Solve on 4 correspondences ([4][2] each) -> H, row-major 3x3 (one model).
int ref_four_point(const double* x1, const double* x2, double* H) {
  Mat a(2, 4), b(2, 4);
  for (int i = 0; i < 4; ++i) { a(0, i) = x1[2 * i]; a(1, i) = x1[2 * i + 1]; b(0, i) = x2[2 * i]; b(1, i) = x2[2 * i + 1]; }
  std::vector<Mat3> models;
  homography::FourPointSolver::Solve(a, b, &models);
  for (size_t k = 0; k < models.size(); ++k)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) H[9 * k + 3 * r + c] = models[k](r, c);
  return (int)models.size();
}

double ref_homography_error(const double* H, double x1, double y1, double x2, double y2) {
  Mat3 M;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M(r, c) = H[3 * r + c];
  return homography::AsymmetricError::Error(M, Vec2(x1, y1), Vec2(x2, y2));
}

// NormalizePoints(points, &out, &T, width, height) on n float points ([n][2]); T: row-major 3x3.
void ref_normalize(const float* pts, int n, int width, int height, double* out, double* T) {
  Mat x(2, n), xn;
  for (int i = 0; i < n; ++i) x.col(i) = Vec2f(pts[2 * i], pts[2 * i + 1]).cast<double>();
  Mat3 N;
  NormalizePoints(x, &xn, &N, width, height);
  for (int i = 0; i < n; ++i) { out[2 * i] = xn(0, i); out[2 * i + 1] = xn(1, i); }
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) T[3 * r + c] = N(r, c);
}

double ref_epipolar_error(const double* F, double x1, double y1, double x2, double y2) {
  Mat3 M;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M(r, c) = F[3 * r + c];
  return fundamental::SimpleError::Error(M, Vec2(x1, y1), Vec2(x2, y2));
}

void ref_rand(unsigned seed, int n, unsigned* out) {
  srand(seed);
  for (int i = 0; i < n; ++i) out[i] = (unsigned)rand();
}

// How many values the process-wide rand() stream has yielded since srand(seed): the next four values are drawn and
// located in a regenerated stream (the state is left re-seeded -- call this after a run, not inside one).  -1: beyond limit.
long ref_rand_position(unsigned seed, long limit) {
  unsigned w[4];
  for (int i = 0; i < 4; ++i) w[i] = (unsigned)rand();
  srand(seed);
  unsigned a = (unsigned)rand(), b = (unsigned)rand(), c = (unsigned)rand(), d = (unsigned)rand();
  for (long pos = 0; pos <= limit; ++pos) {
    if (a == w[0] && b == w[1] && c == w[2] && d == w[3]) return pos;
    a = b; b = c; c = d; d = (unsigned)rand();
  }
  return -1;
}

void ref_random_sample7(unsigned seed, int skip, int n, int* out) {
  srand(seed);
  for (int i = 0; i < skip; ++i) rand();
  std::vector<size_t> s;
  RandomSample(7, (size_t)n, &s);
  for (int i = 0; i < 7; ++i) out[i] = (int)s[i];
}

void ref_logc(int n, float* logc_n, float* logc_k) {
  std::vector<float> a, b;
  makelogcombi_n((size_t)n, a);
  makelogcombi_k(7, (size_t)n, b);
  for (int k = 0; k <= n; ++k) { logc_n[k] = a[k]; logc_k[k] = b[k]; }
}

// One pair through the reference's own kernel + ACRANSAC (fundamental_acransac.h:23-47 without the final 2.5 x 7 floor):
// xI, xJ [n][2] float feature coordinates.  inliers: capacity n.  out[0] = errorMax, out[1] = minNFA, out[2] = rand() calls
// consumed (counted by re-running the stream).  Returns the number of inliers.
}  // extern "C"

// ACRANSAC on one pair with the kernel a filter functor builds; out = {errorMax, minNFA, rand() values consumed}.
template <typename KernelType>
static int run_acransac(const float* xI, const float* xJ, int n, int wI, int hI, int wJ, int hJ, bool point_to_line, double precision,
                        int iterations, unsigned seed, int* inliers, double* out, long skip = 0, bool verbose = false) {
  Mat a(2, n), b(2, n);
  for (int i = 0; i < n; ++i) {
    a.col(i) = Vec2f(xI[2 * i], xI[2 * i + 1]).cast<double>();
    b.col(i) = Vec2f(xJ[2 * i], xJ[2 * i + 1]).cast<double>();
  }
  KernelType kernel(a, wI, hI, b, wJ, hJ, point_to_line);
  std::vector<size_t> vec_inliers;
  Mat3 M;
  srand(seed);
  for (long i = 0; i < skip; ++i) rand();  // the pair starts where the pairs before it left the process-wide stream
  std::pair<double, double> r = ACRANSAC(kernel, vec_inliers, (size_t)iterations, &M, precision, verbose);
  // how far the stream moved: the position (a multiple of the sample size) whose value is what rand() returns now
  const unsigned next = (unsigned)rand();
  srand(seed);
  for (long i = 0; i < skip; ++i) rand();
  const size_t sample = KernelType::MINIMUM_SAMPLES;
  std::vector<unsigned> all((size_t)iterations * sample + 8);
  for (size_t i = 0; i < all.size(); ++i) all[i] = (unsigned)rand();
  long used = -1;
  for (size_t i = 0; i < all.size(); ++i) if (all[i] == next && (i % sample) == 0) { used = (long)i; break; }
  for (size_t i = 0; i < vec_inliers.size(); ++i) inliers[i] = (int)vec_inliers[i];
  out[0] = r.first; out[1] = r.second; out[2] = (double)used;
  return (int)vec_inliers.size();
}

extern "C" {

// GeometricFilter_FMatrix_AC::Fit's kernel (fundamental_acransac.h:30-42): point-to-line.
int ref_acransac_f(const float* xI, const float* xJ, int n, int wI, int hI, int wJ, int hJ, double precision, int iterations,
                   unsigned seed, int* inliers, double* out) {
  typedef ACKernelAdaptor<fundamental::SevenPointSolver, fundamental::SimpleError, UnnormalizerT, Mat3> KernelType;
  return run_acransac<KernelType>(xI, xJ, n, wI, hI, wJ, hJ, true, precision, iterations, seed, inliers, out);
}

// The same with the stream advanced by `skip` values first (a pair in the middle of a collection); model 'f' / 'h'.
int ref_acransac_at(char model, const float* xI, const float* xJ, int n, int wI, int hI, int wJ, int hJ, double precision, int iterations,
                    unsigned seed, long skip, int verbose, int* inliers, double* out) {
  typedef ACKernelAdaptor<fundamental::SevenPointSolver, fundamental::SimpleError, UnnormalizerT, Mat3> KernelF;
  typedef ACKernelAdaptor<homography::FourPointSolver, homography::AsymmetricError, UnnormalizerI, Mat3> KernelH;
  if (model == 'h') return run_acransac<KernelH>(xI, xJ, n, wI, hI, wJ, hJ, false, precision, iterations, seed, inliers, out, skip, verbose != 0);
  return run_acransac<KernelF>(xI, xJ, n, wI, hI, wJ, hJ, true, precision, iterations, seed, inliers, out, skip, verbose != 0);
}

// GeometricFilter_HMatrix_AC::Fit's kernel (homography_acransac.h:35-46): point-to-point.
int ref_acransac_h(const float* xI, const float* xJ, int n, int wI, int hI, int wJ, int hJ, double precision, int iterations,
                   unsigned seed, int* inliers, double* out) {
  typedef ACKernelAdaptor<homography::FourPointSolver, homography::AsymmetricError, UnnormalizerI, Mat3> KernelType;
  return run_acransac<KernelType>(xI, xJ, n, wI, hI, wJ, hJ, false, precision, iterations, seed, inliers, out);
}

}  // extern "C"
