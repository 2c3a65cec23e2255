// test_adaptors.cpp -- the C++ drop-in adaptors (include/mvgcuda/*.h) against the REFERENCE'S OWN classes, in one
// binary: ArrayMatcherCuda vs ArrayMatcherBruteForce, MatcherCudaAllInMemory vs MatcherAllInMemory,
// ImageCollectionGeometricFilterCuda vs ImageCollectionGeometricFilter (GeometricFilter_FMatrix_AC, images of 640 x 480).
// Built by tests/native/build_native.sh where the reference tree is mounted (it needs the reference headers); the
// binary travels to the GPU box.  Usage: test_adaptors <match_dir> <name1> <name2> ...   (names of images whose
// .feat/.desc live in match_dir).  Exit code 0 and "ADAPTORS OK" on success.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "mvg/utils/progress.h"
using namespace std;
using namespace mvg::utils;

#include "mvg/feature/features.h"
#include "mvg/feature/indexed_match_utils.h"
#include "mvg/feature/matcher_all_in_memory.h"
#include "mvg/feature/matcher_brute_force.h"
#include "mvg/feature/two_view_matches.h"
#include "mvg/feature/geometric_filter.h"
#include "mvg/multiview/fundamental_acransac.h"
#include "mvg/multiview/homography_acransac.h"
#include "mvgcuda/array_matcher_cuda.h"
#include "mvgcuda/geometric_filter_cuda.h"
#include "mvgcuda/matcher_cuda_all_in_memory.h"

using namespace mvg::feature;
typedef Descriptor<unsigned char, 128> DescriptorT;
typedef ScalePointFeature FeatureT;
typedef KeypointSet<std::vector<FeatureT>, std::vector<DescriptorT> > KeypointSetT;
typedef SquaredEuclideanDistanceVectorized<unsigned char> MetricT;
typedef ArrayMatcherBruteForce<unsigned char, MetricT> MatcherBF;

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::cerr << "CHECK failed: " #c " (line " << __LINE__ << ")" << std::endl; ++fails; } } while (0)

static std::vector<unsigned char> rnd(int rows, int alphabet, unsigned seed) {
  std::mt19937 g(seed);
  std::vector<unsigned char> v((size_t)rows * 128);
  for (auto& b : v) b = (unsigned char)(g() % alphabet);
  return v;
}

// the two-view entry GetPutativesMatches<DescriptorT, MatcherT> (two_view_matches.h:9-47, used by
// apps/sift_match/sift_match.cpp:89-99) with the CUDA matcher as MatcherT, against the brute-force matcher.
static void two_view_entry(int nL, int nR, int alphabet, unsigned seed, float ratio) {
  std::vector<unsigned char> l = rnd(nL, alphabet, seed), r = rnd(nR, alphabet, seed + 1);
  std::vector<DescriptorT> L(nL), R(nR);
  for (int i = 0; i < nL; ++i) memcpy(L[i].getData(), l.data() + (size_t)i * 128, 128);
  for (int i = 0; i < nR; ++i) memcpy(R[i].getData(), r.data() + (size_t)i * 128, 128);
  std::vector<IndexedMatch> m0, m1;
  GetPutativesMatches<DescriptorT, MatcherBF>(L, R, mvg::math::Square(ratio), m0);
  GetPutativesMatches<DescriptorT, ArrayMatcherCuda<unsigned char, MetricT> >(L, R, mvg::math::Square(ratio), m1);
  CHECK(m0.size() == m1.size());
  bool same = m0.size() == m1.size();
  for (size_t k = 0; same && k < m0.size(); ++k) same = (m0[k] == m1[k]);
  CHECK(same);
  std::cout << "two-view " << nL << "x" << nR << " ratio " << ratio << ": " << m0.size() << " matches, identical=" << same << std::endl;
}

static void array_level(int db_rows, int nq, int alphabet, unsigned seed) {
  std::vector<unsigned char> db = rnd(db_rows, alphabet, seed), q = rnd(nq, alphabet, seed + 1);
  MatcherBF ref;
  ArrayMatcherCuda<unsigned char, MetricT> gpu;
  ArrayMatcher<unsigned char, MetricT>* base = &gpu;   // used through the reference's abstract interface
  CHECK(ref.Build(db.data(), db_rows, 128) == base->Build(db.data(), db_rows, 128));
  std::vector<int> i0(1, 42), i1(1, 42);
  std::vector<float> d0(1, 7.f), d1(1, 7.f);   // pre-filled: both APPEND
  const bool r0 = ref.SearchNeighbours(q.data(), nq, &i0, &d0, 2);
  const bool r1 = base->SearchNeighbours(q.data(), nq, &i1, &d1, 2);
  CHECK(r0 == r1);
  CHECK(i0 == i1);
  CHECK(d0 == d1);
  if (i0 != i1) std::cerr << "  array_level(" << db_rows << "," << nq << "," << alphabet << ") indices differ" << std::endl;
}

// k = 1 under heavy ties (tiny alphabet, duplicated rows): ArrayMatcherBruteForce runs partial_sort(first, first + 1, last),
// which keeps the FIRST minimum; and the reference's usage pattern Build once / search many times (matcher_all_in_memory.h:84-107)
static void array_level_k1_and_reuse(int db_rows, int nq, int alphabet, unsigned seed) {
  std::vector<unsigned char> db = rnd(db_rows, alphabet, seed);
  for (int r = 1; r < db_rows; r += 3) std::copy(db.begin(), db.begin() + 128, db.begin() + (size_t)r * 128);  // many exact duplicates
  MatcherBF ref;
  ArrayMatcherCuda<unsigned char, MetricT> gpu;
  CHECK(ref.Build(db.data(), db_rows, 128) == gpu.Build(db.data(), db_rows, 128));
  for (int round = 0; round < 3; ++round) {  // the db stays in HBM: three searches, one upload of it
    std::vector<unsigned char> q = rnd(nq, alphabet, seed + 10 + round);
    std::copy(db.begin(), db.begin() + 128, q.begin());  // a query equal to the duplicated row: distance-0 ties
    for (size_t k : {size_t(1), size_t(2)}) {
      std::vector<int> i0, i1;
      std::vector<float> d0, d1;
      CHECK(ref.SearchNeighbours(q.data(), nq, &i0, &d0, k) == gpu.SearchNeighbours(q.data(), nq, &i1, &d1, k));
      CHECK(d0 == d1);
      CHECK(i0 == i1);
      if (i0 != i1) std::cerr << "  k=" << k << " round " << round << ": indices differ under ties" << std::endl;
    }
  }
}

int main(int argc, char** argv) {
  // array level: random, tie-heavy, degenerate
  array_level_k1_and_reuse(600, 300, 2, 31);
  array_level_k1_and_reuse(257, 129, 3, 32);
  array_level(300, 200, 256, 1);
  array_level(1000, 513, 3, 2);
  array_level(2, 5, 2, 3);
  array_level(1, 4, 256, 4);   // k=2 > rows: both return false, outputs untouched
  two_view_entry(700, 650, 4, 21, 0.8f);
  two_view_entry(1500, 1200, 256, 22, 0.95f);
  {
    ArrayMatcherCuda<unsigned char, MetricT> gpu;
    CHECK(!gpu.Build(NULL, 0, 128));
    std::vector<unsigned char> db = rnd(50, 256, 9), q = rnd(1, 256, 10);
    CHECK(gpu.Build(db.data(), 50, 128));
    int idx = -1; float dist = -1.f;
    CHECK(gpu.SearchNeighbour(q.data(), &idx, &dist));
    MetricT m; float best = 1e30f; int bi = -1;
    for (int r = 0; r < 50; ++r) { float d = m(q.data(), db.data() + r * 128, 128); if (d < best) { best = d; bi = r; } }
    CHECK(idx == bi && dist == best);
  }
  // collection level on files
  if (argc >= 4) {
    const std::string dir = argv[1];
    std::vector<std::string> names;
    for (int k = 2; k < argc; ++k) names.push_back(argv[k]);
    for (float ratio : {0.6f, 0.8f}) {
      MatcherAllInMemory<KeypointSetT, MatcherBF> ref(ratio);
      MatcherCudaAllInMemory<KeypointSetT> gpu(ratio, 1);
      CHECK(ref.LoadData(names, dir));
      CHECK(gpu.LoadData(names, dir));
      PairWiseMatches m0, m1;
      std::streambuf* old = std::cout.rdbuf();
      std::ostringstream sink;
      std::cout.rdbuf(sink.rdbuf());
      ref.Match(names, m0);
      std::cout.rdbuf(old);
      const Matcher& as_base = gpu;   // through the reference's abstract Matcher interface
      as_base.Match(names, m1);
      std::ostringstream s0, s1;
      PairedIndexedMatchToStream(m0, s0);
      PairedIndexedMatchToStream(m1, s1);
      CHECK(m0.size() == m1.size());
      CHECK(s0.str() == s1.str());
      std::cout << "collection ratio " << ratio << ": " << m0.size() << " pairs, " << s0.str().size() << " bytes, identical="
                << (s0.str() == s1.str()) << std::endl;
      if (ratio == 0.6f) {
        // the step after (compute_matches.cpp:250-268): the reference's filter with its never-seeded rand() against the GPU one
        const std::vector<std::pair<size_t, size_t> > sizes(names.size(), std::make_pair((size_t)640, (size_t)480));
        ImageCollectionGeometricFilter<FeatureT> ref_filter;
        ImageCollectionGeometricFilterCuda<FeatureT> gpu_filter;
        CHECK(ref_filter.LoadData(names, dir));
        CHECK(gpu_filter.LoadData(names, dir));
        PairWiseMatches g0, g1;
        std::cout.rdbuf(sink.rdbuf());
        srand(1);
        ref_filter.Filter(mvg::multiview::GeometricFilter_FMatrix_AC(4.0), m0, g0, sizes);
        std::cout.rdbuf(old);
        gpu_filter.Filter(mvg::multiview::GeometricFilter_FMatrix_AC(4.0), m1, g1, sizes);
        std::ostringstream t0, t1;
        PairedIndexedMatchToStream(g0, t0);
        PairedIndexedMatchToStream(g1, t1);
        CHECK(g0.size() == g1.size());
        CHECK(t0.str() == t1.str());
        std::cout << "geometric filter (F, AC-RANSAC): " << g0.size() << " pairs kept, " << t0.str().size() << " bytes, identical="
                  << (t0.str() == t1.str()) << std::endl;
        PairWiseMatches h0, h1;
        std::cout.rdbuf(sink.rdbuf());
        srand(1);
        ref_filter.Filter(mvg::multiview::GeometricFilter_HMatrix_AC(4.0), m0, h0, sizes);
        std::cout.rdbuf(old);
        gpu_filter.Filter(mvg::multiview::GeometricFilter_HMatrix_AC(4.0), m1, h1, sizes);
        std::ostringstream u0, u1;
        PairedIndexedMatchToStream(h0, u0);
        PairedIndexedMatchToStream(h1, u1);
        CHECK(h0.size() == h1.size());
        CHECK(u0.str() == u1.str());
        std::cout << "geometric filter (H, AC-RANSAC): " << h0.size() << " pairs kept, " << u0.str().size() << " bytes, identical="
                  << (u0.str() == u1.str()) << std::endl;
      }
    }
  }
  std::cout << (fails ? "ADAPTORS FAILED" : "ADAPTORS OK") << std::endl;
  return fails ? 1 : 0;
}
