// ref_driver.cpp -- ORACLE L0 (test infrastructure, NOT product code).
//
// A thin extern "C" driver around the REFERENCE'S OWN headers, compiled from where they lie under
// /root/reference by oracle/build_ref.sh into oracle/_ref/ (git-ignored).  Nothing of the reference is
// copied into this repository; this file only calls it:
//   ArrayMatcherBruteForce<uchar, SquaredEuclideanDistanceVectorized<uchar>>   matcher_brute_force.h:42-134
//   DistanceRatioFilter                                                       matching_filters.h:27-47
//   IndexedMatch::getDeduplicated                                             indexed_match.h:49-55
//   IndexedMatchDecorator<float>::getDeduplicated                             indexed_match_decorator.h:60-104
//   MatcherAllInMemory<...>::LoadData / Match                                 matcher_all_in_memory.h:44-141
//   PairedIndexedMatchToStream / pairedIndexedMatchImport                     indexed_match_utils.h:22-73
//   ArrayMatcherKdtreeFlann<uchar, flann::L2<uchar>> (CPU baseline only)      matcher_kdtree_flann.h:15-128
//   Square                                                                    numeric.h:108-111
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the
// resulting library.
//
// Prelude required by the reference headers under g++ (SURVEY.md 8(c) recipe): they rely on MSVC's
// lax two-phase lookup for ControlProgressDisplay / make_pair.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "mvg/utils/progress.h"
using namespace std;
using namespace mvg::utils;

#include "mvg/feature/features.h"
#include "mvg/feature/indexed_match_utils.h"
#include "mvg/feature/matcher_all_in_memory.h"
#include "mvg/feature/matcher_brute_force.h"  // the typename-patched temp copy, earlier on the include path
#include "mvg/feature/matching_filters.h"
#ifdef ORACLE_WITH_FLANN
#include "mvg/feature/matcher_kdtree_flann.h"
#endif

using namespace mvg::feature;

typedef Descriptor<unsigned char, 128> DescriptorT;                    // compute_matches.cpp:182
typedef ScalePointFeature FeatureT;                                    // :183
typedef KeypointSet<std::vector<FeatureT>, std::vector<DescriptorT> > KeypointSetT;
typedef SquaredEuclideanDistanceVectorized<unsigned char> MetricT;     // :226
typedef ArrayMatcherBruteForce<unsigned char, MetricT> MatcherBF;      // :227

extern "C" {

int ref_has_flann() {
#ifdef ORACLE_WITH_FLANN
  return 1;
#else
  return 0;
#endif
}
int ref_has_openmp() {
#ifdef USE_OPENMP
  return 1;
#else
  return 0;
#endif
}
// Team size of the bench loops below.  (The environment is only read when libgomp loads; torchrun exports
// OMP_NUM_THREADS=1 to its workers, which silently made the "all cores" baseline single-threaded in round 1.)
int ref_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

// metric.h:51-82 on two rows
float ref_metric(const unsigned char* a, const unsigned char* b, int n) { return MetricT()(a, b, (size_t)n); }

// numeric.h:108-111
float ref_square(float r) { return mvg::math::Square(r); }

// Build + SearchNeighbours(k).  Returns 1 on success (outputs hold nq*k entries), 0 if the reference returned false.
int ref_knn(const unsigned char* db, int db_rows, const unsigned char* q, int nq, int k, int* idx, float* dist) {
  MatcherBF m;
  if (!m.Build(db, db_rows, 128)) return 0;
  std::vector<int> vi;
  std::vector<float> vd;
  if (!m.SearchNeighbours(q, nq, &vi, &vd, (size_t)k)) return 0;
  std::copy(vi.begin(), vi.end(), idx);
  std::copy(vd.begin(), vd.end(), dist);
  return 1;
}

// DistanceRatioFilter over dist[n] laid out [q][nn]; ratio is the (already squared) fp32 threshold.
int ref_ratio_filter(const float* dist, int n, int nn, float ratio, int* out) {
  std::vector<float> v(dist, dist + n);
  std::vector<int> keep;
  DistanceRatioFilter(v.begin(), v.end(), nn, keep, ratio);
  std::copy(keep.begin(), keep.end(), out);
  return (int)keep.size();
}

// The per-pair sequence of matcher_all_in_memory.h:102-129 WITHOUT the coordinate de-dup: knn2 + ratio filter +
// drop-last loop + IndexedMatch::getDeduplicated.  out holds (_i,_j) pairs; returns their number.
int ref_pair_matches(const unsigned char* db, int db_rows, const unsigned char* q, int nq, float ratio_sq, int* out) {
  MatcherBF m;
  m.Build(db, db_rows, 128);
  std::vector<int> vi;
  std::vector<float> vd;
  if (db_rows >= 1) m.SearchNeighbours(q, nq, &vi, &vd, 2);
  std::vector<int> pass;
  DistanceRatioFilter(vd.begin(), vd.end(), 2, pass, ratio_sq);
  std::vector<IndexedMatch> ms;
  for (size_t k = 0; k < pass.size() - 1 && pass.size() > 0; ++k) {
    const size_t index = pass[k];
    ms.push_back(IndexedMatch(vi[index * 2], index));
  }
  IndexedMatch::getDeduplicated(ms);
  for (size_t k = 0; k < ms.size(); ++k) { out[2 * k] = (int)ms[k]._i; out[2 * k + 1] = (int)ms[k]._j; }
  return (int)ms.size();
}

// IndexedMatch::getDeduplicated on n (_i,_j) pairs, in place; returns the new count.
int ref_dedup_indexed(int* m, int n) {
  std::vector<IndexedMatch> v;
  for (int k = 0; k < n; ++k) v.push_back(IndexedMatch(m[2 * k], m[2 * k + 1]));
  IndexedMatch::getDeduplicated(v);
  for (size_t k = 0; k < v.size(); ++k) { m[2 * k] = (int)v[k]._i; m[2 * k + 1] = (int)v[k]._j; }
  return (int)v.size();
}

// IndexedMatchDecorator<float>(matches, featsI, featsJ).getDeduplicated; feats are [rows][2] (x,y).
int ref_dedup_xy(int* m, int n, const float* fI, int nI, const float* fJ, int nJ) {
  std::vector<FeatureT> a, b;
  for (int k = 0; k < nI; ++k) a.push_back(FeatureT(fI[2 * k], fI[2 * k + 1]));
  for (int k = 0; k < nJ; ++k) b.push_back(FeatureT(fJ[2 * k], fJ[2 * k + 1]));
  std::vector<IndexedMatch> v;
  for (int k = 0; k < n; ++k) v.push_back(IndexedMatch(m[2 * k], m[2 * k + 1]));
  IndexedMatchDecorator<float> deco(v, a, b);
  deco.getDeduplicated(v);
  for (size_t k = 0; k < v.size(); ++k) { m[2 * k] = (int)v[k]._i; m[2 * k + 1] = (int)v[k]._j; }
  return (int)v.size();
}

// The whole collection path: MatcherAllInMemory(distRatio).LoadData(names, dir) + Match + PairedIndexedMatchToStream
// into out_path (compute_matches.cpp:237-246).  names = '\n'-separated image file names.  The progress bar goes to
// stdout (silenced).  Returns 1 on success.
int ref_match_dir(const char* match_dir, const char* names, float dist_ratio, const char* out_path) {
  std::vector<std::string> file_names;
  {
    std::istringstream ss(names);
    std::string line;
    while (std::getline(ss, line))
      if (!line.empty()) file_names.push_back(line);
  }
  MatcherAllInMemory<KeypointSetT, MatcherBF> coll(dist_ratio);
  if (!coll.LoadData(file_names, match_dir)) return 0;
  PairWiseMatches map_putatives;
  std::streambuf* old = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());
  coll.Match(file_names, map_putatives);
  std::cout.rdbuf(old);
  std::ofstream file(out_path);
  if (!file.is_open()) return 0;
  PairedIndexedMatchToStream(map_putatives, file);
  file.close();
  return 1;
}

// import + re-export through the reference's own reader/writer (file-boundary acceptance check, config 4)
int ref_roundtrip_matches(const char* in_path, const char* out_path) {
  PairWiseMatches m;
  std::streambuf* old = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());
  const bool ok = pairedIndexedMatchImport(in_path, m);
  std::cout.rdbuf(old);
  if (!ok) return 0;
  std::ofstream file(out_path);
  if (!file.is_open()) return 0;
  PairedIndexedMatchToStream(m, file);
  return 1;
}

// CPU baseline timing helper: knn2 + ratio filter for `n_pairs` (db, q) pairs, parallel over pairs when built
// with OpenMP (the reference parallelises over j for fixed i, matcher_all_in_memory.h:87-90).  Returns the total
// number of passing queries (so the work cannot be optimised away).
long long ref_bench_bf(const unsigned char* const* dbs, const int* db_rows, const unsigned char* const* qs,
                       const int* q_rows, int n_pairs, float ratio_sq) {
  long long total = 0;
#ifdef USE_OPENMP
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
#endif
  for (int p = 0; p < n_pairs; ++p) {
    MatcherBF m;
    m.Build(dbs[p], db_rows[p], 128);
    std::vector<int> vi;
    std::vector<float> vd;
    m.SearchNeighbours(qs[p], q_rows[p], &vi, &vd, 2);
    std::vector<int> pass;
    DistanceRatioFilter(vd.begin(), vd.end(), 2, pass, ratio_sq);
    total += (long long)pass.size();
  }
  return total;
}

#ifdef ORACLE_WITH_FLANN
typedef ArrayMatcherKdtreeFlann<unsigned char, flann::L2<unsigned char> > MatcherFlann;  // compute_matches.cpp:222-223
long long ref_bench_flann(const unsigned char* const* dbs, const int* db_rows, const unsigned char* const* qs,
                          const int* q_rows, int n_pairs, float ratio_sq) {
  long long total = 0;
#ifdef USE_OPENMP
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
#endif
  for (int p = 0; p < n_pairs; ++p) {
    MatcherFlann m;
    m.Build(dbs[p], db_rows[p], 128);
    std::vector<int> vi;
    std::vector<float> vd;
    m.SearchNeighbours(qs[p], q_rows[p], &vi, &vd, 2);
    std::vector<int> pass;
    DistanceRatioFilter(vd.begin(), vd.end(), 2, pass, ratio_sq);
    total += (long long)pass.size();
  }
  return total;
}
#endif

}  // extern "C"
