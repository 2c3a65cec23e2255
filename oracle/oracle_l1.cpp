// oracle_l1.cpp -- ORACLE L1: independent CPU restatement of the reference's brute-force putative
// matching path.  TEST INFRASTRUCTURE ONLY: nothing under 3dreconstruction_b200/ may include, link or
// call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// Parity status: PINNED.  Every function below is checked in tests/test_oracle.py against (a) the
// reference's own known-answer tests (metric_unittest.cpp:19-35, matching_unittest.cpp:13-74,
// indexed_match_unittest.cpp:6-54), (b) oracle L0 = the reference's own headers compiled by
// oracle/build_ref.sh, on random / tie-heavy / real (data/et) inputs, and (c) the golden fixtures in
// tests/golden/ that L0 generated (sha256 of matches.putative.txt at ratio 0.6 and 0.8).
//
// What is restated (reference file:line, relative to the reference tree):
//   l1_sqdist        SquaredEuclideanDistanceVectorized<uchar>      libs/feature/include/mvg/feature/metric.h:51-82
//                    (float accumulation of exact integers < 2^24 == exact int32)
//   l1_knn2          ArrayMatcherBruteForce::SearchNeighbours k=2   matcher_brute_force.h:102-134
//                    + SortIndexHelper/std::partial_sort tie rule   libs/base/include/mvg/utils/indexed_sort.h:52-66
//   l1_ratio_pass    DistanceRatioFilter                            matching_filters.h:27-47, numeric.h:108-111
//   l1_pair_matches  drop-last loop + IndexedMatch::getDeduplicated matcher_all_in_memory.h:102-125, indexed_match.h:39-55
//   l1_dedup_xy      IndexedMatchDecorator<float>::getDeduplicated  indexed_match_decorator.h:33-53,90-104
//   l1_export_text   PairedIndexedMatchToStream                     indexed_match_utils.h:22-38
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <set>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const int kDim = 128;

// Exact squared L2 distance of two 128-byte rows.  The reference accumulates float(a-b)^2 in float
// (metric.h:57-81); every partial sum is an integer <= 128*255^2 = 8,323,200 < 2^24, hence exact, hence
// equal to this int32 sum regardless of summation order.
__attribute__((target_clones("avx2", "default")))
int sqdist128(const uint8_t* a, const uint8_t* b) {
  int s = 0;
  for (int k = 0; k < kDim; ++k) {
    const int d = (int)a[k] - (int)b[k];
    s += d * d;
  }
  return s;
}

__attribute__((target_clones("avx2", "default")))
void sqdist_row(const uint8_t* q, const uint8_t* db, int n, int* out) {
  for (int r = 0; r < n; ++r) {
    const uint8_t* b = db + (size_t)r * kDim;
    int s = 0;
    for (int k = 0; k < kDim; ++k) {
      const int d = (int)q[k] - (int)b[k];
      s += d * d;
    }
    out[r] = s;
  }
}

// Which two packets std::partial_sort(begin, begin+2, end) with operator< on the value only leaves in
// front (libstdc++ __heap_select + __sort_heap).  Not "lowest index wins": e.g. d = [5,9,5] -> nearest is
// index 2.  (verified against L0 in tests/test_oracle.py.)
void top2_reference(const int* d, int n, int& s_out, int& t_out) {
  int T, S;
  if (d[1] < d[0]) { T = 0; S = 1; } else { T = 1; S = 0; }
  for (int v = 2; v < n; ++v) {
    if (d[v] < d[T]) {
      if (d[S] < d[v]) { T = v; } else { T = S; S = v; }
    }
  }
  s_out = S;
  t_out = T;
}

void top2_lowest_index(const int* d, int n, int& s_out, int& t_out) {
  int S = -1, T = -1;
  for (int v = 0; v < n; ++v) {
    if (S < 0 || d[v] < d[S]) { T = S; S = v; }
    else if (T < 0 || d[v] < d[T]) { T = v; }
  }
  s_out = S;
  t_out = T;
}

struct Decorated {
  float x1, y1, x2, y2;
  int i, j;
};
inline bool deco_eq(const Decorated& a, const Decorated& b) {
  return a.x1 == b.x1 && a.y1 == b.y1 && a.x2 == b.x2 && a.y2 == b.y2;
}
struct DecoLess {  // indexed_match_decorator.h:33-45 -- deliberately NOT a strict weak order
  bool operator()(const Decorated& a, const Decorated& b) const {
    if (deco_eq(a, b)) return false;
    if (a.x1 < b.x1) return a.y1 < b.y1;
    else if (a.x1 > b.x1) return a.y1 < b.y1;
    return a.x1 < b.x1;
  }
};

}  // namespace

extern "C" {

int l1_sqdist(const uint8_t* a, const uint8_t* b, int n) {
  int s = 0;
  for (int k = 0; k < n; ++k) { const int d = (int)a[k] - (int)b[k]; s += d * d; }
  return s;
}

// tie_mode 0 = lowest index, 1 = reference (partial_sort machine).  idx/dist are [nq][2].
// Returns 0 ("Too much asked nearest neighbors", matcher_brute_force.h:107-110) when db_rows < 2 or nq < 1.
int l1_knn2(const uint8_t* db, int db_rows, const uint8_t* q, int nq, int tie_mode, int32_t* idx, int32_t* dist) {
  if (db_rows < 2 || nq < 1) return 0;
#ifdef _OPENMP
#pragma omp parallel
#endif
  {
    std::vector<int> d(db_rows);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int qi = 0; qi < nq; ++qi) {
      sqdist_row(q + (size_t)qi * kDim, db, db_rows, d.data());
      int S, T;
      if (tie_mode == 1) top2_reference(d.data(), db_rows, S, T); else top2_lowest_index(d.data(), db_rows, S, T);
      idx[2 * qi] = S; idx[2 * qi + 1] = T;
      dist[2 * qi] = d[S]; dist[2 * qi + 1] = d[T];
    }
  }
  return 1;
}

// DistanceRatioFilter: all three operands fp32, one fp32 multiply, strict <.
int l1_ratio_pass(int d1, int d2, float ratio_sq) {
  volatile float rhs = ratio_sq * (float)d2;  // volatile: no excess precision / contraction
  return (float)d1 < rhs ? 1 : 0;
}

// knn2 + ratio filter + drop-last + unique-on-_i.  out: (_i,_j) pairs (capacity nq), returns count.
int l1_pair_matches(const uint8_t* db, int db_rows, const uint8_t* q, int nq, float ratio_sq, int32_t* out) {
  if (db_rows < 2 || nq < 1) return 0;  // SearchNeighbours returns false, outputs stay empty
  std::vector<int32_t> idx(2 * (size_t)nq), dist(2 * (size_t)nq);
  l1_knn2(db, db_rows, q, nq, 1, idx.data(), dist.data());
  std::vector<int> pass;
  for (int qi = 0; qi < nq; ++qi)
    if (l1_ratio_pass(dist[2 * qi], dist[2 * qi + 1], ratio_sq)) pass.push_back(qi);
  int n = 0;
  // matcher_all_in_memory.h:117: `k < size()-1 && size() > 0` -- the LAST passing query is dropped
  for (size_t k = 0; k + 1 < pass.size(); ++k) {
    const int i = idx[2 * pass[k]], j = pass[k];
    // IndexedMatch::getDeduplicated on this ascending-_j sequence == keep iff _i differs from the previous kept _i
    if (n > 0 && out[2 * (n - 1)] == i) continue;
    out[2 * n] = i;
    out[2 * n + 1] = j;
    ++n;
  }
  return n;
}

// IndexedMatch::getDeduplicated restated for ascending-_j input (what the path produces).
int l1_dedup_indexed_sorted(int32_t* m, int n) {
  int w = 0;
  for (int k = 0; k < n; ++k) {
    if (w > 0 && m[2 * (w - 1)] == m[2 * k]) continue;
    m[2 * w] = m[2 * k];
    m[2 * w + 1] = m[2 * k + 1];
    ++w;
  }
  return w;
}

// Coordinate de-duplication; feats are [rows][2] (x,y).  In place, returns the new count.
int l1_dedup_xy(int32_t* m, int n, const float* fI, const float* fJ) {
  std::vector<Decorated> v(n);
  for (int k = 0; k < n; ++k) {
    const int I = m[2 * k], J = m[2 * k + 1];
    Decorated d = {fI[2 * I], fI[2 * I + 1], fJ[2 * J], fJ[2 * J + 1], I, J};
    v[k] = d;
  }
  std::set<Decorated, DecoLess> s(v.begin(), v.end());  // same container + range ctor as the reference (:93-95)
  int w = 0;
  for (std::set<Decorated, DecoLess>::const_iterator it = s.begin(); it != s.end(); ++it, ++w) {
    m[2 * w] = it->i;
    m[2 * w + 1] = it->j;
  }
  return w;
}

// Full collection: for each pair (I,J) rows 7-13.  descs[i] -> [rows[i]][128]; feats[i] -> [rows[i]][2] or NULL
// (then de-dup-2 is skipped).  Results are appended to a caller-sized arena: out_counts[n_pairs],
// out_matches (capacity sum of q rows).  Returns total matches.
long long l1_match_collection(const uint8_t* const* descs, const float* const* feats, const int32_t* rows,
                              const int32_t* pairs, int n_pairs, float ratio_sq, int32_t* out_counts,
                              int32_t* out_matches) {
  long long total = 0;
  for (int p = 0; p < n_pairs; ++p) {
    const int I = pairs[2 * p], J = pairs[2 * p + 1];
    int32_t* out = out_matches + 2 * total;
    int n = l1_pair_matches(descs[I], rows[I], descs[J], rows[J], ratio_sq, out);
    if (feats && feats[I] && feats[J]) n = l1_dedup_xy(out, n, feats[I], feats[J]);
    out_counts[p] = n;
    total += n;
  }
  return total;
}

// PairedIndexedMatchToStream: pairs must already be unique and in lexicographic order (std::map iteration).
int l1_export_text(const int32_t* pairs, const int32_t* counts, const int32_t* matches, int n_pairs, const char* path) {
  FILE* f = fopen(path, "wb");
  if (!f) return 0;
  long long off = 0;
  for (int p = 0; p < n_pairs; ++p) {
    fprintf(f, "%d %d\n%d\n", pairs[2 * p], pairs[2 * p + 1], counts[p]);
    for (int k = 0; k < counts[p]; ++k, ++off) fprintf(f, "%d %d\n", matches[2 * off], matches[2 * off + 1]);
  }
  return fclose(f) == 0;
}

// CPU-baseline timing leg ("port"): knn2 + ratio test over n_pairs pairs; returns total passing queries.
long long l1_bench_bf(const uint8_t* const* dbs, const int32_t* db_rows, const uint8_t* const* qs, const int32_t* q_rows,
                      int n_pairs, float ratio_sq) {
  long long total = 0;
  for (int p = 0; p < n_pairs; ++p) {
    std::vector<int32_t> idx(2 * (size_t)q_rows[p]), dist(2 * (size_t)q_rows[p]);
    if (!l1_knn2(dbs[p], db_rows[p], qs[p], q_rows[p], 1, idx.data(), dist.data())) continue;
    for (int qi = 0; qi < q_rows[p]; ++qi) total += l1_ratio_pass(dist[2 * qi], dist[2 * qi + 1], ratio_sq);
  }
  return total;
}

int l1_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
