// stdsort_restated.cuh -- libstdc++'s std::sort (introsort) restated step by step over a (residual, index) array.
//
// Why: ACRANSAC sorts ALL residuals of a model with std::sort(vec_residuals) (estimator_acransac.h:176-181) under
// std::pair<double, size_t>::operator<.  A degenerate sample can give a model whose residual is NaN for some points
// (homography: 0 / 0 where H maps a point to (0, 0, 0); fundamental matrix: 0 / 0 where F x = 0).  With a NaN,
//     a < b  :=  a.e < b.e || (!(b.e < a.e) && a.i < b.i)
// is no strict weak ordering any more (a NaN entry is "equivalent" to every value and ordered against it by INDEX), so the
// result of std::sort is whatever the algorithm's steps leave -- typically NaN entries at the FRONT, which bestNFA never
// looks at (its scan starts at k = MINIMUM_SAMPLES + 1 and reads e[k - 1]), followed by tiny residuals: a very negative NFA
// and a "meaningful" model the filter then draws its samples from.  The reference binary's answer on such pairs therefore
// depends on the exact sequence of comparisons and moves; this file repeats that sequence (GCC 13 libstdc++,
// bits/stl_algo.h:1871-1950 + bits/stl_heap.h:135-430; the algorithm is unchanged since GCC 4.7's median-of-3 variant):
//     __sort              = __introsort_loop(first, last, 2 * floor(log2 n)) ; __final_insertion_sort
//     __introsort_loop    = while (last - first > 16): depth 0 -> heap sort of the range, return;
//                             cut = __unguarded_partition_pivot ; recurse on [cut, last) ; last = cut
//     __unguarded_partition_pivot = median of (first + 1, mid, last - 1) moved to first ; __unguarded_partition(first + 1, last, first)
//     __final_insertion_sort      = __insertion_sort on the first 16 ; __unguarded_linear_insert for the rest
// Only iterations whose residuals contain a NaN come here (acransac_kernels.cuh: model_candidates_warp); everything else
// uses the parallel sort of the candidates, which is the same list whenever the order is a strict weak one.
//
// The "unguarded" loops rely on the ordering to stop inside the range; without one they may leave it.  Inside the array they
// are followed as they are; at the ends of the ARRAY they stop here and set *left_array (the reference would read
// whatever lies next to the vector: nothing to reproduce).
#pragma once
#include "acransac_core.cuh"

namespace mvgcuda {
namespace geo {

#if defined(MVG_SORT_STATS) && !defined(__CUDA_ARCH__)
static long g_ss_heap_sorts = 0;
#endif

struct SortArr {
  double* e;
  int* i;
  int n;
  int left_array;
};

MVG_GEO_HD bool ss_less_v(double ae, int ai, double be, int bi) { return ae < be || (!(be < ae) && ai < bi); }
MVG_GEO_HD bool ss_less(const SortArr& A, int a, int b) { return ss_less_v(A.e[a], A.i[a], A.e[b], A.i[b]); }
MVG_GEO_HD void ss_swap(SortArr& A, int a, int b) {
  const double te = A.e[a]; A.e[a] = A.e[b]; A.e[b] = te;
  const int ti = A.i[a]; A.i[a] = A.i[b]; A.i[b] = ti;
}

// stl_heap.h:135-150
MVG_GEO_HD void ss_push_heap(SortArr& A, int first, int hole, int top, double ve, int vi) {
  int parent = (hole - 1) / 2;
  while (hole > top && ss_less_v(A.e[first + parent], A.i[first + parent], ve, vi)) {
    A.e[first + hole] = A.e[first + parent]; A.i[first + hole] = A.i[first + parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  A.e[first + hole] = ve; A.i[first + hole] = vi;
}
// stl_heap.h:224-250
MVG_GEO_HD void ss_adjust_heap(SortArr& A, int first, int hole, int len, double ve, int vi) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (ss_less(A, first + child, first + child - 1)) --child;
    A.e[first + hole] = A.e[first + child]; A.i[first + hole] = A.i[first + child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    A.e[first + hole] = A.e[first + child - 1]; A.i[first + hole] = A.i[first + child - 1];
    hole = child - 1;
  }
  ss_push_heap(A, first, hole, top, ve, vi);
}
// __partial_sort(first, last, last): __make_heap (stl_heap.h:340-362) then __sort_heap (:419-427)
MVG_GEO_HD void ss_heap_sort(SortArr& A, int first, int last) {
  const int len = last - first;
  if (len >= 2) {
    int parent = (len - 2) / 2;
    while (true) {
      ss_adjust_heap(A, first, parent, len, A.e[first + parent], A.i[first + parent]);
      if (parent == 0) break;
      --parent;
    }
  }
  while (last - first > 1) {
    --last;
    const double ve = A.e[last]; const int vi = A.i[last];   // __pop_heap(first, last, last)
    A.e[last] = A.e[first]; A.i[last] = A.i[first];
    ss_adjust_heap(A, first, 0, last - first, ve, vi);
  }
}

// stl_algo.h:85-103
MVG_GEO_HD void ss_move_median_to_first(SortArr& A, int result, int a, int b, int c) {
  if (ss_less(A, a, b)) {
    if (ss_less(A, b, c)) ss_swap(A, result, b);
    else if (ss_less(A, a, c)) ss_swap(A, result, c);
    else ss_swap(A, result, a);
  } else if (ss_less(A, a, c)) ss_swap(A, result, a);
  else if (ss_less(A, b, c)) ss_swap(A, result, c);
  else ss_swap(A, result, b);
}
// stl_algo.h:1871-1888
MVG_GEO_HD int ss_unguarded_partition(SortArr& A, int first, int last, int pivot) {
  while (true) {
    while (first < A.n && ss_less(A, first, pivot)) ++first;
    if (first >= A.n) { A.left_array = 1; return A.n; }
    --last;
    while (last >= 0 && ss_less(A, pivot, last)) --last;
    if (last < 0) { A.left_array = 1; return first; }
    if (!(first < last)) return first;
    ss_swap(A, first, last);
    ++first;
  }
}
// stl_algo.h:1792-1807
MVG_GEO_HD void ss_unguarded_linear_insert(SortArr& A, int last) {
  const double ve = A.e[last]; const int vi = A.i[last];
  int next = last - 1;
  while (next >= 0 && ss_less_v(ve, vi, A.e[next], A.i[next])) {
    A.e[last] = A.e[next]; A.i[last] = A.i[next];
    last = next;
    --next;
  }
  if (next < 0) A.left_array = 1;  // still "less" than element 0: the reference goes on below the vector
  A.e[last] = ve; A.i[last] = vi;
}
// stl_algo.h:1812-1830
MVG_GEO_HD void ss_insertion_sort(SortArr& A, int first, int last) {
  if (first == last) return;
  for (int k = first + 1; k != last; ++k) {
    if (ss_less(A, k, first)) {
      const double ve = A.e[k]; const int vi = A.i[k];
      for (int q = k; q > first; --q) { A.e[q] = A.e[q - 1]; A.i[q] = A.i[q - 1]; }  // move_backward(first, k, k + 1)
      A.e[first] = ve; A.i[first] = vi;
    } else {
      ss_unguarded_linear_insert(A, k);
    }
  }
}

// std::sort(vec_residuals.begin(), vec_residuals.end()) -- sequential; e / idx are reordered in place.
MVG_GEO_HD int libstdcxx_sort(double* e, int* idx, int n) {
  SortArr A{e, idx, n, 0};
  if (n < 1) return 0;
  int lg = 0;
  while ((n >> (lg + 1)) > 0) ++lg;  // std::__lg
  // __introsort_loop with its recursion on the right part unrolled onto a stack: a frame is one pending call
  // __introsort_loop(first, last, depth); the call on [cut, last) runs to completion before [first, cut) continues.
  int st_first[72], st_last[72], st_depth[72];
  int sp = 0;
  st_first[0] = 0; st_last[0] = n; st_depth[0] = 2 * lg; sp = 1;
  while (sp > 0) {
    --sp;
    int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
    while (last - first > 16) {
      if (depth == 0) {
#if defined(MVG_SORT_STATS) && !defined(__CUDA_ARCH__)
        ++g_ss_heap_sorts;  // (tests: how often the depth limit was reached)
#endif
        ss_heap_sort(A, first, last);
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      ss_move_median_to_first(A, first, first + 1, mid, last - 1);
      const int cut = ss_unguarded_partition(A, first + 1, last, first);
      if (sp < 72) { st_first[sp] = first; st_last[sp] = cut; st_depth[sp] = depth; ++sp; }  // continues after the right part
      first = cut;
    }
  }
  // __final_insertion_sort
  if (n > 16) {
    ss_insertion_sort(A, 0, 16);
    for (int k = 16; k != n; ++k) ss_unguarded_linear_insert(A, k);
  } else {
    ss_insertion_sort(A, 0, n);
  }
  return A.left_array;
}

// What bestNFA's scan sees of a sorted residual array (estimator_acransac.h:86): k runs from sample + 1 while
// e[k - 1] <= max_threshold -- the first `sample` entries are never looked at.  Returns the last such k (sample if none).
MVG_GEO_HD int nfa_scan_end(const double* e, int n, int sample, double max_threshold) {
  int k = sample + 1;
  while (k <= n && e[k - 1] <= max_threshold) ++k;
  return k - 1;
}

}  // namespace geo
}  // namespace mvgcuda
