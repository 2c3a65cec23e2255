// acransac_engine.cuh -- the control flow of ACRANSAC (estimator_acransac.h:125-245) restated so that whole RANGES of
// iterations can be evaluated in parallel and then accounted for in order.
//
// The reference loop is sequential only through its state (minNFA, the best inlier set, vec_index, iter_num, nIterReserve):
//   * every iteration draws exactly MINIMUM_SAMPLES values from rand() (RandomSample, random_sampling.h:46-58), so iteration
//     `it` of a pair reads the stream at  pair_offset + 7 * it  whatever happened before;
//   * the sample is mapped through vec_index, which changes only when  better && minNFA < 0  (the inliers of the new best
//     model) or once at the end of the first phase (iter + 1 == iter_num && nIterReserve);
//   * a model is accepted only by the strict  best.first < minNFA.
// So all iterations of [iter, iter_num) can be evaluated SPECULATIVELY with the current vec_index; the first iteration that
// changes vec_index (the "trigger") invalidates only the evaluations after it, which are repeated with the new index set.
// In the first phase (no meaningful model yet, minNFA >= 0) the trigger is the first iteration whose best NFA is negative;
// afterwards it is the first one that improves minNFA.  With no trigger the state is the strict first-occurrence minimum.
//
// `Range` supplies two searches over the evaluated results of [lo, hi) -- sequential on the host (tests), warp-parallel on
// the device:
//     int  first_below(lo, hi, thr)      lowest it with nfa[it] < thr, or -1
//     int  argmin_first(lo, hi)          lowest it attaining the minimum nfa over the range (hi > lo)
//     double nfa(it); int n_inl(it); int model(it)
#pragma once
#include "acransac_core.cuh"

namespace mvgcuda {
namespace geo {

constexpr int kSampleF = 7;  // SevenPointSolver::MINIMUM_SAMPLES (fundamental matrix)
constexpr int kSampleH = 4;  // homography::FourPointSolver::MINIMUM_SAMPLES

struct AcState {
  double min_nfa;      // +inf: no model yet
  int best_it, best_model, n_inl;  // n_inl == 0: vec_inliers is empty
  int iter;            // first iteration not accounted for yet
  int iter_num;        // current loop bound
  int reserve;         // nIterReserve
  int index_it, index_model, n_index;  // vec_index = first n_index inliers of (index_it, index_model); index_it < 0: identity
  int extend_to;       // > 0: nothing found in the first phase, the loop grows one iteration at a time up to this bound
  int done;
  int sample;          // MINIMUM_SAMPLES of the model (7: fundamental matrix, 4: homography)
  int pad;
};

MVG_GEO_HD double ac_inf() { return __builtin_huge_val(); }

MVG_GEO_HD void ac_init(AcState& S, int n_data, int max_iterations, int sample_size = kSampleF) {
  S.sample = sample_size; S.pad = 0;
  S.min_nfa = ac_inf();
  S.best_it = -1; S.best_model = 0; S.n_inl = 0;
  S.iter = 0;
  S.reserve = max_iterations / 10;            // estimator_acransac.h:162-163
  S.iter_num = max_iterations - S.reserve;
  S.index_it = -1; S.index_model = 0; S.n_index = n_data;
  S.extend_to = 0;
  S.done = (n_data <= sample_size) ? 1 : 0;   // :137-138 (nothing is drawn from rand())
  if (S.done) S.iter_num = 0;
}

// Upper end of the range to evaluate next with the current vec_index.
MVG_GEO_HD int ac_range_end(const AcState& S) { return S.extend_to > 0 ? S.extend_to : S.iter_num; }

// Account for the evaluated range [S.iter, hi) -- hi = ac_range_end(S), or less (hi_eval) when only the head of the first
// phase has been evaluated so far.  Returns 1 when there is more to evaluate from S.iter (vec_index changed, or the rest of
// a partly evaluated range), 0 otherwise; S.done is set when the loop has ended.
template <typename Range>
MVG_GEO_HD int ac_account(AcState& S, const Range& R, int hi_eval = -1) {
  const int lo = S.iter, hi_full = ac_range_end(S);
  const int hi = (hi_eval >= 0 && hi_eval < hi_full) ? hi_eval : hi_full;
  if (S.extend_to > 0) {
    // The first phase ended with no model at all (vec_inliers empty): every further iteration that finds nothing does
    // iter_num++, nIterReserve-- (:223-226); the first one with ANY model makes its inliers the sampling set and spends what
    // is left of the reserve (:228-234), so the loop always ends at the same total.  (Never evaluated in parts.)
    const int t = lo < hi_full ? R.first_below(lo, hi_full, ac_inf()) : -1;
    const int total = S.extend_to;
    S.extend_to = 0;
    S.reserve = 0;
    S.iter_num = total;
    if (t < 0) { S.iter = total; S.done = 1; return 0; }
    S.min_nfa = R.nfa(t); S.best_it = t; S.best_model = R.model(t); S.n_inl = R.n_inl(t);
    S.iter = t + 1;
    // iteration total - 1 no longer has a reserve to spend: the sampling set changes only if its model is meaningful
    if (t + 1 == total) { S.done = 1; return 0; }
    S.index_it = t; S.index_model = S.best_model; S.n_index = S.n_inl;
    return 1;
  }
  if (lo >= hi) { S.done = 1; return 0; }
  const double thr = S.min_nfa < 0.0 ? S.min_nfa : 0.0;
  const int t = R.first_below(lo, hi, thr);
  if (t >= 0) {
    // iterations before t may have improved a non-negative minNFA without changing the sampling set; t itself is better
    // than all of them (negative) and becomes the state:  better && minNFA < 0  (:221-235)
    S.min_nfa = R.nfa(t); S.best_it = t; S.best_model = R.model(t); S.n_inl = R.n_inl(t);
    S.index_it = t; S.index_model = S.best_model; S.n_index = S.n_inl;
    if (S.reserve) { S.iter_num = t + 1 + S.reserve; S.reserve = 0; }
    S.iter = t + 1;
    if (S.iter >= S.iter_num) S.done = 1;
    return S.done ? 0 : 1;
  }
  if (hi < hi_full) {
    // the head of the first phase holds no trigger: go on with the same (identity) sampling set; the non-negative minimum is
    // folded once, when the whole phase has been evaluated (below: from iteration 0)
    S.iter = hi;
    return 1;
  }
  // no trigger: only a non-negative minNFA can still have improved (strictly, first occurrence).  In the first phase the
  // range starts at iteration 0 whatever parts it was evaluated in.
  const int a = R.argmin_first(S.reserve ? 0 : lo, hi);
  if (R.nfa(a) < S.min_nfa) { S.min_nfa = R.nfa(a); S.best_it = a; S.best_model = R.model(a); S.n_inl = R.n_inl(a); }
  S.iter = hi;
  if (S.reserve) {  // iter + 1 == iter_num && nIterReserve (:221)
    if (S.n_inl == 0) {
      S.extend_to = S.iter_num + S.reserve;
      return 1;  // same (identity) index set, but new iterations to evaluate
    }
    S.index_it = S.best_it; S.index_model = S.best_model; S.n_index = S.n_inl;
    S.iter_num = hi + S.reserve;
    S.reserve = 0;
    return 1;
  }
  S.done = 1;
  return 0;
}

// What ACRANSAC + GeometricFilter_FMatrix_AC::Fit leave (fundamental_acransac.h:44-47): the inliers are kept only for a
// meaningful model (minNFA < 0, :240-241) with at least 2.5 x MINIMUM_SAMPLES of them (homography_acransac.h:55-57 alike).
MVG_GEO_HD int ac_final_inliers(const AcState& S) {
  if (!(S.min_nfa < 0.0)) return 0;
  return (static_cast<double>(S.n_inl) < S.sample * 2.5) ? 0 : S.n_inl;
}

}  // namespace geo
}  // namespace mvgcuda

// ------------------------------------------------------------------------------------------ per-pair constants and scalar evaluation
// (host side of the product: log tables; host tests: a sequential evaluator.  The device evaluator in acransac_kernels.cuh
// is warp-parallel but calls the same core functions and must produce the same (nfa, n_inl, model) per iteration.)
namespace mvgcuda {
namespace geo {

struct PairGeo {
  int n;                 // putative matches of the pair == nData
  const double* x1;      // normalised left / right points, [n][2]
  const double* x2;
  double max_threshold;  // precision * N2(0,0)^2 (estimator_acransac.h:140-142)
  double logalpha0;      // log10(2 D / A / N2(0,0)) (estimator_acransac_kernel_adaptator.h:53-58)
  double loge0;          // log10(MAX_MODELS * (n - sample)) (estimator_acransac.h:153)
  const float* logc_n;   // [n + 1] log10 C(n, k)        (makelogcombi_n, :50-56)
  const float* logc_k;   // [n + 1] log10 C(k, sample)   (makelogcombi_k, :59-65)
  int sample;            // MINIMUM_SAMPLES
  double mult_error;     // 0.5 point-to-line, 1.0 point-to-point
};

struct Cand { double e; int i; };
// std::pair<double, size_t>::operator< (what std::sort(vec_residuals) uses, estimator_acransac.h:181)
MVG_GEO_HD bool cand_less(const Cand& a, const Cand& b) { return a.e < b.e || (!(b.e < a.e) && a.i < b.i); }

// bestNFA (estimator_acransac.h:73-96) over the sorted candidates (all residuals <= max_threshold): strict <, ascending k.
MVG_GEO_HD void best_nfa_scalar(const PairGeo& P, const Cand* list, int m, double& nfa, int& k_best) {
  nfa = ac_inf();
  k_best = P.sample;
  for (int k = P.sample + 1; k <= m; ++k) {
    const double v = nfa_term(P.logalpha0, P.loge0, P.mult_error, list[k - 1].e, k, P.sample, P.logc_n[k], P.logc_k[k]);
    if (v < nfa) { nfa = v; k_best = k; }
  }
}

// log10 C(n, k) tables exactly as logcombi accumulates them (estimator_acransac.h:39-48): r += log10(n - i + 1) - log10(i)
// for i = 1 .. min(k, n - k); the partial sums for growing k are the SAME additions, so one pass fills the whole table.
inline void make_logc_n(int n, float* out /*[n + 1]*/) {
  double r = 0.0;
  out[0] = 0.0f;
  for (int i = 1; 2 * i <= n; ++i) {  // k = i <= n - i: logcombi(i, n) is the i-th partial sum
    r += log10(static_cast<double>(n - i + 1)) - log10(static_cast<double>(i));
    out[i] = static_cast<float>(r);
  }
  for (int k = n / 2 + 1; k <= n; ++k) out[k] = out[n - k];  // n - k < k: logcombi uses k' = n - k; k == n gives 0
  if (n >= 1) out[n] = 0.0f;
}
inline void make_logc_k(int k, int nmax, float* out /*[nmax + 1]*/) {
  for (int n = 0; n <= nmax; ++n) {
    double r = 0.0;
    if (!(k >= n || k <= 0)) {
      const int kk = (n - k < k) ? n - k : k;
      for (int i = 1; i <= kk; ++i) r += log10(static_cast<double>(n - i + 1)) - log10(static_cast<double>(i));
    }
    out[n] = static_cast<float>(r);
  }
}

}  // namespace geo
}  // namespace mvgcuda
