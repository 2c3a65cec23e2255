// acransac_kernels.cuh -- AC-RANSAC geometric filter on the GPU, fundamental-matrix and homography models (SURVEY.md
// 8(f)-1; replaces ImageCollectionGeometricFilter::Filter + GeometricFilter_FMatrix_AC / GeometricFilter_HMatrix_AC,
// geometric_filter.h:37-102, fundamental_acransac.h:13-57, homography_acransac.h:18-62, estimator_acransac.h:125-245).
// Compiled with -fmad=false (see acransac_core.cuh).
//
// The reference's results are defined by ONE global rand() stream consumed pair after pair (7 or 4 values per iteration;
// the iteration count of a pair depends on when its first meaningful model turns up), so pairs form a chain: a pair's
// stream offset is known only when the previous pair's iteration count is final.  Only the STARTS of the pairs are
// ordered by that chain (geometric_api.cu: up to kGeoSlots pairs in flight, each on its own stream, some started ahead
// speculatively); the parallelism inside a pair comes from evaluating whole RANGES of iterations at once
// (geo_eval_kernel: sample, solve -- 9x9 Jacobi SVD bit-faithful to Eigen, three iterations side by side per warp, + cubic
// or QR preconditioner --, then per model all residuals with the lanes over the points, ordered ballot compaction of the
// candidates <= max_threshold, bitonic sort in shared memory (global scratch for long lists; the reference's std::sort
// restated step by step when a residual is NaN), parallel NFA scan), after which ONE warp accounts for them in order
// (decide_slot, acransac_engine.cuh) and writes its verdict into page-locked host memory.
// Exactness: the evaluation uses CUDA's acos / cos / pow for the cubic, which differ from the host C library's in the last
// bit of ~10 % of the roots; a model an ACRANSAC decision hinges on (a trigger, the best model at the end of the first
// phase, and anything within a guard band of those thresholds) is re-evaluated from the same null vectors with the roots
// the HOST's libm gives for its cubic (geo_exact_kernel) before the decision is taken, so accepted models -- and with them
// the sampling sets, the inlier lists and their order -- carry the reference's bits.  The homography model has no
// transcendental function: every bit is the device's own.
#pragma once
#include <cuda_runtime.h>

#include "acransac_engine.cuh"
#include "stdsort_restated.cuh"

namespace mvgcuda {
namespace geo {

constexpr int kGeoThreads = 1024;            // one CTA per SM
constexpr int kGeoWarps = kGeoThreads / 32;
constexpr int kListCap = 256;                // candidates a warp sorts in shared memory (power of two)
constexpr int kMaxModelsF = 3;

struct GeoPairDev {  // one ACTIVE pair (more than 7 putative matches) of a batch
  int n;             // putative matches == nData
  int m_off;         // first match of the pair in the batch arrays
  int row0_i, row0_j;  // first arena row of image I / J (feature coordinates)
  int logc_off;      // the pair's log10 C(n, .) table in the pool
  int out_slot;      // index of the pair in the caller's pair list
  Normalizer N1, N2;
  double max_threshold, logalpha0, loge0;
};

struct IterRes { double nfa; int n_inl; int model; };
constexpr int kBasisDoubles = 22;  // per iteration: f1[9], f2[9] (null vectors), P[4] (cubic, ascending powers)
constexpr double kGuardAbs = 1e-3, kGuardRel = 1e-6;  // NFA band around a decision threshold that forces exact roots

struct RoundInfo {   // what a warp needs to evaluate the current range of a slot; written by the accounting warp
  int pair, lo, hi, n_index;
  long long offset;  // absolute rand() position of iteration 0 of the pair
};

struct DecideOut {   // the accounting warp's verdict, per slot: written straight into page-locked HOST memory, `seq` last
  int status;        // 0: evaluate `next`, 1: iteration `it` needs exact roots (P), 2: the pair is finished
  int it;
  int iters_final;   // >= 0 once the number of iterations the pair will run is known (the next pair's rand() offset follows)
  int seq;           // the launch's sequence number, stored after a system-wide fence: the host polls it
  int iters_guess;   // status 1 in the first phase: the count if iteration `it` is confirmed as the first trigger (else -1)
  int pad;
  double P[4];
  RoundInfo next;
};

// Several pairs are in flight at once, each in its own SLOT (state + per-iteration scratch): the chain only orders the
// STARTS of the pairs -- a pair's rand() offset is known as soon as its predecessor's iteration count is final, which is
// after the predecessor's first accepted model -- so the later rounds of a pair run beside the first rounds of the next.
constexpr int kGeoSlots = 16;

// Launch arguments.  The host sends ONE slot per launch (every pair in flight has its own stream), so the lists hold a
// single entry: kernel parameters stay a few dozen bytes.
constexpr int kListSlots = 1;
struct EvalList {    // range evaluation: slot[k] covers warps [first_warp[k], first_warp[k + 1])
  int n;
  int group;         // iterations a warp takes: kEvalGroup when the range is wide (throughput), 1 when it is narrow (latency)
  int scratch_base;  // first warp-scratch list of this launch (launches of different slots run side by side)
  int slot[kListSlots];
  int first_warp[kListSlots + 1];
};
struct ExactList {   // iterations to re-evaluate with the host's roots (the accounting follows in the same warp)
  int n;
  int slot[kListSlots], it[kListSlots], nr[kListSlots], seq[kListSlots];
  double roots[kListSlots][3];
};
struct DecideList {
  int n;
  int slot[kListSlots], seq[kListSlots];
};

struct GeoBatchDev {
  const GeoPairDev* pairs;
  int n_pairs;
  const int2* matches;      // putative (_i, _j) of the batch, dense, pair after pair
  const float2* feats;      // (x, y) of every arena row
  double2* x1;              // normalised points, [matches of the batch]
  double2* x2;
  const float* logc_pool;   // log10 C(n, k) tables
  const float* logc_k;      // log10 C(k, 7), k = 0 .. n_max
  const uint32_t* stream;   // stream[k] = rand() value number stream_base + k of the process
  long long stream_base;
  int max_iterations;
  int model;                // 0: fundamental matrix (7-point solver, point-to-line residual), 1: homography (4 points, point-to-point)
  int sample;               // MINIMUM_SAMPLES of the model: rand() values drawn per iteration
  double mult_error;        // exponent of the residual in the NFA: 0.5 / 1.0
  int it_stride;            // max_iterations + 8: per-slot stride of the per-iteration arrays
  int n_max;                // per-slot stride of vec_index
  int n_cap;                // power of two >= n_max: stride of the candidate-list scratch
  IterRes* res;             // [slots][it_stride]
  double* models;           // [slots][it_stride][27]
  double* basis;            // [slots][it_stride][kBasisDoubles]
  int* exact;               // [slots][it_stride]: models computed with the host's roots
  int* vec_index;           // [slots][n_max]
  AcState* state;           // [slots]
  RoundInfo* round;         // [slots]
  DecideOut* decide;        // [slots]
  double* g_e;              // long candidate lists: [slots (accounting / exact warps) + slots x warps of a range evaluation][n_cap]
  int* g_i;
  int* out_idx;             // [matches of the batch]: inlier positions (into the pair's putative list), in residual order
  int* out_count;           // [n_pairs]
  int* out_iters;           // [n_pairs]: iterations the reference would have run (7 rand() values each)
};

struct SlotView {
  IterRes* res; double* models; double* basis; int* exact; int* vec_index; AcState* state; RoundInfo* round; DecideOut* decide;
  double* ge; int* gi;  // the slot's own candidate scratch (accounting / exact warps)
};
__device__ __forceinline__ SlotView slot_view(const GeoBatchDev& B, int slot) {
  SlotView v;
  v.res = B.res + static_cast<size_t>(slot) * B.it_stride;
  v.models = B.models + static_cast<size_t>(slot) * B.it_stride * 27;
  v.basis = B.basis + static_cast<size_t>(slot) * B.it_stride * kBasisDoubles;
  v.exact = B.exact + static_cast<size_t>(slot) * B.it_stride;
  v.vec_index = B.vec_index + static_cast<size_t>(slot) * B.n_max;
  v.state = B.state + slot; v.round = B.round + slot; v.decide = B.decide + slot;
  v.ge = B.g_e + static_cast<size_t>(slot) * B.n_cap;
  v.gi = B.g_i + static_cast<size_t>(slot) * B.n_cap;
  return v;
}

// ------------------------------------------------------------------------------------------ prep
// Normalised coordinates of every putative match (NormalizePoints of both images, once per pair in the reference).
__global__ void __launch_bounds__(256)
geo_prep_kernel(GeoBatchDev B) {
  const GeoPairDev P = B.pairs[blockIdx.x];
  for (int k = threadIdx.x; k < P.n; k += blockDim.x) {
    const int2 m = B.matches[P.m_off + k];
    const float2 a = B.feats[P.row0_i + m.x], b = B.feats[P.row0_j + m.y];
    double2 o;
    normalize_point(P.N1, a.x, a.y, o.x, o.y);
    B.x1[P.m_off + k] = o;
    normalize_point(P.N2, b.x, b.y, o.x, o.y);
    B.x2[P.m_off + k] = o;
  }
}

// ------------------------------------------------------------------------------------------ warp-level pieces
struct WarpScratch {
  double W[81], V[81], F[27];
  double le[kListCap];
  int li[kListCap];
  int n_models, pad;
};

// Candidates of one model: residual of every point, those <= max_threshold kept in index order (ordered compaction), then
// sorted by (residual, index) == the head of std::sort(vec_residuals) (estimator_acransac.h:176-181).  Returns their number
// m; e / idx point to the sorted list (shared memory when m <= kListCap, else the warp's global scratch).
// one copy of the sequential sort per kernel image (it is inlined nowhere: the path is rare and long)
__device__ __noinline__ int nan_order_scan(double* ge, int* gi, int n, int sample, double max_threshold) {
  libstdcxx_sort(ge, gi, n);
  return nfa_scan_end(ge, n, sample, max_threshold);
}

__device__ __forceinline__ int model_candidates_warp(const GeoBatchDev& B, const GeoPairDev& P, const double* F, double* list_e, int* list_i,
                                                     double* ge, int* gi, int lane, double*& e_out, int*& i_out) {
  double f[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) f[k] = F[k];
  const double2* x1 = B.x1 + P.m_off;
  const double2* x2 = B.x2 + P.m_off;
  int m = 0;
  double* le = list_e;
  int* li = list_i;
  int cap = kListCap;
  unsigned any_nan = 0u;
  for (int pass = 0; pass < 2; ++pass) {
    m = 0;
    for (int base = 0; base < P.n; base += 32) {
      const int i = base + lane;
      bool keep = false;
      double e = 0.0;
      if (i < P.n) {
        const double2 a = x1[i], b = x2[i];
        e = B.model ? homography_error(f, a.x, a.y, b.x, b.y) : epipolar_error(f, a.x, a.y, b.x, b.y);
        keep = e <= P.max_threshold;
      }
      any_nan |= __ballot_sync(0xffffffffu, e != e);
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int slot = m + __popc(mask & ((1u << lane) - 1u));
        if (slot < cap) { le[slot] = e; li[slot] = i; }
      }
      m += __popc(mask);
    }
    if (any_nan || m <= cap) break;
    le = ge; li = gi; cap = B.n_cap;  // a long list: once more, into global scratch
  }
  if (any_nan) {
    // A residual is NaN (0 / 0 under a degenerate model): the reference's std::sort over ALL residuals has no strict weak
    // ordering to work with and its result is defined by its steps only -- repeat them (stdsort_restated.cuh), then take
    // what bestNFA's scan reaches.  Rare (degenerate samples over exactly collinear / coincident points): one lane.
    for (int i = lane; i < P.n; i += 32) {
      const double2 a = x1[i], b = x2[i];
      ge[i] = B.model ? homography_error(f, a.x, a.y, b.x, b.y) : epipolar_error(f, a.x, a.y, b.x, b.y);
      gi[i] = i;
    }
    __syncwarp();
    int m_scan = 0;
    if (lane == 0) m_scan = nan_order_scan(ge, gi, P.n, B.sample, P.max_threshold);
    m_scan = __shfl_sync(0xffffffffu, m_scan, 0);
    __syncwarp();
    e_out = ge; i_out = gi;
    return m_scan;
  }
  e_out = le; i_out = li;
  if (m < 2) { __syncwarp(); return m; }
  int p2 = 32;
  while (p2 < m) p2 <<= 1;
  for (int k = m + lane; k < p2; k += 32) { le[k] = ac_inf(); li[k] = 0x7fffffff; }
  __syncwarp();
  // bitonic network; every lane owns compare-exchanges (pair t = the t-th index with bit j clear and its partner), so no
  // lane idles on the upper element of a pair
  for (int k = 2; k <= p2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (p2 >> 1); t += 32) {
        const int idx = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = idx | j;
        const Cand a{le[idx], li[idx]}, b{le[ixj], li[ixj]};
        const bool up = (idx & k) == 0;
        if (up ? cand_less(b, a) : cand_less(a, b)) { le[idx] = b.e; li[idx] = b.i; le[ixj] = a.e; li[ixj] = a.i; }
      }
      __syncwarp();
    }
  }
  return m;
}

// bestNFA over the sorted candidates (estimator_acransac.h:73-96): minimum over k = sample + 1 .. m, lowest k among equal values.
__device__ __forceinline__ void best_nfa_warp(const GeoBatchDev& B, const GeoPairDev& P, const double* e, int m, int lane, double& nfa, int& k_best) {
  double v = ac_inf();
  int kb = 0x7fffffff;
  const float* lcn = B.logc_pool + P.logc_off;
  for (int k = B.sample + 1 + lane; k <= m; k += 32) {
    const double t = nfa_term(P.logalpha0, P.loge0, B.mult_error, e[k - 1], k, B.sample, lcn[k], B.logc_k[k]);
    if (t < v) { v = t; kb = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int ok = __shfl_xor_sync(0xffffffffu, kb, o);
    if (ov < v || (ov == v && ok < kb)) { v = ov; kb = ok; }
  }
  nfa = v;
  k_best = v < ac_inf() ? kb : B.sample;
}

// All models of one iteration (F: nm x 9 in shared memory): best NFA over the models, strict <, in order
// (estimator_acransac.h:173-217).  list_e / list_i: the warp's kListCap-entry candidate list in shared memory.
__device__ __forceinline__ IterRes evaluate_models_warp(const GeoBatchDev& B, const GeoPairDev& P, const double* F, double* list_e, int* list_i, int nm,
                                                        double* ge, int* gi, int lane) {
  double best = ac_inf();
  int best_k = 0, best_model = 0;
  for (int k = 0; k < nm; ++k) {
    double* e;
    int* idx;
    const int m = model_candidates_warp(B, P, F + 9 * k, list_e, list_i, ge, gi, lane, e, idx);
    if (m > B.sample) {
      double v;
      int kb;
      best_nfa_warp(B, P, e, m, lane, v, kb);
      if (v < best) { best = v; best_k = kb; best_model = k; }
    }
    __syncwarp();
  }
  IterRes r;
  r.nfa = best; r.n_inl = best < ac_inf() ? best_k : 0; r.model = best_model;
  return r;
}

// ------------------------------------------------------------------------------------------ evaluation of ranges
// A warp takes kEvalGroup consecutive iterations of [round.lo, round.hi) of a slot (estimator_acransac.h:166-218).
//   solve:     the Jacobi SVD is a chain of ~250 dependent 2x2 steps (six fp64 divisions and three square roots each) in
//              which nine lanes at most have element pairs to rotate -- the fp64 pipe, which a warp instruction occupies
//              whatever its active lanes, is what a GPU full of evaluations runs out of.  So THREE iterations are solved side
//              by side, lanes 10 g .. 10 g + 8 on the matrices of iteration g: one instruction stream, every lane computes
//              the 2x2 step of its own group, a step is skipped only when no group needs it (a converged matrix asks for
//              nothing more, exactly as its own loop would have ended).  Per element exactly the scalar operations.
//   evaluate:  the models of the three iterations one after the other, the whole warp on the residuals of each.
constexpr int kEvalWarps = 4;
constexpr int kEvalGroup = 3;
constexpr int kGroupLanes = 10;
struct SolveScratch {
  double W[81], V[81], F[27];
  int n_models, pad;
};
struct EvalScratch {
  SolveScratch s[kEvalGroup];
  union {
    struct { double le[kListCap]; int li[kListCap]; } list;
    double A[kEvalGroup][144];   // homography: the 16 x 9 action matrices while the QR preconditioner runs (before any list)
  } u;
  double colmax[kEvalGroup][9];
};

// jacobi_svd9_sweeps (acransac_core.cuh) on kEvalGroup work matrices at once; V already initialised.
__device__ __forceinline__ void jacobi_svd9_sweeps_groups(EvalScratch& es, int lane) {
  const double precision = 2.0 * DBL_EPSILON;
  const double consider_as_zero = 2.0 * 4.9406564584124654e-324;
  const int g = lane / kGroupLanes;              // 3: lanes 30, 31 have no group
  const int gl = lane - g * kGroupLanes;
  const bool member = g < kEvalGroup;
  const bool worker = member && gl < 9;
  double* W = es.s[member ? g : 0].W;
  double* V = es.s[member ? g : 0].V;
  if (worker) {
    double m = 0.0;
    for (int r = 0; r < 9; ++r) { const double a = fabs(W[r + 9 * gl]); if (a > m) m = a; }
    es.colmax[g][gl] = m;
  }
  __syncwarp();
  double scale = 0.0;
  if (member) for (int c = 0; c < 9; ++c) { const double t = es.colmax[g][c]; if (t > scale) scale = t; }
  if (scale == 0.0) scale = 1.0;
  const double inv = 1.0 / scale;
  if (worker) for (int r = 0; r < 9; ++r) W[r + 9 * gl] *= inv;
  __syncwarp();
  bool again = true;
  while (again) {
    bool rotated = false;
    for (int p = 1; p < 9; ++p) {
      for (int q = 0; q < p; ++q) {
        const double wpp = W[p + 9 * p], wqq = W[q + 9 * q], wpq = W[p + 9 * q], wqp = W[q + 9 * p];
        const double app = fabs(wpp), aqq = fabs(wqq);
        const double mx = app < aqq ? aqq : app;
        const double pm = precision * mx;
        const double threshold = consider_as_zero < pm ? pm : consider_as_zero;
        const double apq = fabs(wpq), aqp = fabs(wqp);
        const double off = apq < aqp ? aqp : apq;
        const bool need = member && off > threshold;   // uniform within a group
        if (!__any_sync(0xffffffffu, need)) continue;
        rotated = rotated || need;
        Rot jl, jr;
        real_2x2_jacobi_svd(wpp, wpq, wqp, wqq, jl, jr);
        __syncwarp();
        if (need && worker && !(jl.c == 1.0 && jl.s == 0.0)) rotate_pair(W[p + 9 * gl], W[q + 9 * gl], jl.c, jl.s);
        __syncwarp();
        if (need && worker && !(jr.c == 1.0 && -jr.s == 0.0)) {
          rotate_pair(W[gl + 9 * p], W[gl + 9 * q], jr.c, -jr.s);
          rotate_pair(V[gl + 9 * p], V[gl + 9 * q], jr.c, -jr.s);
        }
        __syncwarp();
      }
    }
    again = __any_sync(0xffffffffu, rotated);
  }
  // singular values = |diagonal|, selection sort in descending order (first maximum wins), stop at an exact zero
  double sv[9];
  for (int i = 0; i < 9; ++i) sv[i] = fabs(W[i + 9 * i]);
  for (int i = 0; i < 9; ++i) {
    int pos = 0;
    double best = sv[i];
    for (int k = 1; k < 9 - i; ++k) if (sv[i + k] > best) { best = sv[i + k]; pos = k; }
    if (best == 0.0) break;
    if (pos) {
      pos += i;
      const double t = sv[i]; sv[i] = sv[pos]; sv[pos] = t;
      if (worker) { const double u = V[gl + 9 * pos]; V[gl + 9 * pos] = V[gl + 9 * i]; V[gl + 9 * i] = u; }
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(32 * kEvalWarps)
geo_eval_kernel(GeoBatchDev B, EvalList L) {
  __shared__ EvalScratch scratch[kEvalWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * kEvalWarps + warp;
  int k = 0;
  while (k + 1 < L.n && gwarp >= L.first_warp[k + 1]) ++k;
  if (gwarp >= L.first_warp[L.n]) return;
  const SlotView V = slot_view(B, L.slot[k]);
  const RoundInfo R = *V.round;
  const int it0 = R.lo + L.group * (gwarp - L.first_warp[k]);
  if (it0 >= R.hi) return;
  const GeoPairDev P = B.pairs[R.pair];
  const bool identity = V.state->index_it < 0;
  EvalScratch& es = scratch[warp];
  double* ge = B.g_e + static_cast<size_t>(kGeoSlots + L.scratch_base + gwarp) * B.n_cap;
  int* gi = B.g_i + static_cast<size_t>(kGeoSlots + L.scratch_base + gwarp) * B.n_cap;
  const int g = lane / kGroupLanes, gl = lane - g * kGroupLanes;
  const bool member = g < kEvalGroup;
  const int it = it0 + g;
  const bool live = g < L.group && it < R.hi;     // the group has an iteration to solve
  const bool leader = member && gl == 0;
  SolveScratch& sg = es.s[member ? g : 0];
  const uint32_t* rs = B.stream + (R.offset + static_cast<long long>(B.sample) * it - B.stream_base);
  if (B.model == 0) {
    if (member && gl < 9)
      for (int r = 0; r < 9; ++r) sg.V[r + 9 * gl] = r == gl ? 1.0 : 0.0;
    if (leader) {
      for (int i = 0; i < 81; ++i) sg.W[i] = 0.0;   // (an idle group keeps a zero matrix: it never asks for a rotation)
      if (live) {
        uint32_t r[kSampleF];
        for (int q = 0; q < kSampleF; ++q) r[q] = rs[q];
        int s[kSampleF];
        random_sample<kSampleF>(r, R.n_index, s);
        for (int q = 0; q < kSampleF; ++q) {  // EncodeEpipolarEquation (seven_point_basis)
          const int id = identity ? s[q] : V.vec_index[s[q]];
          const double2 a = B.x1[P.m_off + id], b = B.x2[P.m_off + id];
          sg.W[q + 9 * 0] = b.x * a.x; sg.W[q + 9 * 1] = b.x * a.y; sg.W[q + 9 * 2] = b.x;
          sg.W[q + 9 * 3] = b.y * a.x; sg.W[q + 9 * 4] = b.y * a.y; sg.W[q + 9 * 5] = b.y;
          sg.W[q + 9 * 6] = a.x;       sg.W[q + 9 * 7] = a.y;       sg.W[q + 9 * 8] = 1.0;
        }
      }
    }
    __syncwarp();
    jacobi_svd9_sweeps_groups(es, lane);
    if (leader && live) {
      double Pc[4], roots[3];
      cubic_from_null_vectors(sg.V + 9 * 8, sg.V + 9 * 7, Pc);
      const int nr = solve_cubic(Pc, roots);  // CUDA's acos / cos / pow: approximate in the last bit
      models_from_roots(sg.V + 9 * 8, sg.V + 9 * 7, roots, nr, sg.F);
      sg.n_models = nr;
      double* bs = V.basis + static_cast<size_t>(it) * kBasisDoubles;
      for (int q = 0; q < 9; ++q) { bs[q] = sg.V[9 * 8 + q]; bs[9 + q] = sg.V[9 * 7 + q]; }
      for (int q = 0; q < 4; ++q) bs[18 + q] = Pc[q];
      double* out = V.models + static_cast<size_t>(it) * 27;
      for (int q = 0; q < 9 * nr; ++q) out[q] = sg.F[q];
    }
  } else {
    // homography: no transcendental anywhere (QR preconditioner + Jacobi: +, -, *, /, sqrt), so the device model IS the
    // reference's model bit for bit and nothing ever has to be re-evaluated with host values
    if (leader) {
      if (live) {
        uint32_t r[kSampleH];
        for (int q = 0; q < kSampleH; ++q) r[q] = rs[q];
        int s[kSampleH];
        random_sample<kSampleH>(r, R.n_index, s);
        double a[2 * kSampleH], b[2 * kSampleH];
        for (int q = 0; q < kSampleH; ++q) {
          const int id = identity ? s[q] : V.vec_index[s[q]];
          const double2 u = B.x1[P.m_off + id], w = B.x2[P.m_off + id];
          a[2 * q] = u.x; a[2 * q + 1] = u.y; b[2 * q] = w.x; b[2 * q + 1] = w.y;
        }
        four_point_qr(a, b, es.u.A[g], sg.W, sg.V);
      } else {
        for (int i = 0; i < 81; ++i) { sg.W[i] = 0.0; sg.V[i] = 0.0; }
      }
    }
    __syncwarp();
    jacobi_svd9_sweeps_groups(es, lane);
    if (leader && live) {
      for (int q = 0; q < 9; ++q) sg.F[q] = sg.V[q + 9 * 8];
      sg.n_models = 1;
      double* out = V.models + static_cast<size_t>(it) * 27;
      for (int q = 0; q < 9; ++q) out[q] = sg.F[q];
    }
  }
  __syncwarp();
  for (int q = 0; q < L.group && it0 + q < R.hi; ++q) {
    const IterRes r = evaluate_models_warp(B, P, es.s[q].F, es.u.list.le, es.u.list.li, es.s[q].n_models, ge, gi, lane);
    if (lane == 0) { V.res[it0 + q] = r; V.exact[it0 + q] = B.model; }  // a homography model needs no second look
    __syncwarp();
  }
}

struct WarpScratch;
__device__ __forceinline__ void decide_slot(const GeoBatchDev& B, int slot, int seq, WarpScratch& ws, int lane);

// The same iterations again with the roots the host's C library computed for their cubics: bit-identical to the reference.
__global__ void __launch_bounds__(32)
geo_exact_kernel(GeoBatchDev B, ExactList L) {
  __shared__ WarpScratch ws;
  const int lane = threadIdx.x;
  const int k = blockIdx.x;
  const SlotView V = slot_view(B, L.slot[k]);
  const GeoPairDev P = B.pairs[V.round->pair];
  const int it = L.it[k], nr = L.nr[k];
  if (lane == 0) {
    const double* bs = V.basis + static_cast<size_t>(it) * kBasisDoubles;
    models_from_roots(bs, bs + 9, L.roots[k], nr, ws.F);
    ws.n_models = nr;
    double* out = V.models + static_cast<size_t>(it) * 27;
    for (int q = 0; q < 9 * nr; ++q) out[q] = ws.F[q];
  }
  __syncwarp();
  const IterRes r = evaluate_models_warp(B, P, ws.F, ws.le, ws.li, nr, V.ge, V.gi, lane);
  if (lane == 0) { V.res[it] = r; V.exact[it] = 1; }
  __syncwarp();
  decide_slot(B, L.slot[k], L.seq[k], ws, lane);   // the decision that was waiting for this model, in the same launch
}

// The accounting warp's view of the evaluated results (warp-parallel searches; every lane returns the same value).
struct WarpRange {
  const IterRes* res;
  int lane;
  __device__ int first_below(int lo, int hi, double thr) const {
    int f = 0x7fffffff;
    for (int i = lo + lane; i < hi; i += 32)
      if (res[i].nfa < thr) { f = i; break; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) f = min(f, __shfl_xor_sync(0xffffffffu, f, o));
    return f == 0x7fffffff ? -1 : f;
  }
  __device__ int argmin_first(int lo, int hi) const {
    double v = ac_inf();
    int a = 0x7fffffff;
    for (int i = lo + lane; i < hi; i += 32) {
      const double t = res[i].nfa;
      if (a == 0x7fffffff || t < v) { v = t; a = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oa = __shfl_xor_sync(0xffffffffu, a, o);
      if (oa != 0x7fffffff && (a == 0x7fffffff || ov < v || (!(v < ov) && oa < a))) { v = ov; a = oa; }
    }
    return a;
  }
  __device__ double nfa(int i) const { return res[i].nfa; }
  __device__ int n_inl(int i) const { return res[i].n_inl; }
  __device__ int model(int i) const { return res[i].model; }
};

// Lowest iteration of [lo, hi) that may lie below `thr`: exact entries are compared with thr, approximate ones with
// thr + guard.  -1 if none.
__device__ __forceinline__ int first_candidate(const SlotView& V, int lo, int hi, double thr, double guard, int lane) {
  int f = 0x7fffffff;
  for (int i = lo + lane; i < hi; i += 32) {
    const double v = V.res[i].nfa;
    if (V.exact[i] ? v < thr : v < thr + guard) { f = i; break; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) f = min(f, __shfl_xor_sync(0xffffffffu, f, o));
  return f == 0x7fffffff ? -1 : f;
}

// ------------------------------------------------------------------------------------------ accounting (one warp per slot)
// The range of the slot's `round` has been evaluated: ask for exact roots where a decision needs them, else account for
// the range (ac_account), materialise a new sampling set or the final inliers, and tell the host what comes next.
__device__ __forceinline__ void publish_verdict(DecideOut* dst, const DecideOut& D, int seq) {
  dst->status = D.status; dst->it = D.it; dst->iters_final = D.iters_final; dst->iters_guess = D.iters_guess;
  dst->P[0] = D.P[0]; dst->P[1] = D.P[1]; dst->P[2] = D.P[2]; dst->P[3] = D.P[3];
  dst->next = D.next;
  __threadfence_system();
  *reinterpret_cast<volatile int*>(&dst->seq) = seq;
}

__device__ __forceinline__ void decide_slot(const GeoBatchDev& B, int slot, int seq, WarpScratch& ws, int lane) {
  const SlotView V = slot_view(B, slot);
  DecideOut D;
  D.status = 0; D.it = 0; D.iters_final = -1; D.seq = 0; D.iters_guess = -1; D.pad = 0; D.P[0] = D.P[1] = D.P[2] = D.P[3] = 0.0;
  const RoundInfo R = *V.round;
  const GeoPairDev P = B.pairs[R.pair];
  AcState S = *V.state;
  // the evaluated range: the whole pending one, or only the head of the first phase (the host starts a pair with a short
  // first range when pairs with geometry are about: their trigger sits in the first few dozen iterations)
  const int lo = S.iter, hi_full = ac_range_end(S);
  const int hi = R.hi < hi_full ? R.hi : hi_full;
  const bool head_only = hi < hi_full;
  if (!S.done && lo < hi) {
    const bool ext = S.extend_to > 0;
    const double thr = ext ? ac_inf() : (S.min_nfa < 0.0 ? S.min_nfa : 0.0);
    const double guard = ext ? 0.0 : kGuardAbs + kGuardRel * fabs(thr);
    int need = -1;
    bool no_trigger_in_phase_one = false;
    const int t = first_candidate(V, lo, hi, thr, guard, lane);
    if (t >= 0) {
      if (!V.exact[t]) need = t;
    } else if (!ext && S.reserve > 0 && !head_only) {
      // no iteration of the whole first phase can be a trigger (none within the guard band of 0), so the loop will run
      // iter_num + reserve iterations whichever model the fold below settles on: the next pair may start already
      no_trigger_in_phase_one = true;
      // end of the first phase without a trigger: the best model so far becomes the sampling set -- the minimum must be
      // exact (over the whole phase, from iteration 0, whatever parts it was evaluated in)
      const WarpRange WR{V.res, lane};
      const int a = WR.argmin_first(0, hi);
      const double m = V.res[a].nfa;
      if (m < ac_inf()) {
        const double g2 = kGuardAbs + kGuardRel * fabs(m);
        int f = 0x7fffffff;
        for (int i = lane; i < hi; i += 32)
          if (!V.exact[i] && V.res[i].nfa <= m + g2) { f = i; break; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) f = min(f, __shfl_xor_sync(0xffffffffu, f, o));
        if (f != 0x7fffffff) need = f;
      }
    }
    if (need >= 0) {
      D.status = 1; D.it = need;
      const double* bs = V.basis + static_cast<size_t>(need) * kBasisDoubles;
      for (int q = 0; q < 4; ++q) D.P[q] = bs[18 + q];
      D.next = R;
      if (S.reserve == 0) D.iters_final = S.iter_num;
      else if (no_trigger_in_phase_one) D.iters_final = S.iter_num + S.reserve;
      else if (!ext && need == t) D.iters_guess = t + 1 + S.reserve;   // confirmed, t is the first trigger (:228-234)
      if (lane == 0) publish_verdict(V.decide, D, seq);
      return;
    }
  }
  const WarpRange WR{V.res, lane};
  const int old_it = S.index_it, old_model = S.index_model;
  ac_account(S, WR, head_only ? hi : -1);
  if (!S.done && (S.index_it != old_it || S.index_model != old_model)) {
    double* e;
    int* idx;
    model_candidates_warp(B, P, V.models + static_cast<size_t>(S.index_it) * 27 + 9 * S.index_model, ws.le, ws.li, V.ge, V.gi, lane, e, idx);
    for (int q = lane; q < S.n_index; q += 32) V.vec_index[q] = idx[q];
  }
  if (S.done) {
    const int n_final = ac_final_inliers(S);
    if (n_final > 0) {
      double* e;
      int* idx;
      model_candidates_warp(B, P, V.models + static_cast<size_t>(S.best_it) * 27 + 9 * S.best_model, ws.le, ws.li, V.ge, V.gi, lane, e, idx);
      for (int q = lane; q < n_final; q += 32) B.out_idx[P.m_off + q] = idx[q];
    }
    if (lane == 0) { B.out_count[R.pair] = n_final; B.out_iters[R.pair] = S.iter_num; }
  }
  __syncwarp();
  RoundInfo N = R;
  N.lo = S.iter; N.hi = ac_range_end(S); N.n_index = S.n_index;
  D.status = S.done ? 2 : 0;
  D.next = N;
  // the loop bound is final once the reserve has been spent (or, with nothing found so far, when the extension is fixed)
  if (S.done || S.reserve == 0) D.iters_final = S.iter_num;
  else if (S.extend_to > 0) D.iters_final = S.extend_to;
  if (lane == 0) { *V.state = S; *V.round = N; publish_verdict(V.decide, D, seq); }
}

__global__ void __launch_bounds__(32)
geo_decide_kernel(GeoBatchDev B, DecideList L) {
  __shared__ WarpScratch ws;
  decide_slot(B, L.slot[blockIdx.x], L.seq[blockIdx.x], ws, threadIdx.x);
}

// Filtered matches of the batch: pair p keeps putative[out_idx[k]] for k < out_count[p], in that (residual) order
// (geometric_filter.h:87-92).  offsets: exclusive scan of out_count (computed by the host).
__global__ void __launch_bounds__(256)
geo_gather_kernel(GeoBatchDev B, const long long* __restrict__ offsets, int2* __restrict__ out) {
  const GeoPairDev P = B.pairs[blockIdx.x];
  const int n = B.out_count[blockIdx.x];
  const long long o = offsets[blockIdx.x];
  for (int k = threadIdx.x; k < n; k += blockDim.x) out[o + k] = B.matches[P.m_off + B.out_idx[P.m_off + k]];
}

// ------------------------------------------------------------------------------------------ self-test
// The scalar core on the device, one thread per case, so that tests can compare its bits with the reference's
// (tests/test_gpu_geometric.py): 7-point models of a sample, the residual of a probe point under the first model, and one
// NFA term.
__global__ void geo_selftest_kernel(int n, const double* __restrict__ x1 /*[n][14]*/, const double* __restrict__ x2, const double* __restrict__ probe /*[n][4]*/,
                                    double* __restrict__ F /*[n][27]*/, int* __restrict__ n_models, double* __restrict__ err, double* __restrict__ nfa) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double W[81], V[81], a[14], b[14], f[27];
  for (int k = 0; k < 14; ++k) { a[k] = x1[14 * t + k]; b[k] = x2[14 * t + k]; }
  for (int k = 0; k < 27; ++k) f[k] = 0.0;
  const int nm = seven_point_models(a, b, W, V, f);
  n_models[t] = nm;
  for (int k = 0; k < 27; ++k) F[27 * t + k] = f[k];
  const double e = epipolar_error(f, probe[4 * t], probe[4 * t + 1], probe[4 * t + 2], probe[4 * t + 3]);
  err[t] = e;
  nfa[t] = nfa_term(-1.25, 2.5, 0.5, e, 8 + (t & 63), kSampleF, 3.5f, 1.25f);
}

// The homography path exactly as geo_eval_kernel runs it (three cases per warp: the group leaders build the action matrix
// and run the QR preconditioner, the groups sweep side by side): H of a 4-point sample and the transfer error of a probe
// point under it.
__global__ void __launch_bounds__(32)
geo_selftest_h_kernel(int n, const double* __restrict__ x1 /*[n][8]*/, const double* __restrict__ x2, const double* __restrict__ probe /*[n][4]*/,
                      double* __restrict__ H /*[n][9]*/, double* __restrict__ err) {
  __shared__ EvalScratch es;
  const int lane = threadIdx.x;
  const int g = lane / kGroupLanes, gl = lane - g * kGroupLanes;
  const bool member = g < kEvalGroup;
  const int t = blockIdx.x * kEvalGroup + g;
  const bool live = member && t < n, leader = member && gl == 0;
  SolveScratch& sg = es.s[member ? g : 0];
  if (leader) {
    if (live) {
      double a[2 * kSampleH], b[2 * kSampleH];
      for (int k = 0; k < 2 * kSampleH; ++k) { a[k] = x1[8 * t + k]; b[k] = x2[8 * t + k]; }
      four_point_qr(a, b, es.u.A[g], sg.W, sg.V);
    } else {
      for (int i = 0; i < 81; ++i) { sg.W[i] = 0.0; sg.V[i] = 0.0; }
    }
  }
  __syncwarp();
  jacobi_svd9_sweeps_groups(es, lane);
  if (leader && live) {
    for (int q = 0; q < 9; ++q) { sg.F[q] = sg.V[q + 9 * 8]; H[9 * t + q] = sg.F[q]; }
    err[t] = homography_error(sg.F, probe[4 * t], probe[4 * t + 1], probe[4 * t + 2], probe[4 * t + 3]);
  }
}

}  // namespace geo
}  // namespace mvgcuda
