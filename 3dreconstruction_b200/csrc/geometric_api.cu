// geometric_api.cu -- C ABI of the AC-RANSAC geometric filter (include/mvgcuda.h: mvgcuda_geometric_filter), the step of
// apps/compute_matches right after putative matching (compute_matches.cpp:250-318, geometric_filter.h:37-102).
// Compiled with -fmad=false: the double-precision arithmetic of the solver must not be contracted into FMAs
// (acransac_core.cuh).  Host side: per-pair constants that go through the C library's log10 exactly as the reference's do
// (log-combinatorial tables, logalpha0, loge0), the glibc rand() stream, batching; device side: acransac_kernels.cuh.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <new>
#include <algorithm>
#include <atomic>
#include <deque>
#include <vector>

#include "../../include/mvgcuda.h"
#include "acransac_kernels.cuh"
#include "host_util.cuh"
#include "pair_chain.h"

using namespace mvgcuda;
using namespace mvgcuda::geo;

namespace {

constexpr int kGeoBatchPairs = 1024;        // active pairs per launch
constexpr long long kGeoBatchMatches = 4ll << 20;
constexpr int kGeoMaxMatchesPerPair = 16384;
constexpr int kNarrowRangeIterations = 148 * 4;  // a range of up to four warps per SM: latency-bound, nothing to gain from packing three iterations into one warp

struct GeoState {
  DevBuf<GeoPairDev> d_pairs;
  DevBuf<int2> d_matches, d_out;
  DevBuf<double2> d_x1, d_x2;
  DevBuf<float> d_logc_pool, d_logc_k;
  DevBuf<uint32_t> d_stream;
  DevBuf<IterRes> d_res;
  DevBuf<double> d_models, d_basis, d_ge;
  DevBuf<int> d_vec_index, d_gi, d_out_idx, d_out_count, d_out_iters, d_exact;
  DevBuf<RoundInfo> d_round;
  DevBuf<AcState> d_state;
  PinnedBuf<DecideOut> h_decide;
  DevBuf<long long> d_offsets;
  PinnedBuf<GeoPairDev> h_pairs;
  PinnedBuf<int2> h_matches;
  PinnedBuf<uint32_t> h_stream;
  PinnedBuf<int> h_out_count, h_out_iters;
  PinnedBuf<long long> h_offsets;
  PinnedBuf<AcState> h_state;
  PinnedBuf<RoundInfo> h_round;
  // results of the last call
  PinnedBuf<int> r_counts;
  PinnedBuf<long long> r_offsets;
  PinnedBuf<int> r_matches;
  // log tables (host copies; the pool only grows within a call)
  std::vector<float> logc_k;
  int logc_k_sample = 0;  // the MINIMUM_SAMPLES the table was built for
  std::vector<float> logc_pool;
  std::map<int, int> logc_off;  // n -> offset in the pool
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_prep = nullptr;
  cudaStream_t slot_stream[kGeoSlots] = {};   // every pair in flight advances on its own stream
};

void geo_free(void* p) {
  GeoState* G = static_cast<GeoState*>(p);
  if (!G) return;
  G->d_pairs.release(); G->d_matches.release(); G->d_out.release(); G->d_x1.release(); G->d_x2.release();
  G->d_logc_pool.release(); G->d_logc_k.release(); G->d_stream.release(); G->d_res.release(); G->d_models.release(); G->d_ge.release();
  G->d_vec_index.release(); G->d_gi.release(); G->d_out_idx.release(); G->d_out_count.release(); G->d_out_iters.release();
  G->d_round.release(); G->d_state.release(); G->h_decide.release(); G->d_basis.release(); G->d_exact.release();
  G->d_offsets.release(); G->h_state.release(); G->h_round.release();
  G->h_pairs.release(); G->h_matches.release(); G->h_stream.release(); G->h_out_count.release(); G->h_out_iters.release();
  G->h_offsets.release(); G->r_counts.release(); G->r_offsets.release(); G->r_matches.release();
  if (G->ev0) cudaEventDestroy(G->ev0);
  if (G->ev1) cudaEventDestroy(G->ev1);
  if (G->ev_prep) cudaEventDestroy(G->ev_prep);
  for (int q = 0; q < kGeoSlots; ++q) {
    if (G->slot_stream[q]) cudaStreamDestroy(G->slot_stream[q]);
  }
  delete G;
}

#define GEO_CHECK(ctx, expr)                                                                              \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      char _b[512];                                                                                       \
      snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      ctx_set_error(ctx, _b);                                                                             \
      return _e == cudaErrorMemoryAllocation ? MVGCUDA_ERR_NOMEM : MVGCUDA_ERR_CUDA;                      \
    }                                                                                                     \
  } while (0)

int fail(mvgcuda_ctx* ctx, const char* msg) {
  ctx_set_error(ctx, msg);
  return MVGCUDA_ERR_INVALID;
}

}  // namespace

extern "C" int mvgcuda_geometric_filter(mvgcuda_ctx* ctx, char model, double precision, int iterations, unsigned seed, int64_t n_pairs,
                                        const int32_t* pairs, const int32_t* counts, const int64_t* offsets, const int32_t* matches,
                                        const int32_t* image_sizes, mvgcuda_pair_matches* out) try {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  if (n_pairs < 0 || (n_pairs > 0 && (!pairs || !counts || !offsets || !matches)) || !image_sizes) return fail(ctx, "geometric_filter: null argument");
  if (model != 'f' && model != 'h')
    return fail(ctx, "geometric_filter: only the fundamental-matrix ('f') and homography ('h') models are built (SURVEY.md 8(f)-1)");
  // GeometricFilter_FMatrix_AC: 7 points, <= 3 models, point-to-line residual; GeometricFilter_HMatrix_AC: 4 points, 1 model,
  // point-to-point residual (homography_acransac.h:35-46)
  const int sample = model == 'f' ? kSampleF : kSampleH;
  const int max_models = model == 'f' ? kMaxModelsF : 1;
  if (iterations < 10 || iterations > (1 << 20)) return fail(ctx, "geometric_filter: iterations out of range");
  const CtxView V = ctx_view(ctx);
  if (!V.feats) return fail(ctx, "geometric_filter: call mvgcuda_set_features (or stream the features) first");
  GEO_CHECK(ctx, cudaSetDevice(V.device));
  if (!*V.geo) {
    *V.geo = new GeoState();
    *V.geo_free = geo_free;
  }
  GeoState& G = *static_cast<GeoState*>(*V.geo);
  if (!G.ev0) { GEO_CHECK(ctx, cudaEventCreate(&G.ev0)); GEO_CHECK(ctx, cudaEventCreate(&G.ev1)); }
  if (!G.ev_prep) {
    GEO_CHECK(ctx, cudaEventCreateWithFlags(&G.ev_prep, cudaEventDisableTiming));
    for (int q = 0; q < kGeoSlots; ++q) {
      GEO_CHECK(ctx, cudaStreamCreateWithFlags(&G.slot_stream[q], cudaStreamNonBlocking));
    }
  }
  cudaStream_t st = V.stream;
  int rc = ctx_wait_uploads(ctx);
  if (rc) return rc;

  // ---- validate, find the active pairs (ACRANSAC returns at once, drawing nothing, unless nData > 7: estimator_acransac.h:137-138)
  std::vector<int64_t> active;
  int n_max = 0;
  for (int64_t p = 0; p < n_pairs; ++p) {
    const int i = pairs[2 * p], j = pairs[2 * p + 1];
    if (i < 0 || j < 0 || i >= V.n_images || j >= V.n_images || counts[p] < 0) return fail(ctx, "geometric_filter: bad pair / count");
    for (int k = 0; k < counts[p]; ++k) {
      const int32_t* m = matches + 2 * (offsets[p] + k);
      if (m[0] < 0 || m[0] >= V.rows[i] || m[1] < 0 || m[1] >= V.rows[j]) return fail(ctx, "geometric_filter: match index out of range");
    }
    if (counts[p] > sample) {
      // the normalisation divides by sqrt(width * height) (conditioning.cpp:46-56): a missing size would give NaNs
      if (image_sizes[2 * i] < 1 || image_sizes[2 * i + 1] < 1 || image_sizes[2 * j] < 1 || image_sizes[2 * j + 1] < 1)
        return fail(ctx, "geometric_filter: image size (width, height) missing for an image with matches");
      active.push_back(p);
      n_max = std::max(n_max, counts[p]);
    }
  }
  // every warp of a range evaluation owns a candidate list of n_cap entries in global scratch (used when more than 256
  // residuals lie under the threshold): bounded so that the scratch stays below ~1.5 GB
  if (n_max > kGeoMaxMatchesPerPair) return fail(ctx, "geometric_filter: a pair with more than 16,384 putative matches is not supported");
  GEO_CHECK(ctx, G.r_counts.reserve(std::max<int64_t>(n_pairs, 1)));
  GEO_CHECK(ctx, G.r_offsets.reserve(n_pairs + 1));
  GEO_CHECK(ctx, G.r_matches.reserve(2));
  for (int64_t p = 0; p < n_pairs; ++p) G.r_counts.p[p] = 0;

  // ---- log tables through the C library's log10, accumulated exactly as logcombi does (estimator_acransac.h:39-65)
  if ((int)G.logc_k.size() < n_max + 1 || G.logc_k_sample != sample) {
    G.logc_k.resize(n_max + 1);
    make_logc_k(sample, n_max, G.logc_k.data());
    G.logc_k_sample = sample;
  }
  G.logc_pool.clear();
  G.logc_off.clear();
  for (int64_t p : active) {
    const int n = counts[p];
    if (G.logc_off.count(n)) continue;
    const int off = (int)G.logc_pool.size();
    G.logc_off[n] = off;
    G.logc_pool.resize(off + n + 1);
    make_logc_n(n, G.logc_pool.data() + off);
  }
  if (!active.empty()) {
    GEO_CHECK(ctx, G.d_logc_k.reserve(G.logc_k.size()));
    GEO_CHECK(ctx, G.d_logc_pool.reserve(std::max<size_t>(G.logc_pool.size(), 1)));
    GEO_CHECK(ctx, cudaMemcpyAsync(G.d_logc_k.p, G.logc_k.data(), G.logc_k.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    GEO_CHECK(ctx, cudaMemcpyAsync(G.d_logc_pool.p, G.logc_pool.data(), G.logc_pool.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  }

  // ---- scratch: kGeoSlots pairs in flight, each with its per-iteration arrays and its region of warp-scratch lists
  int n_cap = 32;
  while (n_cap < n_max) n_cap <<= 1;
  const int it_stride = iterations + 8;
  // warps of one range evaluation (one iteration per warp in a narrow range, kEvalGroup in a wide one); every slot has its
  // own region of warp-scratch lists, so that the launches of different slots can run side by side
  const int slot_warps = ((std::max(std::min(iterations, kNarrowRangeIterations), iterations / kEvalGroup + 1) + kEvalWarps - 1) / kEvalWarps) * kEvalWarps;
  // pairs in flight: all the slots, fewer when their scratch lists would be huge (pairs with many thousands of matches)
  int n_slots = kGeoSlots;
  while (n_slots > 4 && (size_t)n_slots * slot_warps * n_cap * (sizeof(double) + sizeof(int)) > ((size_t)3 << 30)) n_slots >>= 1;
  const int scratch_warps = n_slots * slot_warps;
  if (!active.empty()) {
    GEO_CHECK(ctx, G.d_res.reserve((size_t)kGeoSlots * it_stride));
    GEO_CHECK(ctx, G.d_models.reserve((size_t)kGeoSlots * it_stride * 27));
    GEO_CHECK(ctx, G.d_basis.reserve((size_t)kGeoSlots * it_stride * kBasisDoubles));
    GEO_CHECK(ctx, G.d_exact.reserve((size_t)kGeoSlots * it_stride));
    GEO_CHECK(ctx, G.d_vec_index.reserve((size_t)kGeoSlots * n_max));
    GEO_CHECK(ctx, G.d_ge.reserve((size_t)(kGeoSlots + scratch_warps) * n_cap));
    GEO_CHECK(ctx, G.d_gi.reserve((size_t)(kGeoSlots + scratch_warps) * n_cap));
    GEO_CHECK(ctx, G.d_round.reserve(kGeoSlots));
    GEO_CHECK(ctx, G.d_state.reserve(kGeoSlots));
    GEO_CHECK(ctx, G.h_decide.reserve(kGeoSlots));
    GEO_CHECK(ctx, G.h_state.reserve(kGeoSlots));
    GEO_CHECK(ctx, G.h_round.reserve(kGeoSlots));
  }
  long long exact_requests = 0, rounds = 0, respeculated = 0;

  // ---- the process-wide rand() stream: srand(seed) here, consumed pair after pair (the reference never seeds: seed 1)
  GlibcRand gen;
  glibc_srand(gen, seed);
  long long gen_pos = 0;   // values generated so far
  long long offset = 0;    // values the pairs processed so far have consumed
  std::vector<uint32_t> window;  // stream values [win_base, gen_pos)
  long long win_base = 0;
  float gpu_ms = 0.f;
  int launches = 0;

  // (tests: MVGCUDA_GEO_BATCH_PAIRS cuts a small collection into several batches -- offsets and generated-but-unconsumed
  // rand() values carry over from one to the next)
  int batch_pairs = kGeoBatchPairs;
  if (const char* v = getenv("MVGCUDA_GEO_BATCH_PAIRS")) batch_pairs = std::max(1, std::min(kGeoBatchPairs, atoi(v)));
  std::vector<std::vector<int32_t> > kept(active.size());  // filtered matches of the active pairs
  size_t a0 = 0;
  while (a0 < active.size()) {
    size_t a1 = a0;
    long long tot = 0;
    while (a1 < active.size() && (a1 - a0) < (size_t)batch_pairs) {
      const long long c = counts[active[a1]];
      if (a1 > a0 && tot + c > kGeoBatchMatches) break;
      tot += c;
      ++a1;
    }
    const int nb = (int)(a1 - a0);
    GEO_CHECK(ctx, G.h_pairs.reserve(nb));
    GEO_CHECK(ctx, G.h_matches.reserve(tot));
    GEO_CHECK(ctx, G.d_pairs.reserve(nb));
    GEO_CHECK(ctx, G.d_matches.reserve(tot));
    GEO_CHECK(ctx, G.d_x1.reserve(tot));
    GEO_CHECK(ctx, G.d_x2.reserve(tot));
    GEO_CHECK(ctx, G.d_out_idx.reserve(tot));
    GEO_CHECK(ctx, G.d_out.reserve(tot));
    GEO_CHECK(ctx, G.d_out_count.reserve(nb));
    GEO_CHECK(ctx, G.d_out_iters.reserve(nb));
    GEO_CHECK(ctx, G.d_offsets.reserve(nb + 1));
    GEO_CHECK(ctx, G.h_out_count.reserve(nb));
    GEO_CHECK(ctx, G.h_out_iters.reserve(nb));
    GEO_CHECK(ctx, G.h_offsets.reserve(nb + 1));
    long long mo = 0;
    for (int k = 0; k < nb; ++k) {
      const int64_t p = active[a0 + k];
      const int i = pairs[2 * p], j = pairs[2 * p + 1];
      GeoPairDev D;
      D.n = counts[p];
      D.m_off = (int)mo;
      D.row0_i = V.row0[i];
      D.row0_j = V.row0[j];
      D.logc_off = G.logc_off[D.n];
      D.out_slot = (int)p;
      const int wi = image_sizes[2 * i], hi = image_sizes[2 * i + 1], wj = image_sizes[2 * j], hj = image_sizes[2 * j + 1];
      D.N1 = make_normalizer(wi, hi);
      D.N2 = make_normalizer(wj, hj);
      // estimator_acransac.h:140-142, estimator_acransac_kernel_adaptator.h:53-58, estimator_acransac.h:153
      D.max_threshold = std::isinf(precision) ? precision : precision * D.N2.d * D.N2.d;
      const double diag = sqrt(wj * (double)wj + hj * (double)hj);
      const double area = wj * (double)hj;
      // point to line: ratio of the image diagonal over its area; point to point: unit circle over the image area
      // (estimator_acransac_kernel_adaptator.h:53-63)
      D.logalpha0 = model == 'f' ? log10(2.0 * diag / area / D.N2.d) : log10(M_PI / (wj * (double)hj) / (D.N2.d * D.N2.d));
      D.loge0 = log10((double)max_models * (size_t)(D.n - sample));
      G.h_pairs.p[k] = D;
      for (int q = 0; q < D.n; ++q) {
        const int32_t* m = matches + 2 * (offsets[p] + q);
        G.h_matches.p[mo + q] = make_int2(m[0], m[1]);
      }
      mo += D.n;
    }
    // stream window [offset, offset + sample * iterations * nb): room for the whole batch, but the values are GENERATED as
    // the pairs are started (start_pair below: 28,672 values = ~90 us of host time per pair, beside the GPU's work instead
    // of in front of it) and each pair uploads its own range on its own stream.  Values generated for the previous batch
    // and not consumed by it are carried over.
    const long long need_end = offset + (long long)sample * iterations * nb;
    const size_t n_stream = (size_t)(need_end - offset);
    GEO_CHECK(ctx, G.h_stream.reserve(n_stream));
    GEO_CHECK(ctx, G.d_stream.reserve(n_stream));
    if (offset > win_base) {
      window.erase(window.begin(), window.begin() + (size_t)std::min<long long>(offset - win_base, (long long)window.size()));
      win_base = offset;
    }
    if (!window.empty()) memcpy(G.h_stream.p, window.data(), std::min(window.size(), n_stream) * sizeof(uint32_t));
    const long long batch_base = offset;
    GEO_CHECK(ctx, cudaMemcpyAsync(G.d_pairs.p, G.h_pairs.p, nb * sizeof(GeoPairDev), cudaMemcpyHostToDevice, st));
    GEO_CHECK(ctx, cudaMemcpyAsync(G.d_matches.p, G.h_matches.p, tot * sizeof(int2), cudaMemcpyHostToDevice, st));
    GEO_CHECK(ctx, cudaEventRecord(G.ev0, st));

    GeoBatchDev B;
    B.pairs = G.d_pairs.p; B.n_pairs = nb; B.matches = G.d_matches.p; B.feats = V.feats; B.x1 = G.d_x1.p; B.x2 = G.d_x2.p;
    B.logc_pool = G.d_logc_pool.p; B.logc_k = G.d_logc_k.p; B.stream = G.d_stream.p; B.stream_base = offset;
    B.max_iterations = iterations; B.model = model == 'f' ? 0 : 1; B.sample = sample; B.mult_error = model == 'f' ? 0.5 : 1.0; B.it_stride = it_stride; B.n_max = n_max; B.n_cap = n_cap;
    B.res = G.d_res.p; B.models = G.d_models.p; B.basis = G.d_basis.p; B.exact = G.d_exact.p; B.vec_index = G.d_vec_index.p;
    B.state = G.d_state.p; B.round = G.d_round.p; B.decide = G.h_decide.p /* page-locked host memory (unified addressing): the verdicts are polled, not copied */; B.g_e = G.d_ge.p; B.g_i = G.d_gi.p;
    B.out_idx = G.d_out_idx.p; B.out_count = G.d_out_count.p; B.out_iters = G.d_out_iters.p;
    geo_prep_kernel<<<nb, 256, 0, st>>>(B);
    GEO_CHECK(ctx, cudaGetLastError());
    GEO_CHECK(ctx, cudaEventRecord(G.ev_prep, st));
    ++launches;

    // The pipeline.  The chain orders only the STARTS of the pairs: pair p + 1 is admitted (into a free slot) as soon as
    // pair p's iteration count is final -- its rand() offset follows -- and from then on each advances on its own stream:
    // evaluation of its pending range (or the re-evaluation of one iteration with roots from THIS machine's C library),
    // its accounting warp, the copy of the verdict, an event the host polls.  No pair waits for another one's kernels.
    //
    // The order of the starts, the offsets they assume and what happens when an assumption fails: pair_chain.h.
    enum { kNeedEval = 1, kNeedExact = 2 };
    struct HostSlot { int need = kNeedEval; int lo = 0, hi = 0; int it = 0; double P[4]; bool in_flight = false, discard = false; int seq = 0; int idle_polls = 0; };
    HostSlot slots[kGeoSlots];
    for (int q = 0; q < kGeoSlots; ++q) G.h_decide.p[q].seq = 0;
    int launch_seq = 0;
    PairChain chain(n_slots, nb, sample, iterations, offset);
    constexpr int kFirstRange = 192;  // iterations of a pair's first range while pairs with geometry are about
    auto start_pair = [&](int sl, int pair, long long off) -> int {
      AcState S0;
      ac_init(S0, G.h_pairs.p[pair].n, iterations, sample);
      RoundInfo R0;
      R0.pair = pair; R0.lo = S0.iter; R0.hi = ac_range_end(S0); R0.n_index = S0.n_index; R0.offset = off;
      // pairs with geometry about: only the head of the first phase at first -- their trigger is almost always in it, and
      // everything evaluated after a trigger is thrown away (a no-model pair pays one more round for the rest)
      if (chain.geometry_about() && R0.hi > kFirstRange) R0.hi = kFirstRange;
      G.h_state.p[sl] = S0;
      G.h_round.p[sl] = R0;
      GEO_CHECK(ctx, cudaStreamWaitEvent(G.slot_stream[sl], G.ev_prep, 0));  // the batch's points are normalised
      {  // the pair's part of the rand() stream: generated up to its end if it is not yet, uploaded on the pair's stream
        const long long end = off + (long long)sample * iterations;
        while (gen_pos < end) { G.h_stream.p[gen_pos - batch_base] = glibc_rand(gen); ++gen_pos; }
        GEO_CHECK(ctx, cudaMemcpyAsync(G.d_stream.p + (off - batch_base), G.h_stream.p + (off - batch_base),
                                       (size_t)sample * iterations * sizeof(uint32_t), cudaMemcpyHostToDevice, G.slot_stream[sl]));
      }
      GEO_CHECK(ctx, cudaMemcpyAsync(G.d_state.p + sl, G.h_state.p + sl, sizeof(AcState), cudaMemcpyHostToDevice, G.slot_stream[sl]));
      GEO_CHECK(ctx, cudaMemcpyAsync(G.d_round.p + sl, G.h_round.p + sl, sizeof(RoundInfo), cudaMemcpyHostToDevice, G.slot_stream[sl]));
      HostSlot& H = slots[sl];
      // (a restart while the slot's previous launches are still running: they come first in its stream; their verdict is dropped)
      if (H.in_flight) H.discard = true;
      H.need = kNeedEval; H.lo = R0.lo; H.hi = R0.hi;
      return MVGCUDA_OK;
    };
    while (!chain.finished()) {
      // admissions
      {
        int sl, pair;
        long long off;
        while (chain.admit(&sl, &pair, &off)) {
          rc = start_pair(sl, pair, off);
          if (rc) return rc;
        }
      }
      // launches: every slot with something to do and nothing in flight gets, on ITS stream, the evaluation of its
      // pending range followed by its accounting warp -- or, in one launch, the re-evaluation of one iteration with roots
      // from this machine's C library and the accounting that was waiting for it.  The accounting warp stores its verdict
      // straight into page-locked host memory, sequence number last; the host polls that number: no copy, no event
      for (int q = 0; q < kGeoSlots; ++q) {
        HostSlot& H = slots[q];
        if (H.in_flight || chain.slot(q).state != PairChain::kActive) continue;
        cudaStream_t sq = G.slot_stream[q];
        H.seq = ++launch_seq;
        if (H.need == kNeedEval) {
          EvalList EL;
          EL.n = 1; EL.slot[0] = q; EL.first_warp[0] = 0;
          // a narrow range is bound by the latency of one warp (the SVD chain, then its evaluations one after the other):
          // one iteration per warp; a wide one by the fp64 pipe: three
          EL.group = (H.hi - H.lo) <= kNarrowRangeIterations ? 1 : kEvalGroup;
          EL.scratch_base = q * slot_warps;
          const int n_w = (H.hi - H.lo + EL.group - 1) / EL.group;
          EL.first_warp[1] = n_w;
          if (n_w > 0) {
            geo_eval_kernel<<<(n_w + kEvalWarps - 1) / kEvalWarps, 32 * kEvalWarps, 0, sq>>>(B, EL);
            GEO_CHECK(ctx, cudaGetLastError());
            ++launches;
          }
          DecideList DL;
          DL.n = 1; DL.slot[0] = q; DL.seq[0] = H.seq;
          geo_decide_kernel<<<1, 32, 0, sq>>>(B, DL);
          GEO_CHECK(ctx, cudaGetLastError());
          ++launches;
        } else {
          ExactList XL;
          XL.n = 1; XL.slot[0] = q; XL.it[0] = H.it; XL.seq[0] = H.seq;
          XL.nr[0] = solve_cubic(H.P, XL.roots[0]);  // host libm: acos / cos / pow as the reference's process would call them
          geo_exact_kernel<<<1, 32, 0, sq>>>(B, XL);
          GEO_CHECK(ctx, cudaGetLastError());
          ++launches;
          ++exact_requests;
        }
        ++rounds;
        H.in_flight = true;
      }
      // wait until at least one slot has its verdict, then take every verdict that is there
      int completed = 0;
      long long spins = 0;
      while (!completed) {
        bool any_in_flight = false;
        for (int q = 0; q < kGeoSlots; ++q) {
          HostSlot& H = slots[q];
          if (!H.in_flight) continue;
          any_in_flight = true;
          const volatile DecideOut* hv = G.h_decide.p + q;
          if (hv->seq != H.seq) continue;
          std::atomic_thread_fence(std::memory_order_acquire);
          H.in_flight = false;
          ++completed;
          if (H.discard) { H.discard = false; continue; }   // the verdict of a refuted start
          const DecideOut D = *const_cast<const DecideOut*>(G.h_decide.p + q);
          chain.note_counts(q, D.iters_final, D.status == 1 ? D.iters_guess : -1);
          if (D.status == 1) {
            H.need = kNeedExact; H.it = D.it;
            for (int c = 0; c < 4; ++c) H.P[c] = D.P[c];
          } else if (D.status == 2) {
            chain.pair_done(q);      // counted now, or held until its offset is confirmed
          } else {
            H.need = kNeedEval; H.lo = D.next.lo; H.hi = D.next.hi;
          }
        }
        if (!any_in_flight) return fail(ctx, "geometric_filter: internal error (no pair in flight)");
        if (!completed && (++spins & 0xffff) == 0) {
          // nothing for a while: a failed launch would never publish -- ask the streams
          for (int q = 0; q < kGeoSlots; ++q) {
            if (!slots[q].in_flight) continue;
            const cudaError_t qe = cudaStreamQuery(G.slot_stream[q]);
            if (qe != cudaSuccess && qe != cudaErrorNotReady) GEO_CHECK(ctx, qe);
            if (qe == cudaSuccess && static_cast<const volatile DecideOut*>(G.h_decide.p + q)->seq != slots[q].seq && (++slots[q].idle_polls > 2))
              return fail(ctx, "geometric_filter: internal error (a slot's launches ended without a verdict)");
          }
        }
      }
      // the chain: front pairs whose count is final leave it and confirm or refute the offset of the next one
      rc = chain.resolve(start_pair);
      if (rc) return rc;
    }
    respeculated += chain.refuted();
    for (int q = 0; q < kGeoSlots; ++q) GEO_CHECK(ctx, cudaStreamSynchronize(G.slot_stream[q]));  // (idle by now: the verdicts are in)
    GEO_CHECK(ctx, cudaMemcpyAsync(G.h_out_count.p, G.d_out_count.p, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
    GEO_CHECK(ctx, cudaMemcpyAsync(G.h_out_iters.p, G.d_out_iters.p, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
    GEO_CHECK(ctx, cudaStreamSynchronize(st));
    const long long offset_after = chain.next_offset();
    // values generated beyond what the batch consumed belong to the next batch (or stay unused after the last one)
    window.assign(G.h_stream.p + (offset_after - batch_base), G.h_stream.p + (gen_pos - batch_base));
    win_base = offset_after;
    long long kept_total = 0;
    for (int k = 0; k < nb; ++k) { G.h_offsets.p[k] = kept_total; kept_total += G.h_out_count.p[k]; }
    G.h_offsets.p[nb] = kept_total;
    if (kept_total > 0) {
      GEO_CHECK(ctx, cudaMemcpyAsync(G.d_offsets.p, G.h_offsets.p, (nb + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
      geo_gather_kernel<<<nb, 256, 0, st>>>(B, G.d_offsets.p, G.d_out.p);
      GEO_CHECK(ctx, cudaGetLastError());
      GEO_CHECK(ctx, G.h_matches.reserve(kept_total));
      GEO_CHECK(ctx, cudaMemcpyAsync(G.h_matches.p, G.d_out.p, kept_total * sizeof(int2), cudaMemcpyDeviceToHost, st));
      ++launches;
    }
    GEO_CHECK(ctx, cudaEventRecord(G.ev1, st));
    GEO_CHECK(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    GEO_CHECK(ctx, cudaEventElapsedTime(&ms, G.ev0, G.ev1));
    gpu_ms += ms;
    for (int k = 0; k < nb; ++k) {
      const int c = G.h_out_count.p[k];
      std::vector<int32_t>& dst = kept[a0 + k];
      dst.resize((size_t)2 * c);
      for (int q = 0; q < c; ++q) {
        const int2 m = G.h_matches.p[G.h_offsets.p[k] + q];
        dst[2 * q] = m.x; dst[2 * q + 1] = m.y;
      }
    }
    offset = offset_after;
    a0 = a1;
  }

  // ---- results in the caller's pair order
  long long total = 0;
  for (size_t a = 0; a < active.size(); ++a) G.r_counts.p[active[a]] = (int)(kept[a].size() / 2);
  for (int64_t p = 0; p < n_pairs; ++p) { G.r_offsets.p[p] = total; total += G.r_counts.p[p]; }
  G.r_offsets.p[n_pairs] = total;
  GEO_CHECK(ctx, G.r_matches.reserve((size_t)std::max<long long>(total, 1) * 2));
  for (size_t a = 0; a < active.size(); ++a)
    if (!kept[a].empty()) memcpy(G.r_matches.p + 2 * G.r_offsets.p[active[a]], kept[a].data(), kept[a].size() * sizeof(int32_t));
  if (const char* v = getenv("MVGCUDA_GEO_STATS"))
    if (*v == '1')
      fprintf(stderr, "mvgcuda_geometric_filter('%c'): %zu active pairs, %lld rounds, %d launches, %lld models re-evaluated with host roots, "
                      "%lld speculative starts refuted, %lld rand() values\n", model, active.size(), rounds, launches, exact_requests, respeculated, offset);
  if (out) {
    out->n_pairs = n_pairs;
    out->counts = G.r_counts.p;
    out->offsets = reinterpret_cast<const int64_t*>(G.r_offsets.p);
    out->matches = G.r_matches.p;
    out->gpu_ms = gpu_ms;
    out->knn_kernel_ms = 0.f;
    out->knn_kernel_launches = (int32_t)std::min<long long>(exact_requests, 0x7fffffff);  // models re-evaluated with host roots
    out->total_launches = launches;
    out->rescanned_queries = offset;  // rand() values the filter consumed (the reference's stream position afterwards)
  }
  return MVGCUDA_OK;
} catch (const std::bad_alloc&) {
  if (ctx) ctx_set_error(ctx, "out of host memory");
  return MVGCUDA_ERR_NOMEM;
} catch (...) {
  if (ctx) ctx_set_error(ctx, "unexpected exception");
  return MVGCUDA_ERR_INVALID;
}

// Instrumentation: the scalar solver core on the device (one thread per case), for bit comparisons in the tests.
extern "C" int mvgcuda_geo_selftest(mvgcuda_ctx* ctx, int n, const double* x1, const double* x2, const double* probe, double* F,
                                    int32_t* n_models, double* err, double* nfa) try {
  if (!ctx || n < 1 || !x1 || !x2 || !probe || !F || !n_models || !err || !nfa) return MVGCUDA_ERR_INVALID;
  const CtxView V = ctx_view(ctx);
  GEO_CHECK(ctx, cudaSetDevice(V.device));
  DevBuf<double> d_x1, d_x2, d_p, d_F, d_e, d_n;
  DevBuf<int> d_nm;
  auto cleanup = [&]() { d_x1.release(); d_x2.release(); d_p.release(); d_F.release(); d_e.release(); d_n.release(); d_nm.release(); };
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = d_x1.reserve((size_t)n * 14);
  if (e == cudaSuccess) e = d_x2.reserve((size_t)n * 14);
  if (e == cudaSuccess) e = d_p.reserve((size_t)n * 4);
  if (e == cudaSuccess) e = d_F.reserve((size_t)n * 27);
  if (e == cudaSuccess) e = d_e.reserve(n);
  if (e == cudaSuccess) e = d_n.reserve(n);
  if (e == cudaSuccess) e = d_nm.reserve(n);
  if (e == cudaSuccess) e = cudaMemcpy(d_x1.p, x1, (size_t)n * 14 * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_x2.p, x2, (size_t)n * 14 * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_p.p, probe, (size_t)n * 4 * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    geo_selftest_kernel<<<(n + 63) / 64, 64>>>(n, d_x1.p, d_x2.p, d_p.p, d_F.p, d_nm.p, d_e.p, d_n.p);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(F, d_F.p, (size_t)n * 27 * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(n_models, d_nm.p, (size_t)n * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(err, d_e.p, (size_t)n * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(nfa, d_n.p, (size_t)n * 8, cudaMemcpyDeviceToHost);
  cleanup();
  GEO_CHECK(ctx, e);
  return MVGCUDA_OK;
} catch (...) {
  return MVGCUDA_ERR_NOMEM;
}

extern "C" int mvgcuda_geo_selftest_h(mvgcuda_ctx* ctx, int n, const double* x1, const double* x2, const double* probe, double* H, double* err) try {
  if (!ctx || n < 1 || !x1 || !x2 || !probe || !H || !err) return MVGCUDA_ERR_INVALID;
  const CtxView V = ctx_view(ctx);
  GEO_CHECK(ctx, cudaSetDevice(V.device));
  DevBuf<double> d_x1, d_x2, d_p, d_H, d_e;
  auto cleanup = [&]() { d_x1.release(); d_x2.release(); d_p.release(); d_H.release(); d_e.release(); };
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = d_x1.reserve((size_t)n * 8);
  if (e == cudaSuccess) e = d_x2.reserve((size_t)n * 8);
  if (e == cudaSuccess) e = d_p.reserve((size_t)n * 4);
  if (e == cudaSuccess) e = d_H.reserve((size_t)n * 9);
  if (e == cudaSuccess) e = d_e.reserve(n);
  if (e == cudaSuccess) e = cudaMemcpy(d_x1.p, x1, (size_t)n * 8 * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_x2.p, x2, (size_t)n * 8 * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_p.p, probe, (size_t)n * 4 * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    geo_selftest_h_kernel<<<(n + kEvalGroup - 1) / kEvalGroup, 32>>>(n, d_x1.p, d_x2.p, d_p.p, d_H.p, d_e.p);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(H, d_H.p, (size_t)n * 9 * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(err, d_e.p, (size_t)n * 8, cudaMemcpyDeviceToHost);
  cleanup();
  GEO_CHECK(ctx, e);
  return MVGCUDA_OK;
} catch (...) {
  return MVGCUDA_ERR_NOMEM;
}

// Write n_pairs match lists in the reference's text format, pairs with no match omitted == PairedIndexedMatchToStream over a
// PairWiseMatches map that only holds non-empty entries (geometric_filter.h:85-98, indexed_match_utils.h:22-38).  `pairs`
// must be in lexicographic (i, j) order (std::map iteration order).
extern "C" int mvgcuda_write_matches(const char* path, int64_t n_pairs, const int32_t* pairs, const int32_t* counts, const int64_t* offsets,
                                     const int32_t* matches, int skip_empty) try {
  if (!path || n_pairs < 0 || (n_pairs > 0 && (!pairs || !counts || !offsets || !matches))) return MVGCUDA_ERR_INVALID;
  FILE* f = fopen(path, "wb");
  if (!f) return MVGCUDA_ERR_IO;
  std::vector<char> buf(1 << 22);
  size_t used = 0;
  auto put_int = [&](int v, char sep) {
    char tmp[12];
    int n = 0;
    unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) buf[used++] = '-';
    while (n) buf[used++] = tmp[--n];
    buf[used++] = sep;
  };
  bool ok = true;
  for (int64_t p = 0; p < n_pairs && ok; ++p) {
    if (skip_empty && counts[p] == 0) continue;
    if (used + 64 > buf.size()) { ok = fwrite(buf.data(), 1, used, f) == used; used = 0; }
    put_int(pairs[2 * p], ' ');
    put_int(pairs[2 * p + 1], '\n');
    put_int(counts[p], '\n');
    const int32_t* m = matches + 2 * offsets[p];
    for (int c = 0; c < counts[p] && ok; ++c) {
      if (used + 64 > buf.size()) { ok = fwrite(buf.data(), 1, used, f) == used; used = 0; }
      put_int(m[2 * c], ' ');
      put_int(m[2 * c + 1], '\n');
    }
  }
  ok = ok && fwrite(buf.data(), 1, used, f) == used;
  if (fclose(f) != 0 || !ok) return MVGCUDA_ERR_IO;
  return MVGCUDA_OK;
} catch (...) {
  return MVGCUDA_ERR_NOMEM;
}
