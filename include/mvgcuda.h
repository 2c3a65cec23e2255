/*
 * mvgcuda.h -- C ABI of libmvgcuda: B200 (sm_100a) exhaustive putative matching of
 * uint8 SIFT-128 descriptors (brute-force squared-L2 2-NN + Lowe ratio test).
 *
 * This is the drop-in boundary for ONE hot path of yueying/3DReconstruction.  The reference
 * has no FFI of its own (it is header-only C++ templates); every entry point below names the
 * reference interface it stands in for.  Citations are relative to the reference tree.
 *
 *   array level      ArrayMatcher<uchar,Metric>::Build / SearchNeighbours
 *                    libs/feature/include/mvg/feature/matching_interface.h:16-64
 *                    libs/feature/include/mvg/feature/matcher_brute_force.h:42-50,102-134
 *   collection level MatcherAllInMemory::LoadData / Match
 *                    libs/feature/include/mvg/feature/matcher_all_in_memory.h:44-60,62-141
 *   filters          DistanceRatioFilter  matching_filters.h:27-47
 *                    IndexedMatch::getDeduplicated  indexed_match.h:49-55
 *                    IndexedMatchDecorator::getDeduplicated  indexed_match_decorator.h:90-104
 *   export           PairedIndexedMatchToStream  indexed_match_utils.h:22-38
 *
 * Conventions: plain C symbols, every function returns an int status (0 = MVGCUDA_OK), no
 * exception crosses the boundary (every entry point catches: std::bad_alloc -> MVGCUDA_ERR_NOMEM), all buffers are caller-owned unless stated, a context is
 * bound to ONE GPU and may be used from one host thread at a time (use one context per GPU).
 * Descriptors are dense row-major [rows][128] uint8 (== std::vector<Descriptor<uchar,128>>,
 * descriptor.h:23-58).  There is no CPU fallback: without a CUDA device every compute entry
 * point fails with MVGCUDA_ERR_CUDA.
 */
#ifndef MVGCUDA_H_
#define MVGCUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVGCUDA_DIM 128 /* descriptor length, Descriptor<unsigned char,128> (compute_matches.cpp:182-186) */

enum mvgcuda_status {
  MVGCUDA_OK = 0,
  MVGCUDA_ERR_INVALID = 1, /* bad argument (null pointer, negative size, unknown image id ...) */
  MVGCUDA_ERR_CUDA = 2,    /* CUDA runtime/driver failure, or no sm_100 device */
  MVGCUDA_ERR_NOMEM = 3,   /* host or device allocation failed */
  MVGCUDA_ERR_IO = 4       /* file could not be written */
};

/* How raw 2-NN indices are chosen when distances tie (matcher_brute_force.h:124-131 +
 * indexed_sort.h:52-66).  Exported MATCHES never depend on this (a tie d1==d2 never passes
 * the ratio test for ratio <= 1, matching_filters.h:44). */
enum mvgcuda_tie_mode {
  MVGCUDA_TIE_LOWEST_INDEX = 0, /* smallest db row wins (BASELINE.json north_star wording) */
  MVGCUDA_TIE_REFERENCE = 1     /* reproduce libstdc++ std::partial_sort(…,2) as used by
                                   SortIndexHelper: bit-identical raw indices */
};

typedef struct mvgcuda_ctx mvgcuda_ctx;

/* Library/ABI version (major*10000 + minor*100 + patch). */
int mvgcuda_version(void);

/* Number of visible CUDA devices with compute capability 10.x (0 if none / no driver). */
int mvgcuda_device_count(void);
/* CUDA ordinal of the k-th such device (k = 0 .. count-1), -1 if there is none: what to pass to mvgcuda_create when
 * the box mixes GPU generations or the sm_100 devices are not ordinals 0..N-1. */
int mvgcuda_device_ordinal(int k);

/* Create a context on CUDA device `device`.  Fails (MVGCUDA_ERR_CUDA) unless it is sm_100. */
int mvgcuda_create(int device, mvgcuda_ctx** out);
void mvgcuda_destroy(mvgcuda_ctx* ctx);

/* Last error text of this context (never NULL; "" when no error).  With ctx == NULL returns the
 * text of the last mvgcuda_create failure on this thread. */
const char* mvgcuda_last_error(const mvgcuda_ctx* ctx);

/* Run the context's work on an externally owned CUDA stream (cudaStream_t as void*), e.g.
 * torch.cuda.current_stream().cuda_stream, so that the caller's CUDA events bracket the kernels.
 * NULL restores the context's own stream. */
int mvgcuda_set_stream(mvgcuda_ctx* ctx, void* cuda_stream);

/* Tuning of the ratio-aware pruning of the pair / collection level (results never depend on it).
 * prune_rho in (0, 1]: a query that currently fails the ratio test only admits db rows with d <= prune_rho * d(best);
 * 1 = plain best-distance bound (no query is ever matched twice), default 0.72.  It is clamped from below to the ratio
 * of the call.  rescan_rows: rows of the buffer the ambiguous queries of a batch are gathered into (0 = default 2^20);
 * a batch with more of them is matched again through a bounded multi-round path (same results, one synchronisation). */
int mvgcuda_set_tuning(mvgcuda_ctx* ctx, float prune_rho, int rescan_rows);

/* ------------------------------------------------------------------------------------------
 * Residency: copy n_images descriptor arrays into the context's HBM arena (replaces the
 * map_descriptors residency of MatcherAllInMemory::LoadData, matcher_all_in_memory.h:44-60,
 * and ArrayMatcherBruteForce::Build, matcher_brute_force.h:42-50 -- but COPIES, so the caller
 * may free its buffers).  Replaces any previously uploaded set.  rows[i] >= 0; desc[i] may be
 * NULL when rows[i] == 0.  Also runs the per-row squared-norm kernel.
 * desc[i] may point to host memory (pageable or page-locked) or to DEVICE memory of any GPU of the box (unified
 * addressing; used to fan a replica out over NVLink instead of re-reading it over PCIe on every GPU).
 * `pinned` != 0 promises that host buffers are page-locked (async H2D). */
int mvgcuda_upload_images(mvgcuda_ctx* ctx, int n_images, const uint8_t* const* desc,
                          const int32_t* rows, int pinned);
/* Replica for another GPU of the box: copy the arena of `src` (descriptors, per-row constants, and the features if they
 * were set) into `ctx` device-to-device -- over NVLink / NVSwitch when the two GPUs are peers -- instead of reading the
 * collection over PCIe once per GPU.  `src` must have finished its upload and must not upload or be destroyed meanwhile;
 * it may be matching.  Replaces any image set of `ctx`. */
int mvgcuda_clone_images(mvgcuda_ctx* ctx, const mvgcuda_ctx* src);

/* Streaming form of the upload, for a loader that parses files while earlier images already travel
 * (apps/compute_matches: the reference loads every file before matching, matcher_all_in_memory.h:44-60):
 *   stream_begin  lays out the arena for n_images images of the given row counts (replaces any previous set);
 *   stream_image  enqueues the copy of ONE image's descriptors ([rows][128] u8) and, optionally, its feature
 *                 coordinates ([rows][2] float, NULL = none) plus the per-row constants kernel for its rows -- it does
 *                 not wait, so the buffers (page-locked for a truly asynchronous copy) must stay untouched until
 *                 stream_end or the next synchronising call; images may arrive in any order, each exactly once;
 *   stream_end    waits for every copy.  Matching calls are stream-ordered after the copies even without it. */
/* Page-locked host memory for the staging buffers of stream_image (usable by every GPU of the box). */
int mvgcuda_host_alloc(size_t bytes, void** out);
void mvgcuda_host_free(void* p);
int mvgcuda_stream_begin(mvgcuda_ctx* ctx, int n_images, const int32_t* rows);
int mvgcuda_stream_image(mvgcuda_ctx* ctx, int image, const uint8_t* desc, const float* feats_xy);
/* The same for a list of images in one call (desc[k] / feats_xy[k] belong to image images[k]; feats_xy may be NULL). */
int mvgcuda_stream_images(mvgcuda_ctx* ctx, int count, const int32_t* images, const uint8_t* const* desc,
                          const float* const* feats_xy);
int mvgcuda_stream_end(mvgcuda_ctx* ctx);
int mvgcuda_num_images(const mvgcuda_ctx* ctx);
int mvgcuda_image_rows(const mvgcuda_ctx* ctx, int image); /* <0 on bad id */

/* ------------------------------------------------------------------------------------------
 * Array level: 2 nearest neighbours of every row of image q_img among the rows of image
 * db_img == ArrayMatcherBruteForce<uchar,SquaredEuclideanDistanceVectorized<uchar>>::
 * SearchNeighbours(query, nq, &idx, &dist, 2)  (matcher_brute_force.h:102-134, metric.h:51-82).
 * idx/dist are [nq][2], distances ascending; dist holds exact integers as float (max
 * 128*255^2 < 2^24).  Returns MVGCUDA_ERR_INVALID when db rows < 2 or nq < 1 -- the reference
 * prints "Too much asked nearest neighbors" and returns false (matcher_brute_force.h:107-110). */
int mvgcuda_knn2(mvgcuda_ctx* ctx, int db_img, int q_img, int tie_mode, int32_t* idx, float* dist);

/* Same, on caller arrays (Build + SearchNeighbours in one call; uses a scratch slot of the
 * context, the uploaded image set is left untouched). */
int mvgcuda_knn2_arrays(mvgcuda_ctx* ctx, const uint8_t* db, int db_rows, const uint8_t* query,
                        int q_rows, int tie_mode, int32_t* idx, float* dist);

/* A database that stays in HBM == ArrayMatcherBruteForce::Build (matcher_brute_force.h:42-50; the reference borrows the
 * caller's pointer, this COPIES the rows to the GPU once) followed by any number of SearchNeighbours calls, each of
 * which uploads only its queries.  rows >= 1 (db_knn2 needs rows >= 2 like mvgcuda_knn2).  A db belongs to the context
 * it was created on; calls on one context are serialised by the caller. */
typedef struct mvgcuda_db mvgcuda_db;
int mvgcuda_db_create(mvgcuda_ctx* ctx, const uint8_t* db, int rows, mvgcuda_db** out);
void mvgcuda_db_destroy(mvgcuda_ctx* ctx, mvgcuda_db* db);
int mvgcuda_db_rows(const mvgcuda_db* db);
int mvgcuda_db_knn2(mvgcuda_ctx* ctx, const mvgcuda_db* db, const uint8_t* query, int q_rows, int tie_mode,
                    int32_t* idx, float* dist);

/* ------------------------------------------------------------------------------------------
 * Pair level: for each pair (I=pairs[2p], J=pairs[2p+1]) run, on the GPU,
 *   SearchNeighbours(J against I, k=2)                       matcher_all_in_memory.h:107
 *   DistanceRatioFilter(..., ratio_sq)                        :111-115  (fp32: d1 < ratio_sq*d2)
 *   the drop-last loop (last passing query discarded)         :117-122
 *   IndexedMatch::getDeduplicated (unique on consecutive _i)  :125
 * and return per pair the surviving (_i,_j) in ascending _j.  ratio_sq must be the fp32 value
 * the reference computes, Square(float distRatio) (numeric.h:108-111): pass r*r evaluated in
 * float, NOT float(double(r)*r).  Pairs whose images have < 2 db rows or < 1 query rows yield
 * zero matches (reference: SearchNeighbours returns false / empty, Appendix B of SURVEY.md).
 * The surviving matches are bit-identical to the reference's for every ratio.  Internally, for ratio_sq <= 1 the kernel
 * prunes rows that can change neither a ratio-test decision nor a reported index ("ratio-aware
 * pruning"); the exact 2 nearest neighbours of EVERY query are what mvgcuda_knn2 / mvgcuda_knn2_arrays return.
 *
 * Result storage is owned by the context and valid until the next match_* / upload / destroy:
 *   counts[p]            matches of pair p
 *   offsets[p]           start of pair p in `matches` (in matches, not bytes); offsets[n_pairs]=total
 *   matches[2*k+0/1]     _i (row in I), _j (row in J)
 */
typedef struct mvgcuda_pair_matches {
  int64_t n_pairs;
  const int32_t* counts;
  const int64_t* offsets;
  const int32_t* matches;
  /* device time of the GPU work of this call (CUDA events on the context's stream), ms */
  float gpu_ms;
  /* of which the fused distance/top-2 kernel */
  float knn_kernel_ms;
  int32_t knn_kernel_launches;
  int32_t total_launches;
  /* queries whose pruned record could not decide the ratio test and were matched a second time, exactly */
  int64_t rescanned_queries;
} mvgcuda_pair_matches;

int mvgcuda_match_pairs(mvgcuda_ctx* ctx, int64_t n_pairs, const int32_t* pairs, float ratio_sq,
                        mvgcuda_pair_matches* out);

/* ------------------------------------------------------------------------------------------
 * Collection level == MatcherAllInMemory<KeypointSet<ScalePointFeature,Descriptor<uchar,128>>,
 * ArrayMatcherBruteForce<...>>::Match  (matcher_all_in_memory.h:62-141) restricted to the given
 * pair list: match_pairs as above, then the coordinate de-duplication
 * IndexedMatchDecorator<float>::getDeduplicated (indexed_match_decorator.h:33-53,90-104) using the features' (x,y) --
 * on the GPU, one thread per pair building the red-black tree libstdc++'s std::set would build (the reference's
 * comparator is not a strict weak order, so the container's algorithm defines the result).  feats_xy[i] is
 * [rows_i][2] float (x,y of ScalePointFeature, feature.h:79-113) for every uploaded image; they are copied to HBM.
 * Result layout as for match_pairs (order within a pair is the std::set iteration order).
 * host_threads is ignored (before version 200 the de-duplication ran on a host pool). */
int mvgcuda_set_features(mvgcuda_ctx* ctx, int n_images, const float* const* feats_xy,
                         const int32_t* rows);
int mvgcuda_match_collection(mvgcuda_ctx* ctx, int64_t n_pairs, const int32_t* pairs,
                             float ratio_sq, int host_threads, mvgcuda_pair_matches* out);

/* Write the result of the last match_pairs / match_collection call in the reference's text format
 * "i j\ncount\n_i _j\n..." in lexicographic (i,j) order == PairedIndexedMatchToStream
 * (indexed_match_utils.h:22-38).  `pairs` must be the list given to that call. */
int mvgcuda_export_matches(mvgcuda_ctx* ctx, const int32_t* pairs, const char* path);

/* ------------------------------------------------------------------------------------------
 * Geometric filter: the step of apps/compute_matches right after putative matching (compute_matches.cpp:250-318) ==
 * ImageCollectionGeometricFilter<ScalePointFeature>::Filter(GeometricFilter_FMatrix_AC(precision, iterations), putatives,
 * geometric, image sizes)  (geometric_filter.h:37-102, fundamental_acransac.h:13-57): per pair an a-contrario RANSAC
 * (ACRANSAC, estimator_acransac.h:125-245) over the 7-point fundamental-matrix solver
 * (solver_fundamental_kernel.cpp:11-66), point-to-epipolar-line residuals, upper bound `precision` pixels (4.0 in
 * compute_matches.cpp:254), `iterations` = 4096 (fundamental_acransac.h:18); pairs whose model is not meaningful or has
 * fewer than 2.5 x 7 inliers come back empty.  On the GPU the control flow, the sample stream (glibc rand(), `seed` = 1 ==
 * the reference's never-seeded default, consumed pair after pair in the given order) and the double-precision arithmetic
 * of the solver are the reference's; see DESIGN.md for what is bit-exact and what is not.
 *   model        'f' = GeometricFilter_FMatrix_AC, or 'h' = GeometricFilter_HMatrix_AC (homography_acransac.h:18-62: 4-point
 *                solver, point-to-point residual, pairs with fewer than 2.5 x 4 inliers dropped; every bit of it is
 *                reproduced on the device, no host values involved); the essential variant is not built
 *   pairs/counts/offsets/matches   the putative matches, laid out like mvgcuda_pair_matches (e.g. the result of
 *                mvgcuda_match_collection, or an imported matches.putative.txt); pairs in std::map order
 *   image_sizes  [n_images][2] = width, height of every uploaded image (lists.txt columns 2, 3)
 * A pair may hold at most 16,384 putative matches (MVGCUDA_ERR_INVALID beyond).
 * The feature coordinates must have been set (mvgcuda_set_features / mvgcuda_stream_image).  The result (matches of every
 * pair in ascending-residual order, as the reference stores them) is owned by the context and valid until the next
 * geometric_filter / destroy; out->rescanned_queries holds the number of rand() values consumed,
 * out->knn_kernel_launches the number of models re-evaluated with roots from the host's C library.
 * Scheduling (pairs in flight on their own streams, pairs started ahead on the assumption that their predecessors find no
 * model) never shows in the result: a start that turns out to be wrong is repeated from the right rand() offset.  The
 * calling thread polls page-locked memory while the GPU works.  MVGCUDA_GEO_STATS=1 in the environment prints one line of
 * counters per call to stderr (launches, re-evaluated models, refuted starts, rand() position); MVGCUDA_GEO_BATCH_PAIRS=n
 * (tests) cuts the pair list into batches of n pairs instead of 1,024. */
int mvgcuda_geometric_filter(mvgcuda_ctx* ctx, char model, double precision, int iterations, unsigned seed,
                             int64_t n_pairs, const int32_t* pairs, const int32_t* counts, const int64_t* offsets,
                             const int32_t* matches, const int32_t* image_sizes, mvgcuda_pair_matches* out);

/* "i j\ncount\n_i _j\n..." for the given lists (PairedIndexedMatchToStream, indexed_match_utils.h:22-38); skip_empty != 0
 * omits pairs without matches, as a PairWiseMatches map that never received them does (geometric_filter.h:85-98). */
int mvgcuda_write_matches(const char* path, int64_t n_pairs, const int32_t* pairs, const int32_t* counts,
                          const int64_t* offsets, const int32_t* matches, int skip_empty);

/* ------------------------------------------------------------------------------------------
 * Instrumentation. */
/* The geometric filter's scalar solver core run on the device, one thread per case: x1/x2 [n][7][2] sampled (normalised)
 * correspondences -> F [n][3][9] models + their number; probe [n][4] = (x1, y1, x2, y2) -> residual under the first
 * model; nfa = one NFA term on that residual.  Lets the tests compare device bits with the reference's. */
int mvgcuda_geo_selftest(mvgcuda_ctx* ctx, int n, const double* x1, const double* x2, const double* probe, double* F,
                         int32_t* n_models, double* err, double* nfa);
/* The homography path as the filter runs it (one warp per case): x1/x2 [n][4][2] -> H [n][9] (row-major, unnormalised
 * null vector of the action matrix); probe [n][4] -> AsymmetricError under it. */
int mvgcuda_geo_selftest_h(mvgcuda_ctx* ctx, int n, const double* x1, const double* x2, const double* probe, double* H, double* err);

typedef struct mvgcuda_device_info {
  char name[128];
  int sm_count;
  int cc_major, cc_minor;
  int clock_khz;
  int64_t hbm_bytes;
} mvgcuda_device_info;
int mvgcuda_get_device_info(const mvgcuda_ctx* ctx, mvgcuda_device_info* out);

/* Tensor-pipe ceiling probe: issue `iters` back-to-back tcgen05.mma kind::i8 M128xN256xK32
 * instructions per CTA on every SM with no loads and no epilogue; returns achieved int8 op/s
 * (2 ops per MAC).  Used by bench.py as the measured int8 roofline denominator. */
int mvgcuda_probe_i8_peak(mvgcuda_ctx* ctx, int iters, double* ops_per_sec, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* MVGCUDA_H_ */
