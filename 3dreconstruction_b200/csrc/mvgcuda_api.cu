// mvgcuda_api.cu -- C ABI of libmvgcuda (include/mvgcuda.h): context, HBM arena, batch pipeline.
//
// Host-side structure (B200-first, not a translation of the reference's per-pair loop,
// matcher_all_in_memory.h:71-139): all descriptor arrays (and feature coordinates) are resident in one HBM arena, the
// pair list is cut into batches, each batch is ONE persistent launch of the fused kernel over (pair, 256-query) work
// items, a second small launch over the queries the pruned pass could not decide (planned on the device), and a handful
// of compaction kernels -- including the reference's coordinate de-duplication, one thread per pair.  Batches are
// software-pipelined over two buffer slots: while batch k computes, batch k-1's matches travel back over PCIe on a
// second stream; the host never synchronises with the compute stream inside a batch.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cfloat>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/mvgcuda.h"
#include "kernels.cuh"
#include "host_util.cuh"

namespace mvgcuda {

static thread_local char g_create_error[512] = "";

#define CU_CHECK(ctx, expr)                                                                         \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      (ctx)->set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return _e == cudaErrorMemoryAllocation ? MVGCUDA_ERR_NOMEM : MVGCUDA_ERR_CUDA;                \
    }                                                                                               \
  } while (0)

constexpr float kDefaultPruneRho = 0.72f;  // profiles/r02w_rho_sweep.log: kernel time vs size of the second pass
constexpr int kDefaultRescanRows = 1 << 20;  // 128 MB of gathered queries per second pass

struct Arena {
  DevBuf<uint8_t> desc;   // [rows_padded][128]
  DevBuf<int> ccol;       // K1 output (ccol_ints)
  DevBuf<float2> feat;    // [rows_padded] (x, y) of every descriptor row (collection level)
  DevBuf<int> img_row0, img_rows;
  std::vector<int> row0, rows;  // host copies
  int arena_rows = 0;
  bool has_feats = false;
  // streaming upload: image i was handed over (streamed), its copy + constants have been seen complete (ready), and the
  // event that marks their completion on the upload stream
  std::vector<char> streamed, ready;
  std::vector<cudaEvent_t> img_ev;
  CUtensorMap tmap_q;    // boxes of 128 rows: one CTA's query block
  CUtensorMap tmap_db;   // boxes of 64 rows: one CTA's share of an N = 128 MMA group
  void release() {
    desc.release(); ccol.release(); feat.release(); img_row0.release(); img_rows.release();
    row0.clear(); rows.clear(); arena_rows = 0; has_feats = false;
    for (cudaEvent_t e : img_ev) if (e) cudaEventDestroy(e);
    img_ev.clear(); streamed.clear(); ready.clear();
  }
};

// Buffers of one in-flight batch that the copy stream / the host still read while the next batch computes.
struct BatchSlot {
  DevBuf<PairJob> d_jobs;
  DevBuf<int> d_item_start;
  DevBuf<KnnRecord> d_knn;        // kept per slot so that a batch can be repaired after the next one was enqueued
  DevBuf<int> d_resc_idx, d_resc_cnt;
  DevBuf<RescanMeta> d_meta;
  DevBuf<int2> d_matches;         // final matches of the batch (dense, batch-local offsets)
  DevBuf<int> d_counts;           // per pair
  DevBuf<long long> d_offsets;    // [nb + 1], batch-local
  PinnedBuf<PairJob> h_jobs;
  PinnedBuf<int> h_item_start;
  PinnedBuf<int> h_counts;
  PinnedBuf<long long> h_offsets;
  PinnedBuf<RescanMeta> h_meta;
  PinnedBuf<int> h_resc_cnt;
  cudaEvent_t ev_start = nullptr, ev_knn = nullptr, ev_ready = nullptr, ev_done = nullptr, ev_copied = nullptr;
  // bookkeeping of the batch currently in the slot
  int64_t p0 = 0;
  int nb = 0, n_items = 0, launches = 0;
  long long n_records = 0;
  bool busy = false;
  void release() {
    d_jobs.release(); d_item_start.release(); d_knn.release(); d_resc_idx.release(); d_resc_cnt.release(); d_meta.release();
    d_matches.release(); d_counts.release(); d_offsets.release();
    h_jobs.release(); h_item_start.release(); h_counts.release(); h_offsets.release(); h_meta.release(); h_resc_cnt.release();
    for (cudaEvent_t* e : {&ev_start, &ev_knn, &ev_ready, &ev_done, &ev_copied}) if (*e) { cudaEventDestroy(*e); *e = nullptr; }
  }
};

}  // namespace mvgcuda

using namespace mvgcuda;

struct mvgcuda_db {  // one resident database image of the array level (ArrayMatcher::Build)
  Arena arena;
  int rows = 0;
};

struct mvgcuda_ctx {
  int device = 0;
  cudaDeviceProp prop{};
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;       // compute (own_stream, or the caller's)
  cudaStream_t copy_stream = nullptr;  // D2H of finished batches
  cudaStream_t post_stream = nullptr;  // rows 10-13 of batch k (small CTAs) under the fused kernel of batch k + 1
  cudaStream_t upload_stream = nullptr;  // H2D of streamed images (+ their constants kernel), concurrent with matching
  char error[1024] = "";
  size_t knn_smem = 0;

  Arena images;   // uploaded collection
  Arena scratch;  // operands of the array-level calls (queries; db of knn2_arrays)

  BatchSlot slot[2];
  // scratch shared by all batches (only used inside the kernel sequence of one batch, which the stream serialises)
  DevBuf<KnnItem> d_items, d_ritems;
  DevBuf<int2> d_tmp, d_tmp2;
  DevBuf<int> d_npass, d_counts_raw, d_counts2;
  DevBuf<long long> d_offsets_raw, d_total;
  DevBuf<RbNode> d_nodes;
  DevBuf<int> d_order;
  // second pass: gathered queries + work lists
  float prune_rho = kDefaultPruneRho;
  int rescan_cap_rows = kDefaultRescanRows;
  DevBuf<int> d_resc_ccol, d_ritem_start, d_used, d_job_of;
  DevBuf<uint8_t> d_resc_desc;
  DevBuf<KnnRecord> d_resc_knn;
  DevBuf<PairJob> d_rjobs;
  DevBuf<RescanSrc> d_rsrc;
  PinnedBuf<int> h_ritem_start;
  PinnedBuf<PairJob> h_rjobs;
  PinnedBuf<RescanSrc> h_rsrc;
  CUtensorMap tmap_resc;
  const uint8_t* tmap_resc_base = nullptr;
  int tmap_resc_rows = 0;
  long long rescanned = 0;  // queries matched a second time by the last match call
  int repaired_batches = 0; // batches whose second pass overflowed the gather buffer and went through the bounded path

  // results of the last match call (pinned host memory)
  PinnedBuf<int> r_counts;
  PinnedBuf<long long> r_offsets;
  PinnedBuf<int> r_matches;  // 2 ints per match
  int64_t r_pairs = 0;

  cudaEvent_t ev[2] = {nullptr, nullptr};

  void* geo = nullptr;               // state of the geometric-filter unit (geometric_api.cu)
  void (*geo_free)(void*) = nullptr;

  void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error, sizeof error, fmt, ap);
    va_end(ap);
  }
};

namespace mvgcuda {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D u8 tensor [rows][128], box {128 bytes, box_rows}, 128-byte swizzle (one row == one swizzle atom).
static int make_tmap(mvgcuda_ctx* ctx, CUtensorMap* tm, const uint8_t* base, int rows, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { ctx->set_error("cuTensorMapEncodeTiled entry point not found"); return MVGCUDA_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)kDim, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)kDim};
  cuuint32_t box[2] = {(cuuint32_t)kDim, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ctx->set_error("cuTensorMapEncodeTiled failed: CUresult %d", (int)r); return MVGCUDA_ERR_CUDA; }
  return MVGCUDA_OK;
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Lay the images out in the arena (each starting at a multiple of kRowAlign rows), allocate, zero the descriptor
// rows, upload the image table and build the tensor maps.  No descriptor is copied yet.
static int layout_arena(mvgcuda_ctx* ctx, Arena& A, int n_images, const int32_t* rows) {
  A.row0.assign(n_images + 1, 0);
  A.rows.assign(n_images, 0);
  A.has_feats = false;
  long long total = 0;
  for (int i = 0; i < n_images; ++i) {
    if (rows[i] < 0) { ctx->set_error("image %d: negative row count", i); return MVGCUDA_ERR_INVALID; }
    A.row0[i] = (int)total;
    A.rows[i] = rows[i];
    total += round_up(rows[i], kRowAlign);
    if (total > 0x7FFF0000ll) { ctx->set_error("arena exceeds 2^31 rows"); return MVGCUDA_ERR_INVALID; }
  }
  A.row0[n_images] = (int)total;
  A.arena_rows = (int)std::max<long long>(total, kRowAlign);
  CU_CHECK(ctx, A.desc.reserve((size_t)A.arena_rows * kDim));
  CU_CHECK(ctx, A.ccol.reserve(ccol_ints(A.arena_rows)));
  CU_CHECK(ctx, A.img_row0.reserve(n_images + 1));
  CU_CHECK(ctx, A.img_rows.reserve(std::max(n_images, 1)));
  cudaStream_t st = ctx->stream;
  CU_CHECK(ctx, cudaMemsetAsync(A.desc.p, 0, (size_t)A.arena_rows * kDim, st));
  // per-128-row minima are accumulated with atomicMin by K1: preset them above every possible norm
  CU_CHECK(ctx, cudaMemsetAsync(A.ccol.p + (size_t)(A.arena_rows / kTileDb) * kTileC, 0x7F, (size_t)(A.arena_rows / kHalfCols) * sizeof(int), st));
  A.streamed.assign(n_images, 0);
  A.ready.assign(n_images, 1);  // nothing pending (the bulk upload is synchronous; stream_image clears the flag of its image)
  CU_CHECK(ctx, cudaMemcpyAsync(A.img_row0.p, A.row0.data(), (n_images + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  if (n_images)
    CU_CHECK(ctx, cudaMemcpyAsync(A.img_rows.p, A.rows.data(), n_images * sizeof(int), cudaMemcpyHostToDevice, st));
  int rc = make_tmap(ctx, &A.tmap_q, A.desc.p, A.arena_rows, kBlockQ);
  if (rc) return rc;
  rc = make_tmap(ctx, &A.tmap_db, A.desc.p, A.arena_rows, 64);
  if (rc) return rc;
  CU_CHECK(ctx, cudaStreamSynchronize(st));  // the host vectors are read by the async copies above
  return MVGCUDA_OK;
}

// K1 over arena rows [row_begin, row_end) (multiples of 256)
static int launch_k1(mvgcuda_ctx* ctx, Arena& A, int row_begin, int row_end, cudaStream_t st) {
  if (row_end <= row_begin) return MVGCUDA_OK;
  row_consts_kernel<<<(row_end - row_begin) / kK1Rows, 8 * kK1Rows, 0, st>>>(
      A.desc.p, A.img_row0.p, A.img_rows.p, (int)A.rows.size(), A.arena_rows, row_begin, A.ccol.p);
  CU_CHECK(ctx, cudaGetLastError());
  return MVGCUDA_OK;
}

// Whole collection at once: layout, copy every image in (host or device pointers), K1.
static int fill_arena(mvgcuda_ctx* ctx, Arena& A, int n_images, const uint8_t* const* desc, const int32_t* rows) {
  for (int i = 0; i < n_images; ++i)
    if (rows[i] > 0 && !desc[i]) { ctx->set_error("image %d: null descriptor pointer", i); return MVGCUDA_ERR_INVALID; }
  int rc = layout_arena(ctx, A, n_images, rows);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  for (int i = 0; i < n_images; ++i) {
    if (rows[i] == 0) continue;
    // cudaMemcpyDefault: desc[i] may be host memory (pageable or pinned) or device memory (e.g. a replica that
    // arrived over NVLink) -- unified addressing sorts it out
    CU_CHECK(ctx, cudaMemcpyAsync(A.desc.p + (size_t)A.row0[i] * kDim, desc[i], (size_t)rows[i] * kDim, cudaMemcpyDefault, st));
  }
  rc = launch_k1(ctx, A, 0, A.arena_rows, st);
  if (rc) return rc;
  CU_CHECK(ctx, cudaStreamSynchronize(st));  // the caller's buffers are free again
  return MVGCUDA_OK;
}

static int upload_features(mvgcuda_ctx* ctx, Arena& A, int image, const float* xy, int rows, cudaStream_t st) {
  if (rows == 0) return MVGCUDA_OK;
  CU_CHECK(ctx, cudaMemcpyAsync(A.feat.p + A.row0[image], xy, (size_t)rows * sizeof(float2), cudaMemcpyDefault, st));
  return MVGCUDA_OK;
}

// Make the compute stream wait for the streamed images a pair list touches (and only for those): matching of the pairs
// among the images that have arrived overlaps the upload of the rest.
static int wait_for_images(mvgcuda_ctx* ctx, Arena& A, const int32_t* pairs, int64_t p0, int64_t p1, std::vector<char>& seen) {
  seen.assign(A.rows.size(), 0);
  for (int64_t p = p0; p < p1; ++p) { seen[pairs[2 * p]] = 1; seen[pairs[2 * p + 1]] = 1; }
  for (size_t i = 0; i < seen.size(); ++i) {
    if (!seen[i] || A.ready[i]) continue;
    if (!A.streamed[i]) { ctx->set_error("image %d was announced by stream_begin but never streamed", (int)i); return MVGCUDA_ERR_INVALID; }
    const cudaError_t q = cudaEventQuery(A.img_ev[i]);
    if (q == cudaSuccess) { A.ready[i] = 1; continue; }
    if (q != cudaErrorNotReady) CU_CHECK(ctx, q);
    (void)cudaGetLastError();
    CU_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, A.img_ev[i], 0));
  }
  return MVGCUDA_OK;
}

static int wait_for_all_images(mvgcuda_ctx* ctx, Arena& A) {
  for (size_t i = 0; i < A.ready.size(); ++i) {
    if (A.ready[i] || !A.streamed[i]) continue;
    CU_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, A.img_ev[i], 0));
  }
  return MVGCUDA_OK;
}

struct BatchPlan {
  int n_jobs = 0;
  int n_items = 0;
  long long n_records = 0;
};

// Fill jobs / item_start for pairs [p0, p1) of the list; db images from arena A, query images from arena Q.
static void plan_batch(PairJob* jobs, int* item_start, const Arena& A, const Arena& Q, const int32_t* pairs, int64_t p0, int64_t p1,
                       BatchPlan& bp) {
  bp.n_jobs = (int)(p1 - p0);
  long long rec = 0;
  int items = 0;
  for (int64_t p = p0; p < p1; ++p) {
    const int I = pairs[2 * p], J = pairs[2 * p + 1];
    PairJob& j = jobs[p - p0];
    j.db_row0 = A.row0[I];
    j.db_rows = A.rows[I];
    j.q_row0 = Q.row0[J];
    j.q_rows = Q.rows[J];
    j.out_off = (int)rec;
    j.valid = (j.db_rows >= 2 && j.q_rows >= 1) ? 1 : 0;
    item_start[p - p0] = items;
    if (j.valid) items += (j.q_rows + 2 * kBlockQ - 1) / (2 * kBlockQ);  // a CTA pair takes two query blocks per item
    rec += j.q_rows;
  }
  item_start[p1 - p0] = items;
  bp.n_items = items;
  bp.n_records = rec;
}

// Work list + one persistent launch of K2.  tmap_q / qcol: the query rows (boxes of 128); A: the arena of the db rows.
// n_items_max bounds the device-side count when kp.meta is set.
static int launch_knn_raw(mvgcuda_ctx* ctx, const CUtensorMap& tmap_q, const Arena& A, KnnParams kp, const PairJob* d_jobs,
                          const int* d_item_start, int n_jobs, int n_items_max, DevBuf<KnnItem>& d_items) {
  if (n_items_max == 0) return MVGCUDA_OK;
  CU_CHECK(ctx, d_items.reserve(n_items_max));
  build_items_kernel<<<(n_items_max + 255) / 256, 256, 0, ctx->stream>>>(d_jobs, d_item_start, n_jobs, kp.n_items, kp.meta, d_items.p);
  CU_CHECK(ctx, cudaGetLastError());
  kp.items = d_items.p;
  kp.ccol = A.ccol.p;
  kp.hmin = A.ccol.p + (size_t)(A.arena_rows / kTileDb) * kTileC;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kKnnThreads);
  cfg.gridDim = dim3(2 * std::min(n_items_max, ctx->prop.multiProcessorCount / 2));  // CTA pairs
  cfg.dynamicSmemBytes = ctx->knn_smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CU_CHECK(ctx, cudaLaunchKernelEx(&cfg, knn2_kernel, tmap_q, A.tmap_db, kp));
  return MVGCUDA_OK;
}

constexpr long long kBatchRecords = 8ll << 20;  // queries per batch (128 MB of KnnRecord per slot)
constexpr int kBatchPairs = 1 << 16;

static int validate_pairs(mvgcuda_ctx* ctx, const Arena& A, int64_t n_pairs, const int32_t* pairs) {
  if (n_pairs < 0 || n_pairs > ((int64_t)1 << 40) || (n_pairs > 0 && !pairs)) { ctx->set_error("bad pair list"); return MVGCUDA_ERR_INVALID; }
  const int n = (int)A.rows.size();
  for (int64_t p = 0; p < n_pairs; ++p) {
    const int I = pairs[2 * p], J = pairs[2 * p + 1];
    if (I < 0 || I >= n || J < 0 || J >= n) { ctx->set_error("pair %lld: image id out of range", (long long)p); return MVGCUDA_ERR_INVALID; }
  }
  return MVGCUDA_OK;
}

// The admission factor actually used for a ratio: never below the ratio itself (a row with d < ratio * d(best) could be
// a passing nearest neighbour and must always be admitted), 1 = plain best-distance bound, no second pass.
static float effective_rho(const mvgcuda_ctx* ctx, float ratio_sq) {
  float rho = ctx->prune_rho;
  if (!(rho < 1.0f)) return 1.0f;
  return std::min(1.0f, std::max(rho, ratio_sq * 1.002f));
}

// Gather buffer of the second pass (rows of gathered queries) and its tensor map.
static int reserve_rescan(mvgcuda_ctx* ctx, int cap, int nb) {
  const size_t ccol_len = (size_t)(cap / kTileDb + 2) * kTileC;
  CU_CHECK(ctx, ctx->d_resc_desc.reserve((size_t)(cap + 2 * kBlockQ) * kDim));
  CU_CHECK(ctx, ctx->d_resc_ccol.reserve(ccol_len));
  CU_CHECK(ctx, ctx->d_resc_knn.reserve(cap));
  const int tm_rows = (int)(ctx->d_resc_desc.cap / kDim);
  if (ctx->tmap_resc_base != ctx->d_resc_desc.p || ctx->tmap_resc_rows != tm_rows) {
    int rc = make_tmap(ctx, &ctx->tmap_resc, ctx->d_resc_desc.p, tm_rows, kBlockQ);
    if (rc) return rc;
    ctx->tmap_resc_base = ctx->d_resc_desc.p;
    ctx->tmap_resc_rows = tm_rows;
  }
  CU_CHECK(ctx, ctx->d_rjobs.reserve(nb + 1));
  CU_CHECK(ctx, ctx->d_rsrc.reserve(nb + 1));
  CU_CHECK(ctx, ctx->d_ritem_start.reserve(nb + 2));
  CU_CHECK(ctx, ctx->d_used.reserve(nb + 2));
  CU_CHECK(ctx, ctx->d_job_of.reserve(nb + 1));
  return MVGCUDA_OK;
}

// Second pass over the ambiguous queries of the batch in slot S, planned ON THE DEVICE (no host round trip): flag,
// plan, gather, the same fused kernel with exact records (prune_ratio = FLT_MAX), scatter.  If the batch has more
// ambiguous queries than the gather buffer holds, the plan kernel raises `overflow` and plans nothing; finish_batch
// then repairs the batch with rescan_bounded below.
static int rescan_planned(mvgcuda_ctx* ctx, const Arena& A, BatchSlot& S, float ratio_sq, float rho) {
  cudaStream_t st = ctx->stream;
  const int nb = S.nb;
  const int cap = std::max(ctx->rescan_cap_rows, 1);
  int rc = reserve_rescan(ctx, cap, nb);
  if (rc) return rc;
  flag_ambiguous_kernel<<<nb, kCompactThreads, 0, st>>>(S.d_jobs.p, S.d_knn.p, ratio_sq, rho, S.d_resc_idx.p, S.d_resc_cnt.p);
  CU_CHECK(ctx, cudaGetLastError());
  plan_rescan_kernel<<<1, 1024, 0, st>>>(S.d_jobs.p, S.d_resc_cnt.p, nb, cap, ctx->d_rsrc.p, ctx->d_rjobs.p, ctx->d_ritem_start.p,
                                         ctx->d_used.p, ctx->d_job_of.p, S.d_meta.p);
  CU_CHECK(ctx, cudaGetLastError());
  rescan_gather_kernel<<<nb, 256, 0, st>>>(ctx->d_rsrc.p, S.d_resc_idx.p, A.desc.p, A.ccol.p, ctx->d_resc_desc.p, ctx->d_resc_ccol.p);
  CU_CHECK(ctx, cudaGetLastError());
  KnnParams kp = {};
  kp.qcol = ctx->d_resc_ccol.p;
  kp.meta = S.d_meta.p;
  kp.out = ctx->d_resc_knn.p;
  kp.prune_ratio = FLT_MAX;
  kp.prune_rho = 1.0f;
  const int n_items_max = cap / (2 * kBlockQ) + nb + 1;
  rc = launch_knn_raw(ctx, ctx->tmap_resc, A, kp, ctx->d_rjobs.p, ctx->d_ritem_start.p, 0, n_items_max, ctx->d_ritems);
  if (rc) return rc;
  rescan_scatter_kernel<<<nb, 256, 0, st>>>(ctx->d_rsrc.p, S.d_resc_idx.p, ctx->d_resc_knn.p, S.d_knn.p);
  CU_CHECK(ctx, cudaGetLastError());
  S.launches += 6;
  return MVGCUDA_OK;
}

// Bounded, host-planned form of the second pass: the flagged queries go through the gather buffer in as many rounds as
// needed (a pair may be split over rounds).  Synchronous; only runs for a batch whose device-planned pass overflowed
// (S.h_resc_cnt holds the per-pair counts, the flagged lists are still in S.d_resc_idx).
static int rescan_bounded(mvgcuda_ctx* ctx, const Arena& A, BatchSlot& S) {
  cudaStream_t st = ctx->stream;
  const int nb = S.nb;
  long long total = 0;
  for (int k = 0; k < nb; ++k) total += S.h_resc_cnt.p[k];
  if (total == 0) return MVGCUDA_OK;
  const int cap = (int)std::min<long long>(std::max(ctx->rescan_cap_rows, 1), round_up((int)std::min<long long>(total, 1 << 30), 256));
  int rc = reserve_rescan(ctx, cap, nb);
  if (rc) return rc;
  CU_CHECK(ctx, ctx->h_rjobs.reserve(nb + 1));
  CU_CHECK(ctx, ctx->h_rsrc.reserve(nb + 1));
  CU_CHECK(ctx, ctx->h_ritem_start.reserve(nb + 2));
  int j = 0, first = 0;
  while (j < nb) {
    // n_rs gather/scatter sources (one per original pair with flagged queries); n_rj second-pass jobs: consecutive
    // pairs against the same db image share ONE job
    int used = 0, n_rs = 0, n_rj = 0, items = 0;
    while (j < nb && used < cap) {
      const int left = S.h_resc_cnt.p[j] - first;
      if (left <= 0) { ++j; first = 0; continue; }
      const int take = std::min(left, cap - used);
      const PairJob& J = S.h_jobs.p[j];
      ctx->h_rsrc.p[n_rs++] = RescanSrc{J.out_off, J.q_row0, first, take, used};
      if (n_rj > 0 && ctx->h_rjobs.p[n_rj - 1].db_row0 == J.db_row0 && ctx->h_rjobs.p[n_rj - 1].db_rows == J.db_rows) {
        ctx->h_rjobs.p[n_rj - 1].q_rows += take;
      } else {
        PairJob& R = ctx->h_rjobs.p[n_rj];
        R = J;
        R.q_row0 = used;
        R.q_rows = take;
        R.out_off = used;
        R.valid = 1;
        ++n_rj;
      }
      used += take;
      first += take;
    }
    for (int k = 0; k < n_rj; ++k) {
      ctx->h_ritem_start.p[k] = items;
      items += (ctx->h_rjobs.p[k].q_rows + 2 * kBlockQ - 1) / (2 * kBlockQ);
    }
    if (n_rj == 0) break;
    ctx->h_ritem_start.p[n_rj] = items;
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->d_rjobs.p, ctx->h_rjobs.p, n_rj * sizeof(PairJob), cudaMemcpyHostToDevice, st));
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->d_rsrc.p, ctx->h_rsrc.p, n_rs * sizeof(RescanSrc), cudaMemcpyHostToDevice, st));
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->d_ritem_start.p, ctx->h_ritem_start.p, (n_rj + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    rescan_gather_kernel<<<n_rs, 256, 0, st>>>(ctx->d_rsrc.p, S.d_resc_idx.p, A.desc.p, A.ccol.p, ctx->d_resc_desc.p, ctx->d_resc_ccol.p);
    CU_CHECK(ctx, cudaGetLastError());
    KnnParams kp = {};
    kp.qcol = ctx->d_resc_ccol.p;
    kp.n_items = items;
    kp.out = ctx->d_resc_knn.p;
    kp.prune_ratio = FLT_MAX;
    kp.prune_rho = 1.0f;
    rc = launch_knn_raw(ctx, ctx->tmap_resc, A, kp, ctx->d_rjobs.p, ctx->d_ritem_start.p, n_rj, items, ctx->d_ritems);
    if (rc) return rc;
    rescan_scatter_kernel<<<n_rs, 256, 0, st>>>(ctx->d_rsrc.p, S.d_resc_idx.p, ctx->d_resc_knn.p, S.d_knn.p);
    CU_CHECK(ctx, cudaGetLastError());
    S.launches += 4;
    CU_CHECK(ctx, cudaStreamSynchronize(st));  // the pinned job lists are rewritten by the next round
  }
  return MVGCUDA_OK;
}

// Rows 10-13 for the batch in slot S: ratio test + ordered compaction + drop-last + unique-on-_i (K3), then -- at the
// collection level -- the coordinate de-duplication, one thread per pair, and a second compaction; finally the small
// per-pair results start their way to the slot's pinned staging.
static int compact_batch(mvgcuda_ctx* ctx, const Arena& A, BatchSlot& S, float ratio_sq, bool dedup_xy, cudaStream_t st, bool beside_knn) {
  const int nb = S.nb;
  const size_t nrec = (size_t)std::max<long long>(S.n_records, 1);
  CU_CHECK(ctx, ctx->d_tmp.reserve(nrec));
  CU_CHECK(ctx, ctx->d_npass.reserve(nb));
  CU_CHECK(ctx, ctx->d_total.reserve(2));
  CU_CHECK(ctx, S.d_matches.reserve(nrec));
  CU_CHECK(ctx, S.d_counts.reserve(nb));
  CU_CHECK(ctx, S.d_offsets.reserve(nb + 1));
  int* counts1 = S.d_counts.p;
  long long* offsets1 = S.d_offsets.p;
  int2* matches1 = S.d_matches.p;
  if (dedup_xy) {
    CU_CHECK(ctx, ctx->d_counts_raw.reserve(nb));
    CU_CHECK(ctx, ctx->d_offsets_raw.reserve(nb + 1));
    CU_CHECK(ctx, ctx->d_tmp2.reserve(nrec));
    counts1 = ctx->d_counts_raw.p;
    offsets1 = ctx->d_offsets_raw.p;
    matches1 = ctx->d_tmp2.p;
  }
  ratio_filter_kernel<<<nb, kCompactThreads, 0, st>>>(S.d_jobs.p, S.d_knn.p, ratio_sq, ctx->d_tmp.p, ctx->d_npass.p, counts1);
  CU_CHECK(ctx, cudaGetLastError());
  scan_counts_kernel<<<1, kScanThreads, 0, st>>>(counts1, nb, 0, offsets1, ctx->d_total.p);
  CU_CHECK(ctx, cudaGetLastError());
  dedup_scatter_kernel<<<nb, kCompactThreads, 0, st>>>(S.d_jobs.p, ctx->d_tmp.p, ctx->d_npass.p, offsets1, 0, matches1);
  CU_CHECK(ctx, cudaGetLastError());
  S.launches += 3;
  if (dedup_xy) {
    CU_CHECK(ctx, ctx->d_nodes.reserve(nrec + nb + 1));
    CU_CHECK(ctx, ctx->d_order.reserve(nrec));
    // survivors of pair p land in d_tmp at the pair's old offset, then move to their final (dense) place
    // a batch with a successor uses the small-footprint variant, which shares the SMs with the successor's fused kernel
    const int cap = beside_knn ? kDedupSmemCapSmall : kDedupSmemCap;
    if (beside_knn) dedup_xy_smem_kernel<kDedupSmemCapSmall><<<nb, 32, 0, st>>>(S.d_jobs.p, matches1, offsets1, counts1, A.feat.p, ctx->d_tmp.p, S.d_counts.p);
    else dedup_xy_smem_kernel<kDedupSmemCap><<<nb, 32, 0, st>>>(S.d_jobs.p, matches1, offsets1, counts1, A.feat.p, ctx->d_tmp.p, S.d_counts.p);
    CU_CHECK(ctx, cudaGetLastError());
    dedup_xy_kernel<<<(nb + 63) / 64, 64, 0, st>>>(S.d_jobs.p, nb, cap, matches1, offsets1, counts1, A.feat.p, ctx->d_nodes.p,
                                                  ctx->d_order.p, ctx->d_tmp.p, S.d_counts.p);
    CU_CHECK(ctx, cudaGetLastError());
    scan_counts_kernel<<<1, kScanThreads, 0, st>>>(S.d_counts.p, nb, 0, S.d_offsets.p, ctx->d_total.p + 1);
    CU_CHECK(ctx, cudaGetLastError());
    compact_pairs_kernel<<<nb, kCompactThreads, 0, st>>>(ctx->d_tmp.p, offsets1, S.d_offsets.p, S.d_counts.p, S.d_matches.p);
    CU_CHECK(ctx, cudaGetLastError());
    S.launches += 4;
  }
  CU_CHECK(ctx, S.h_counts.reserve(nb));
  CU_CHECK(ctx, S.h_offsets.reserve(nb + 1));
  CU_CHECK(ctx, cudaMemcpyAsync(S.h_counts.p, S.d_counts.p, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_CHECK(ctx, cudaMemcpyAsync(S.h_offsets.p, S.d_offsets.p, (nb + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
  return MVGCUDA_OK;
}

struct MatchRun {  // state of one match_pairs / match_collection call
  const int32_t* pairs = nullptr;
  int64_t n_pairs = 0;
  float ratio_sq = 0.f;
  bool dedup_xy = false;
  bool prune = false;
  float rho = 1.f;
  long long match_base = 0;  // matches of all finished batches
  float gpu_ms = 0.f, knn_ms = 0.f;
  int knn_launches = 0, launches = 0;
  std::vector<char> seen;  // scratch of wait_for_images
  long long total_records = 0, done_records = 0;  // queries of the whole call / of the finished batches
};

// Enqueue every kernel of the batch [p0, p1) into slot S.  No synchronisation with the compute stream.
static int enqueue_batch(mvgcuda_ctx* ctx, MatchRun& R, BatchSlot& S, int64_t p0, int64_t p1) {
  const Arena& A = ctx->images;
  cudaStream_t st = ctx->stream;
  const int nb = (int)(p1 - p0);
  CU_CHECK(ctx, S.h_jobs.reserve(nb));
  CU_CHECK(ctx, S.h_item_start.reserve(nb + 1));
  BatchPlan bp;
  plan_batch(S.h_jobs.p, S.h_item_start.p, A, A, R.pairs, p0, p1, bp);
  S.p0 = p0; S.nb = nb; S.n_items = bp.n_items; S.n_records = bp.n_records; S.launches = 0; S.busy = true;
  const size_t nrec = (size_t)std::max<long long>(bp.n_records, 1);
  CU_CHECK(ctx, S.d_jobs.reserve(nb));
  CU_CHECK(ctx, S.d_item_start.reserve(nb + 1));
  CU_CHECK(ctx, S.d_knn.reserve(nrec));
  CU_CHECK(ctx, S.d_meta.reserve(1));
  CU_CHECK(ctx, S.h_meta.reserve(1));
  // the slot's device buffers are free once the copy of the batch that used them last has finished
  CU_CHECK(ctx, cudaStreamWaitEvent(st, S.ev_copied, 0));
  int rc = wait_for_images(ctx, ctx->images, R.pairs, p0, p1, R.seen);
  if (rc) return rc;
  CU_CHECK(ctx, cudaMemcpyAsync(S.d_jobs.p, S.h_jobs.p, nb * sizeof(PairJob), cudaMemcpyHostToDevice, st));
  CU_CHECK(ctx, cudaMemcpyAsync(S.d_item_start.p, S.h_item_start.p, (nb + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  CU_CHECK(ctx, cudaMemsetAsync(S.d_meta.p, 0, sizeof(RescanMeta), st));
  CU_CHECK(ctx, cudaEventRecord(S.ev_start, st));
  // ratio <= 1: records only have to carry what the ratio test and the match list need (ratio-aware pruning);
  // ratio > 1 keeps the exact 2nd neighbour of every query for the tie fix-up below
  KnnParams kp = {};
  kp.qcol = A.ccol.p;
  kp.n_items = bp.n_items;
  kp.out = S.d_knn.p;
  kp.prune_ratio = R.prune ? R.ratio_sq : FLT_MAX;
  kp.prune_rho = R.rho;
  rc = launch_knn_raw(ctx, A.tmap_q, A, kp, S.d_jobs.p, S.d_item_start.p, nb, bp.n_items, ctx->d_items);
  if (rc) return rc;
  if (bp.n_items) S.launches += 2;
  if (R.prune && R.rho < 1.0f && bp.n_items) {
    CU_CHECK(ctx, S.d_resc_idx.reserve(nrec));
    CU_CHECK(ctx, S.d_resc_cnt.reserve(nb));
    CU_CHECK(ctx, S.h_resc_cnt.reserve(nb));
    rc = rescan_planned(ctx, A, S, R.ratio_sq, R.rho);
    if (rc) return rc;
    CU_CHECK(ctx, cudaMemcpyAsync(S.h_resc_cnt.p, S.d_resc_cnt.p, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  CU_CHECK(ctx, cudaMemcpyAsync(S.h_meta.p, S.d_meta.p, sizeof(RescanMeta), cudaMemcpyDeviceToHost, st));
  CU_CHECK(ctx, cudaEventRecord(S.ev_knn, st));  // K2 incl. the exact second pass over ambiguous queries
  if (R.ratio_sq > 1.0f && bp.n_items) {
    // For ratio > 1 a tie d1 == d2 passes the test, so WHICH of the tied rows is reported matters: reproduce the
    // reference's std::partial_sort choice exactly (one extra CUDA-core pass; the usual ratios <= 1 skip it).
    int max_q = 0;
    for (int k = 0; k < nb; ++k) if (S.h_jobs.p[k].valid) max_q = std::max(max_q, S.h_jobs.p[k].q_rows);
    for (int j0 = 0; j0 < nb && max_q > 0; j0 += 32768) {
      const int nj = std::min(32768, nb - j0);
      tie_fixup_kernel<<<dim3((max_q + 7) / 8, nj), 256, 0, st>>>(A.desc.p, A.desc.p, S.d_jobs.p, j0, S.d_knn.p);
      CU_CHECK(ctx, cudaGetLastError());
      ++S.launches;
    }
  }
  // Collection level: rows 10-13 on the post stream -- small CTAs that run beside the NEXT batch's fused kernel instead of
  // in front of it (the coordinate de-dup is a serial chain per pair: 0.6 ms of an otherwise idle GPU per batch).  Pair
  // level: the three compaction kernels take 0.1 ms and stay in line.
  cudaStream_t ps = R.dedup_xy ? ctx->post_stream : st;
  if (R.dedup_xy) {
    CU_CHECK(ctx, cudaEventRecord(S.ev_ready, st));
    CU_CHECK(ctx, cudaStreamWaitEvent(ps, S.ev_ready, 0));
  }
  rc = compact_batch(ctx, A, S, R.ratio_sq, R.dedup_xy, ps, R.dedup_xy && p1 < R.n_pairs);
  if (rc) return rc;
  CU_CHECK(ctx, cudaEventRecord(S.ev_done, ps));
  return MVGCUDA_OK;
}

// Wait for the batch in slot S, repair it if its second pass overflowed, and send its matches to the host on the copy
// stream (the compute stream is already busy with the next batch).
static int finish_batch(mvgcuda_ctx* ctx, MatchRun& R, BatchSlot& S) {
  if (!S.busy) return MVGCUDA_OK;
  const Arena& A = ctx->images;
  CU_CHECK(ctx, cudaEventSynchronize(S.ev_done));
  if (S.h_meta.p[0].overflow) {
    // more ambiguous queries than the gather buffer holds: everything else on the stream is drained, then this batch is
    // matched again through the bounded path (flagged lists and records are still in the slot)
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->post_stream));
    int rc = rescan_bounded(ctx, A, S);
    if (rc) return rc;
    rc = compact_batch(ctx, A, S, R.ratio_sq, R.dedup_xy, ctx->stream, false);
    if (rc) return rc;
    CU_CHECK(ctx, cudaEventRecord(S.ev_done, ctx->stream));
    CU_CHECK(ctx, cudaEventSynchronize(S.ev_done));
    ++ctx->repaired_batches;
  }
  ctx->rescanned += S.h_meta.p[0].total;
  const int nb = S.nb;
  const long long nm = S.h_offsets.p[nb];
  const long long new_total = R.match_base + nm;
  R.done_records += S.n_records;
  if ((size_t)std::max<long long>(new_total, 1) * 2 > ctx->r_matches.cap) {
    // Grow ONCE to what the match density so far predicts for the whole call (+15 %), not step by step: page-locking is
    // slow (~1 GB/s) and serialised across the GPUs of a box, and every step would copy what is already there.
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->copy_stream));  // copies into the old buffer finish before it moves
    const double density = R.done_records > 0 ? (double)new_total / (double)R.done_records : 0.0;
    const long long predicted = (long long)(density * (double)R.total_records * 1.15) + nm;
    CU_CHECK(ctx, ctx->r_matches.reserve((size_t)std::max<long long>(std::max(new_total, predicted), 1) * 2, (size_t)R.match_base * 2));
  }
  if (nm > 0)
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->r_matches.p + R.match_base * 2, S.d_matches.p, nm * sizeof(int2), cudaMemcpyDeviceToHost, ctx->copy_stream));
  CU_CHECK(ctx, cudaEventRecord(S.ev_copied, ctx->copy_stream));
  for (int k = 0; k < nb; ++k) {
    ctx->r_counts.p[S.p0 + k] = S.h_counts.p[k];
    ctx->r_offsets.p[S.p0 + k] = R.match_base + S.h_offsets.p[k];
  }
  float ms = 0.f;
  CU_CHECK(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], S.ev_done));  // batches overlap: time since the first one started
  R.gpu_ms = std::max(R.gpu_ms, ms);
  if (S.n_items) {
    CU_CHECK(ctx, cudaEventElapsedTime(&ms, S.ev_start, S.ev_knn));
    R.knn_ms += ms;
    ++R.knn_launches;
  }
  R.launches += S.launches;
  R.match_base = new_total;
  S.busy = false;
  return MVGCUDA_OK;
}

static int match_pairs_impl(mvgcuda_ctx* ctx, int64_t n_pairs, const int32_t* pairs, float ratio_sq, bool dedup_xy,
                            mvgcuda_pair_matches* out) {
  const Arena& A = ctx->images;
  int rc = validate_pairs(ctx, A, n_pairs, pairs);
  if (rc) return rc;
  if (dedup_xy && !A.has_feats && n_pairs > 0) { ctx->set_error("match_collection: call mvgcuda_set_features first"); return MVGCUDA_ERR_INVALID; }
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  CU_CHECK(ctx, ctx->r_counts.reserve(std::max<int64_t>(n_pairs, 1)));
  CU_CHECK(ctx, ctx->r_offsets.reserve(n_pairs + 1));
  CU_CHECK(ctx, ctx->r_matches.reserve(2));
  ctx->r_pairs = n_pairs;
  ctx->rescanned = 0;
  ctx->repaired_batches = 0;
  MatchRun R;
  R.pairs = pairs; R.n_pairs = n_pairs; R.ratio_sq = ratio_sq; R.dedup_xy = dedup_xy;
  for (int64_t p = 0; p < n_pairs; ++p) R.total_records += A.rows[pairs[2 * p + 1]];
  R.prune = ratio_sq <= 1.0f;
  R.rho = R.prune ? effective_rho(ctx, ratio_sq) : 1.0f;
  for (BatchSlot& S : ctx->slot) {
    S.busy = false;
    for (cudaEvent_t* e : {&S.ev_start, &S.ev_knn, &S.ev_done})
      if (!*e) CU_CHECK(ctx, cudaEventCreate(e));
    if (!S.ev_ready) CU_CHECK(ctx, cudaEventCreateWithFlags(&S.ev_ready, cudaEventDisableTiming));
    if (!S.ev_copied) CU_CHECK(ctx, cudaEventCreateWithFlags(&S.ev_copied, cudaEventDisableTiming));
  }
  int64_t p0 = 0;
  int k = 0;
  bool streaming = false;  // a streamed upload is still in flight
  for (size_t i = 0; i < A.ready.size() && !streaming; ++i) streaming = A.streamed[i] && !A.ready[i];
  if (n_pairs > 0) CU_CHECK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  while (p0 < n_pairs) {
    // batch = as many consecutive pairs as fit the record budget
    int64_t p1 = p0;
    long long rec = 0;
    // while images are still on the wire the first batches are small (1/16, 1/8, ... of the budget): the fused kernel
    // starts after the first handful of images instead of after everything the first full batch touches
    const long long budget = streaming && k < 4 ? (kBatchRecords >> (4 - k)) : kBatchRecords;
    while (p1 < n_pairs && (p1 - p0) < kBatchPairs) {
      const long long qr = A.rows[pairs[2 * p1 + 1]];
      if (p1 > p0 && rec + qr > budget) break;
      rec += qr;
      ++p1;
    }
    BatchSlot& S = ctx->slot[k & 1];
    rc = finish_batch(ctx, R, S);  // batch k-2 (normally long finished): its results leave the slot
    if (rc) return rc;
    rc = enqueue_batch(ctx, R, S, p0, p1);
    if (rc) return rc;
    // while batch k computes, batch k-1 is completed on the host side and its matches start travelling back
    rc = finish_batch(ctx, R, ctx->slot[(k & 1) ^ 1]);
    if (rc) return rc;
    p0 = p1;
    ++k;
  }
  // the batches still in flight, in order
  rc = finish_batch(ctx, R, ctx->slot[k & 1]);
  if (rc) return rc;
  rc = finish_batch(ctx, R, ctx->slot[(k & 1) ^ 1]);
  if (rc) return rc;
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->copy_stream));
  ctx->r_offsets.p[n_pairs] = R.match_base;
  if (out) {
    out->n_pairs = n_pairs;
    out->counts = ctx->r_counts.p;
    out->offsets = reinterpret_cast<const int64_t*>(ctx->r_offsets.p);
    out->matches = ctx->r_matches.p;
    out->gpu_ms = R.gpu_ms;
    out->knn_kernel_ms = R.knn_ms;
    out->knn_kernel_launches = R.knn_launches;
    out->total_launches = R.launches;
    out->rescanned_queries = ctx->rescanned;
  }
  return MVGCUDA_OK;
}

// Array level: the 2 nearest rows of db image `db_img` of arena A for every row of query image `q_img` of arena Q.
static int knn2_impl(mvgcuda_ctx* ctx, const Arena& A, int db_img, const Arena& Q, int q_img, int tie_mode, int32_t* idx, float* dist) {
  if (!idx || !dist) { ctx->set_error("null output"); return MVGCUDA_ERR_INVALID; }
  if (tie_mode != MVGCUDA_TIE_LOWEST_INDEX && tie_mode != MVGCUDA_TIE_REFERENCE) { ctx->set_error("bad tie_mode"); return MVGCUDA_ERR_INVALID; }
  if (db_img < 0 || db_img >= (int)A.rows.size() || q_img < 0 || q_img >= (int)Q.rows.size()) { ctx->set_error("image id out of range"); return MVGCUDA_ERR_INVALID; }
  const int nI = A.rows[db_img], nq = Q.rows[q_img];
  if (nI < 2 || nq < 1) {
    // matcher_brute_force.h:107-110
    ctx->set_error("Too much asked nearest neighbors");
    return MVGCUDA_ERR_INVALID;
  }
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  BatchSlot& S = ctx->slot[0];
  CU_CHECK(ctx, S.h_jobs.reserve(1));
  CU_CHECK(ctx, S.h_item_start.reserve(2));
  const int32_t pr[2] = {db_img, q_img};
  BatchPlan bp;
  plan_batch(S.h_jobs.p, S.h_item_start.p, A, Q, pr, 0, 1, bp);
  CU_CHECK(ctx, S.d_jobs.reserve(1));
  CU_CHECK(ctx, S.d_item_start.reserve(2));
  CU_CHECK(ctx, S.d_knn.reserve(nq));
  cudaStream_t st = ctx->stream;
  if (S.ev_copied) CU_CHECK(ctx, cudaStreamWaitEvent(st, S.ev_copied, 0));
  CU_CHECK(ctx, cudaMemcpyAsync(S.d_jobs.p, S.h_jobs.p, sizeof(PairJob), cudaMemcpyHostToDevice, st));
  CU_CHECK(ctx, cudaMemcpyAsync(S.d_item_start.p, S.h_item_start.p, 2 * sizeof(int), cudaMemcpyHostToDevice, st));
  KnnParams kp = {};
  kp.qcol = Q.ccol.p;
  kp.n_items = bp.n_items;
  kp.out = S.d_knn.p;
  kp.prune_ratio = FLT_MAX;  // exact 2-NN of every query
  kp.prune_rho = 1.0f;
  int rc = launch_knn_raw(ctx, Q.tmap_q, A, kp, S.d_jobs.p, S.d_item_start.p, 1, bp.n_items, ctx->d_items);
  if (rc) return rc;
  if (tie_mode == MVGCUDA_TIE_REFERENCE) {
    tie_fixup_kernel<<<dim3((nq + 7) / 8, 1), 256, 0, st>>>(A.desc.p, Q.desc.p, S.d_jobs.p, 0, S.d_knn.p);
    CU_CHECK(ctx, cudaGetLastError());
  }
  std::vector<KnnRecord> rec(nq);
  CU_CHECK(ctx, cudaMemcpyAsync(rec.data(), S.d_knn.p, nq * sizeof(KnnRecord), cudaMemcpyDeviceToHost, st));
  CU_CHECK(ctx, cudaStreamSynchronize(st));
  for (int q = 0; q < nq; ++q) {
    idx[2 * q] = rec[q].idx1;
    idx[2 * q + 1] = rec[q].idx2;
    dist[2 * q] = (float)rec[q].d1;  // exact: < 2^24
    dist[2 * q + 1] = (float)rec[q].d2;
  }
  return MVGCUDA_OK;
}

}  // namespace mvgcuda

// ================================================================================ C ABI
// Every entry point is a function-try-block: no C++ exception crosses the boundary (std::bad_alloc -> MVGCUDA_ERR_NOMEM).
#define MVG_GUARD(ctx_expr)                                                                                \
  catch (const std::bad_alloc&) {                                                                          \
    if (mvgcuda_ctx* c_ = (ctx_expr)) c_->set_error("out of host memory");                                 \
    return MVGCUDA_ERR_NOMEM;                                                                              \
  } catch (const std::exception& e_) {                                                                     \
    if (mvgcuda_ctx* c_ = (ctx_expr)) c_->set_error("unexpected exception: %s", e_.what());                \
    return MVGCUDA_ERR_INVALID;                                                                            \
  } catch (...) {                                                                                          \
    if (mvgcuda_ctx* c_ = (ctx_expr)) c_->set_error("unexpected exception");                               \
    return MVGCUDA_ERR_INVALID;                                                                            \
  }

namespace mvgcuda {
CtxView ctx_view(mvgcuda_ctx* ctx) {
  CtxView v;
  v.device = ctx->device;
  v.sm_count = ctx->prop.multiProcessorCount;
  v.stream = ctx->stream;
  v.feats = ctx->images.has_feats ? ctx->images.feat.p : nullptr;
  v.row0 = ctx->images.row0.data();
  v.rows = ctx->images.rows.data();
  v.n_images = (int)ctx->images.rows.size();
  v.geo = &ctx->geo;
  v.geo_free = &ctx->geo_free;
  return v;
}
void ctx_set_error(mvgcuda_ctx* ctx, const char* msg) { ctx->set_error("%s", msg); }
int ctx_wait_uploads(mvgcuda_ctx* ctx) { return wait_for_all_images(ctx, ctx->images); }
}  // namespace mvgcuda

extern "C" {

int mvgcuda_version(void) { return 200; }

int mvgcuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}

int mvgcuda_device_ordinal(int k) {
  int n = 0;
  if (k < 0 || cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return -1; }
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10 && k-- == 0) return d;
  }
  return -1;
}

const char* mvgcuda_last_error(const mvgcuda_ctx* ctx) { return ctx ? ctx->error : g_create_error; }

int mvgcuda_create(int device, mvgcuda_ctx** out) try {
  if (!out) { snprintf(g_create_error, sizeof g_create_error, "null out pointer"); return MVGCUDA_ERR_INVALID; }
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    snprintf(g_create_error, sizeof g_create_error, "no CUDA device: %s (libmvgcuda has no CPU fallback)",
             e != cudaSuccess ? cudaGetErrorString(e) : "count is 0");
    return MVGCUDA_ERR_CUDA;
  }
  if (device < 0 || device >= n) { snprintf(g_create_error, sizeof g_create_error, "device index out of range"); return MVGCUDA_ERR_INVALID; }
  mvgcuda_ctx* ctx = new (std::nothrow) mvgcuda_ctx();
  if (!ctx) { snprintf(g_create_error, sizeof g_create_error, "out of host memory"); return MVGCUDA_ERR_NOMEM; }
  ctx->device = device;
  auto fail = [&](const char* what, cudaError_t err) {
    snprintf(g_create_error, sizeof g_create_error, "%s: %s", what, cudaGetErrorString(err));
    delete ctx;
    return MVGCUDA_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
  if ((e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
  if (ctx->prop.major != 10) {
    snprintf(g_create_error, sizeof g_create_error, "device '%s' is sm_%d%d; libmvgcuda is built for sm_100a only", ctx->prop.name,
             ctx->prop.major, ctx->prop.minor);
    delete ctx;
    return MVGCUDA_ERR_CUDA;
  }
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->post_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  ctx->stream = ctx->own_stream;
  for (auto& ev : ctx->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail("cudaEventCreate", e);
  ctx->knn_smem = sizeof(KnnSmem) + 1024;
  // all of the SM's shared memory configured as shared: what knn2_kernel leaves (~30 KB) can then host the small CTAs of the
  // post stream beside it
  if ((e = cudaFuncSetAttribute(knn2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared)) != cudaSuccess)
    return fail("cudaFuncSetAttribute(carveout)", e);
  if ((e = cudaFuncSetAttribute(knn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->knn_smem)) != cudaSuccess)
    return fail("cudaFuncSetAttribute(knn2_kernel)", e);
  if ((e = cudaFuncSetAttribute(i8_peak_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(kBytesA + kBytesB + 1024))) != cudaSuccess)
    return fail("cudaFuncSetAttribute(probe)", e);
  *out = ctx;
  return MVGCUDA_OK;
}
MVG_GUARD(nullptr)

void mvgcuda_destroy(mvgcuda_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->geo && ctx->geo_free) ctx->geo_free(ctx->geo);
  ctx->images.release();
  ctx->scratch.release();
  for (BatchSlot& S : ctx->slot) S.release();
  ctx->d_items.release(); ctx->d_ritems.release(); ctx->d_tmp.release(); ctx->d_tmp2.release(); ctx->d_npass.release();
  ctx->d_counts_raw.release(); ctx->d_counts2.release(); ctx->d_offsets_raw.release(); ctx->d_total.release();
  ctx->d_nodes.release(); ctx->d_order.release();
  ctx->d_resc_ccol.release(); ctx->d_ritem_start.release(); ctx->d_used.release(); ctx->d_job_of.release();
  ctx->d_resc_desc.release(); ctx->d_resc_knn.release(); ctx->d_rjobs.release(); ctx->d_rsrc.release();
  ctx->h_ritem_start.release(); ctx->h_rjobs.release(); ctx->h_rsrc.release();
  ctx->r_counts.release(); ctx->r_offsets.release(); ctx->r_matches.release();
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->upload_stream) cudaStreamDestroy(ctx->upload_stream);
  if (ctx->post_stream) cudaStreamDestroy(ctx->post_stream);
  delete ctx;
}

int mvgcuda_set_tuning(mvgcuda_ctx* ctx, float prune_rho, int rescan_rows) {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  if (!(prune_rho > 0.0f) || prune_rho > 1.0f || rescan_rows < 0) { ctx->set_error("bad tuning values"); return MVGCUDA_ERR_INVALID; }
  ctx->prune_rho = prune_rho;
  ctx->rescan_cap_rows = rescan_rows ? rescan_rows : kDefaultRescanRows;
  return MVGCUDA_OK;
}

int mvgcuda_set_stream(mvgcuda_ctx* ctx, void* cuda_stream) {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return MVGCUDA_OK;
}

int mvgcuda_upload_images(mvgcuda_ctx* ctx, int n_images, const uint8_t* const* desc, const int32_t* rows, int pinned) try {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  (void)pinned;
  if (n_images < 0 || (n_images > 0 && (!desc || !rows))) { ctx->set_error("bad image list"); return MVGCUDA_ERR_INVALID; }
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  ctx->r_pairs = 0;
  return fill_arena(ctx, ctx->images, n_images, desc, rows);
}
MVG_GUARD(ctx)

int mvgcuda_stream_begin(mvgcuda_ctx* ctx, int n_images, const int32_t* rows) try {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  if (n_images < 0 || (n_images > 0 && !rows)) { ctx->set_error("bad image list"); return MVGCUDA_ERR_INVALID; }
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->upload_stream));  // a previous streamed set is complete before its arena is reused
  ctx->r_pairs = 0;
  int rc = layout_arena(ctx, ctx->images, n_images, rows);
  if (rc) return rc;
  Arena& A = ctx->images;
  CU_CHECK(ctx, A.feat.reserve((size_t)A.arena_rows));
  CU_CHECK(ctx, cudaMemsetAsync(A.feat.p, 0, (size_t)A.arena_rows * sizeof(float2), ctx->stream));
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // the upload stream starts from a zeroed arena
  A.has_feats = true;  // every stream_image call supplies them (or none is needed: match_pairs only)
  A.ready.assign(n_images, 0);
  for (int i = 0; i < n_images; ++i) if (rows[i] == 0) A.ready[i] = 1;  // nothing to wait for
  while ((int)A.img_ev.size() < n_images) {
    cudaEvent_t e = nullptr;
    CU_CHECK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    A.img_ev.push_back(e);
  }
  return MVGCUDA_OK;
}
MVG_GUARD(ctx)

int mvgcuda_stream_image(mvgcuda_ctx* ctx, int image, const uint8_t* desc, const float* feats_xy) try {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  Arena& A = ctx->images;
  if (image < 0 || image >= (int)A.rows.size() || image >= (int)A.img_ev.size()) { ctx->set_error("stream_image: image id out of range (call stream_begin first)"); return MVGCUDA_ERR_INVALID; }
  const int rows = A.rows[image];
  if (rows > 0 && !desc) { ctx->set_error("stream_image: null descriptor pointer"); return MVGCUDA_ERR_INVALID; }
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->upload_stream;  // copies and the constants kernel run beside the matching kernels
  if (rows > 0) {
    CU_CHECK(ctx, cudaMemcpyAsync(A.desc.p + (size_t)A.row0[image] * kDim, desc, (size_t)rows * kDim, cudaMemcpyDefault, st));
    if (feats_xy) { int rc = upload_features(ctx, A, image, feats_xy, rows, st); if (rc) return rc; }
  }
  int rc = launch_k1(ctx, A, A.row0[image], A.row0[image] + round_up(rows, kRowAlign), st);
  if (rc) return rc;
  CU_CHECK(ctx, cudaEventRecord(A.img_ev[image], st));
  A.streamed[image] = 1;
  A.ready[image] = 0;
  return MVGCUDA_OK;
}
MVG_GUARD(ctx)

int mvgcuda_stream_images(mvgcuda_ctx* ctx, int count, const int32_t* images, const uint8_t* const* desc, const float* const* feats_xy) try {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  if (count < 0 || (count > 0 && (!images || !desc))) { ctx->set_error("stream_images: bad list"); return MVGCUDA_ERR_INVALID; }
  for (int k = 0; k < count; ++k) {
    const int rc = mvgcuda_stream_image(ctx, images[k], desc[k], feats_xy ? feats_xy[k] : nullptr);
    if (rc) return rc;
  }
  return MVGCUDA_OK;
}
MVG_GUARD(ctx)

int mvgcuda_stream_end(mvgcuda_ctx* ctx) try {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->upload_stream));  // every staging buffer handed to stream_image is free again
  Arena& A = ctx->images;
  for (size_t i = 0; i < A.ready.size(); ++i) if (A.streamed[i]) A.ready[i] = 1;
  return MVGCUDA_OK;
}
MVG_GUARD(ctx)

int mvgcuda_host_alloc(size_t bytes, void** out) {
  if (!out) return MVGCUDA_ERR_INVALID;
  *out = nullptr;
  if (bytes == 0) return MVGCUDA_OK;
  const cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);  // portable: every GPU of the box copies from it asynchronously
  if (e != cudaSuccess) { cudaGetLastError(); *out = nullptr; return MVGCUDA_ERR_NOMEM; }
  return MVGCUDA_OK;
}
void mvgcuda_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int mvgcuda_num_images(const mvgcuda_ctx* ctx) { return ctx ? (int)ctx->images.rows.size() : -1; }
int mvgcuda_image_rows(const mvgcuda_ctx* ctx, int image) {
  if (!ctx || image < 0 || image >= (int)ctx->images.rows.size()) return -1;
  return ctx->images.rows[image];
}

int mvgcuda_knn2(mvgcuda_ctx* ctx, int db_img, int q_img, int tie_mode, int32_t* idx, float* dist) try {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  { int rc = wait_for_all_images(ctx, ctx->images); if (rc) return rc; }
  return knn2_impl(ctx, ctx->images, db_img, ctx->images, q_img, tie_mode, idx, dist);
}
MVG_GUARD(ctx)

int mvgcuda_knn2_arrays(mvgcuda_ctx* ctx, const uint8_t* db, int db_rows, const uint8_t* query, int q_rows,
                        int tie_mode, int32_t* idx, float* dist) try {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  if (db_rows < 2 || q_rows < 1) { ctx->set_error("Too much asked nearest neighbors"); return MVGCUDA_ERR_INVALID; }
  if (!db || !query) { ctx->set_error("null descriptor pointer"); return MVGCUDA_ERR_INVALID; }
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  const uint8_t* d[2] = {db, query};
  const int32_t r[2] = {db_rows, q_rows};
  int rc = fill_arena(ctx, ctx->scratch, 2, d, r);
  if (rc) return rc;
  return knn2_impl(ctx, ctx->scratch, 0, ctx->scratch, 1, tie_mode, idx, dist);
}
MVG_GUARD(ctx)

int mvgcuda_db_create(mvgcuda_ctx* ctx, const uint8_t* db, int rows, mvgcuda_db** out) try {
  if (!ctx || !out) return MVGCUDA_ERR_INVALID;
  *out = nullptr;
  if (rows < 1 || !db) { ctx->set_error("db_create: no rows"); return MVGCUDA_ERR_INVALID; }
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  mvgcuda_db* d = new mvgcuda_db();
  d->rows = rows;
  const uint8_t* dp[1] = {db};
  const int32_t r[1] = {rows};
  int rc = fill_arena(ctx, d->arena, 1, dp, r);
  if (rc) { d->arena.release(); delete d; return rc; }
  *out = d;
  return MVGCUDA_OK;
}
MVG_GUARD(ctx)

void mvgcuda_db_destroy(mvgcuda_ctx* ctx, mvgcuda_db* db) {
  if (!db) return;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  db->arena.release();
  delete db;
}

int mvgcuda_db_rows(const mvgcuda_db* db) { return db ? db->rows : -1; }

int mvgcuda_db_knn2(mvgcuda_ctx* ctx, const mvgcuda_db* db, const uint8_t* query, int q_rows, int tie_mode, int32_t* idx, float* dist) try {
  if (!ctx || !db) return MVGCUDA_ERR_INVALID;
  if (db->rows < 2 || q_rows < 1) { ctx->set_error("Too much asked nearest neighbors"); return MVGCUDA_ERR_INVALID; }
  if (!query) { ctx->set_error("null descriptor pointer"); return MVGCUDA_ERR_INVALID; }
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  const uint8_t* q[1] = {query};
  const int32_t r[1] = {q_rows};
  int rc = fill_arena(ctx, ctx->scratch, 1, q, r);  // only the queries travel
  if (rc) return rc;
  return knn2_impl(ctx, db->arena, 0, ctx->scratch, 0, tie_mode, idx, dist);
}
MVG_GUARD(ctx)

int mvgcuda_match_pairs(mvgcuda_ctx* ctx, int64_t n_pairs, const int32_t* pairs, float ratio_sq,
                        mvgcuda_pair_matches* out) try {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  return match_pairs_impl(ctx, n_pairs, pairs, ratio_sq, false, out);
}
MVG_GUARD(ctx)

int mvgcuda_clone_images(mvgcuda_ctx* ctx, const mvgcuda_ctx* src) try {
  if (!ctx || !src || ctx == src) return MVGCUDA_ERR_INVALID;
  const Arena& S = src->images;
  Arena& A = ctx->images;
  if (S.arena_rows <= 0 || !S.desc.p) { ctx->set_error("clone_images: the source context holds no images"); return MVGCUDA_ERR_INVALID; }
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  if (ctx->device != src->device) {
    int can = 0;
    CU_CHECK(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, src->device));
    if (can) {
      const cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);  // direct NVLink path; without it the copy is staged
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU_CHECK(ctx, e);
      (void)cudaGetLastError();
    }
  }
  ctx->r_pairs = 0;
  A.row0 = S.row0;
  A.rows = S.rows;
  A.arena_rows = S.arena_rows;
  A.has_feats = S.has_feats;
  A.streamed.assign(A.rows.size(), 0);
  A.ready.assign(A.rows.size(), 1);
  const int n_images = (int)A.rows.size();
  const size_t ccol_len = ccol_ints(A.arena_rows);
  CU_CHECK(ctx, A.desc.reserve((size_t)A.arena_rows * kDim));
  CU_CHECK(ctx, A.ccol.reserve(ccol_len));
  CU_CHECK(ctx, A.img_row0.reserve(n_images + 1));
  CU_CHECK(ctx, A.img_rows.reserve(std::max(n_images, 1)));
  cudaStream_t st = ctx->stream;
  CU_CHECK(ctx, cudaMemcpyPeerAsync(A.desc.p, ctx->device, S.desc.p, src->device, (size_t)A.arena_rows * kDim, st));
  CU_CHECK(ctx, cudaMemcpyPeerAsync(A.ccol.p, ctx->device, S.ccol.p, src->device, ccol_len * sizeof(int), st));
  if (S.has_feats) {
    CU_CHECK(ctx, A.feat.reserve((size_t)A.arena_rows));
    CU_CHECK(ctx, cudaMemcpyPeerAsync(A.feat.p, ctx->device, S.feat.p, src->device, (size_t)A.arena_rows * sizeof(float2), st));
  }
  CU_CHECK(ctx, cudaMemcpyAsync(A.img_row0.p, A.row0.data(), (n_images + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  if (n_images)
    CU_CHECK(ctx, cudaMemcpyAsync(A.img_rows.p, A.rows.data(), n_images * sizeof(int), cudaMemcpyHostToDevice, st));
  int rc = make_tmap(ctx, &A.tmap_q, A.desc.p, A.arena_rows, kBlockQ);
  if (rc) return rc;
  rc = make_tmap(ctx, &A.tmap_db, A.desc.p, A.arena_rows, 64);
  if (rc) return rc;
  CU_CHECK(ctx, cudaStreamSynchronize(st));
  return MVGCUDA_OK;
}
MVG_GUARD(ctx)

int mvgcuda_set_features(mvgcuda_ctx* ctx, int n_images, const float* const* feats_xy, const int32_t* rows) try {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  Arena& A = ctx->images;
  if (n_images != (int)A.rows.size()) { ctx->set_error("set_features: image count differs from uploaded set"); return MVGCUDA_ERR_INVALID; }
  if (n_images > 0 && (!feats_xy || !rows)) { ctx->set_error("set_features: null argument"); return MVGCUDA_ERR_INVALID; }
  for (int i = 0; i < n_images; ++i)
    if (rows[i] != A.rows[i] || (rows[i] > 0 && !feats_xy[i])) { ctx->set_error("set_features: image %d rows mismatch", i); return MVGCUDA_ERR_INVALID; }
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  CU_CHECK(ctx, A.feat.reserve((size_t)std::max(A.arena_rows, 1)));
  CU_CHECK(ctx, cudaMemsetAsync(A.feat.p, 0, (size_t)A.arena_rows * sizeof(float2), ctx->stream));
  for (int i = 0; i < n_images; ++i) {
    int rc = upload_features(ctx, A, i, feats_xy[i], rows[i], ctx->stream);
    if (rc) return rc;
  }
  CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // the caller's buffers are free again
  A.has_feats = true;
  return MVGCUDA_OK;
}
MVG_GUARD(ctx)

int mvgcuda_match_collection(mvgcuda_ctx* ctx, int64_t n_pairs, const int32_t* pairs, float ratio_sq,
                             int host_threads, mvgcuda_pair_matches* out) try {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  (void)host_threads;  // row 13 runs on the GPU since version 200; kept for ABI compatibility
  return match_pairs_impl(ctx, n_pairs, pairs, ratio_sq, true, out);
}
MVG_GUARD(ctx)

int mvgcuda_export_matches(mvgcuda_ctx* ctx, const int32_t* pairs, const char* path) try {
  if (!ctx || !pairs || !path) return MVGCUDA_ERR_INVALID;
  const int64_t n = ctx->r_pairs;
  const int* counts = ctx->r_counts.p;
  const long long* offs = ctx->r_offsets.p;
  const int* m = ctx->r_matches.p;
  // std::map<pair<size_t,size_t>,...> iteration order; map::insert keeps the FIRST of duplicate keys
  std::vector<int64_t> order(n);
  for (int64_t p = 0; p < n; ++p) order[p] = p;
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    if (pairs[2 * a] != pairs[2 * b]) return pairs[2 * a] < pairs[2 * b];
    return pairs[2 * a + 1] < pairs[2 * b + 1];
  });
  FILE* f = fopen(path, "wb");
  if (!f) { ctx->set_error("cannot open %s for writing", path); return MVGCUDA_ERR_IO; }
  // hand-rolled decimal formatting: at 499,500 pairs (BASELINE config 4) the text export is ~150 M lines and
  // snprintf would dominate the whole job
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
  for (int64_t k = 0; k < n && ok; ++k) {
    const int64_t p = order[k];
    if (k > 0 && pairs[2 * p] == pairs[2 * order[k - 1]] && pairs[2 * p + 1] == pairs[2 * order[k - 1] + 1]) continue;
    const int* mm = m + 2 * offs[p];
    int c = 0;
    // header + matches, flushing whenever fewer than 64 bytes of head-room remain
    put_int(pairs[2 * p], ' ');
    put_int(pairs[2 * p + 1], '\n');
    put_int(counts[p], '\n');
    while (c < counts[p]) {
      if (used + 64 > buf.size()) { ok = fwrite(buf.data(), 1, used, f) == used; used = 0; if (!ok) break; }
      put_int(mm[2 * c], ' ');
      put_int(mm[2 * c + 1], '\n');
      ++c;
    }
    if (used + 64 > buf.size()) { ok = ok && fwrite(buf.data(), 1, used, f) == used; used = 0; }
  }
  ok = ok && fwrite(buf.data(), 1, used, f) == used;
  if (fclose(f) != 0 || !ok) { ctx->set_error("short write to %s", path); return MVGCUDA_ERR_IO; }
  return MVGCUDA_OK;
}
MVG_GUARD(ctx)

int mvgcuda_get_device_info(const mvgcuda_ctx* ctx, mvgcuda_device_info* out) {
  if (!ctx || !out) return MVGCUDA_ERR_INVALID;
  memset(out, 0, sizeof *out);
  snprintf(out->name, sizeof out->name, "%s", ctx->prop.name);
  out->sm_count = ctx->prop.multiProcessorCount;
  out->cc_major = ctx->prop.major;
  out->cc_minor = ctx->prop.minor;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
  out->clock_khz = khz;
  out->hbm_bytes = (int64_t)ctx->prop.totalGlobalMem;
  return MVGCUDA_OK;
}

int mvgcuda_probe_i8_peak(mvgcuda_ctx* ctx, int iters, double* ops_per_sec, float* ms_out) try {
  if (!ctx || iters < 1) return MVGCUDA_ERR_INVALID;
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  const int grid = ctx->prop.multiProcessorCount;
  const size_t smem = kBytesA + kBytesB + 1024;
  cudaStream_t st = ctx->stream;
  i8_peak_probe_kernel<<<grid, 128, smem, st>>>(iters);  // warm-up
  CU_CHECK(ctx, cudaGetLastError());
  CU_CHECK(ctx, cudaEventRecord(ctx->ev[0], st));
  i8_peak_probe_kernel<<<grid, 128, smem, st>>>(iters);
  CU_CHECK(ctx, cudaGetLastError());
  CU_CHECK(ctx, cudaEventRecord(ctx->ev[1], st));
  CU_CHECK(ctx, cudaStreamSynchronize(st));
  float ms = 0.f;
  CU_CHECK(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  if (ms_out) *ms_out = ms;
  if (ops_per_sec) *ops_per_sec = 2.0 * kBlockQ * kTileDb * 32.0 * (double)iters * grid / (ms * 1e-3);
  return MVGCUDA_OK;
}
MVG_GUARD(ctx)

}  // extern "C"
