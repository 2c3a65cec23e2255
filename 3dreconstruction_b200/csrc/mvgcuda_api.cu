// mvgcuda_api.cu -- C ABI of libmvgcuda (include/mvgcuda.h): context, HBM arena, batch pipeline.
//
// Host-side structure (B200-first, not a translation of the reference's per-pair loop,
// matcher_all_in_memory.h:71-139): all descriptor arrays are resident in one HBM arena, the pair
// list is cut into batches, each batch is ONE persistent launch of the fused kernel over
// (pair, 128-query block) work items followed by three small compaction launches, and only the
// compacted matches travel back over PCIe.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cfloat>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <shared_mutex>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mvgcuda.h"
#include "kernels.cuh"

namespace mvgcuda {

static thread_local std::string g_create_error;

#define CU_CHECK(ctx, expr)                                                                         \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      (ctx)->set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MVGCUDA_ERR_CUDA;                                                                      \
    }                                                                                               \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template <typename T>
struct PinnedBuf {
  T* p = nullptr;
  size_t cap = 0;
  // grow, preserving the first `keep` elements
  cudaError_t reserve(size_t n, size_t keep = 0) {
    if (n <= cap) return cudaSuccess;
    size_t ncap = std::max(n, cap + cap / 2);
    T* np = nullptr;
    cudaError_t e = cudaMallocHost(&np, ncap * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep) memcpy(np, p, keep * sizeof(T));
    if (p) cudaFreeHost(p);
    p = np;
    cap = ncap;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

constexpr float kDefaultPruneRho = 0.8f;
constexpr int kDefaultRescanRows = 1 << 20;  // 128 MB of gathered queries per rescan round

struct Arena {
  DevBuf<uint8_t> desc;   // [rows_padded][128]
  DevBuf<int> ccol;       // K1 output
  DevBuf<int> img_row0, img_rows;
  std::vector<int> row0, rows;  // host copies
  int arena_rows = 0;
  CUtensorMap tmap_q;    // boxes of 128 rows: one CTA's query block
  CUtensorMap tmap_db;   // boxes of 64 rows: one CTA's share of an N = 128 MMA group
  void release() { desc.release(); ccol.release(); img_row0.release(); img_rows.release(); row0.clear(); rows.clear(); arena_rows = 0; }
};

}  // namespace mvgcuda

using namespace mvgcuda;

struct mvgcuda_ctx {
  int device = 0;
  cudaDeviceProp prop{};
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string error;
  size_t knn_smem = 0;

  Arena images;   // uploaded collection
  Arena scratch;  // knn2_arrays operands

  // batch buffers
  DevBuf<PairJob> d_jobs;
  DevBuf<int> d_item_start;
  DevBuf<KnnItem> d_items, d_ritems;  // work lists of K2 (batch / second pass)
  DevBuf<KnnRecord> d_knn;
  DevBuf<int2> d_tmp;
  DevBuf<int> d_npass, d_counts;
  DevBuf<long long> d_offsets;  // [batch_pairs + 1]
  DevBuf<long long> d_total;
  DevBuf<int2> d_matches;
  PinnedBuf<PairJob> h_jobs;
  PinnedBuf<int> h_item_start;

  // ratio-aware pruning: admission factor for failing queries and the rescan of ambiguous ones
  float prune_rho = kDefaultPruneRho;
  int rescan_cap_rows = kDefaultRescanRows;
  DevBuf<int> d_resc_idx, d_resc_cnt, d_resc_ccol, d_ritem_start;
  DevBuf<uint8_t> d_resc_desc;
  DevBuf<KnnRecord> d_resc_knn;
  DevBuf<PairJob> d_rjobs;
  DevBuf<RescanSrc> d_rsrc;
  PinnedBuf<int> h_resc_cnt, h_ritem_start;
  PinnedBuf<PairJob> h_rjobs;
  PinnedBuf<RescanSrc> h_rsrc;
  CUtensorMap tmap_resc;
  const uint8_t* tmap_resc_base = nullptr;
  int tmap_resc_rows = 0;
  long long rescanned = 0;  // queries matched a second time by the last match call
  PinnedBuf<long long> h_total;

  // results of the last match call
  PinnedBuf<int> r_counts;
  PinnedBuf<long long> r_offsets;
  PinnedBuf<int> r_matches;  // 2 ints per match
  int64_t r_pairs = 0;
  // results after host de-duplication (match_collection)
  std::vector<int> c_counts;
  std::vector<long long> c_offsets;
  std::vector<int> c_matches;
  bool last_was_collection = false;

  std::vector<std::vector<float>> feats;  // per image [rows][2]

  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};

  void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    error = buf;
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

// Lay the images out in the arena (each starting at a multiple of kRowAlign rows, zero padded),
// copy them in, run K1, build the tensor maps.
static int fill_arena(mvgcuda_ctx* ctx, Arena& A, int n_images, const uint8_t* const* desc, const int32_t* rows,
                      int pinned) {
  A.row0.assign(n_images + 1, 0);
  A.rows.assign(n_images, 0);
  long long total = 0;
  for (int i = 0; i < n_images; ++i) {
    if (rows[i] < 0 || (rows[i] > 0 && !desc[i])) { ctx->set_error("image %d: bad rows/pointer", i); return MVGCUDA_ERR_INVALID; }
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
  for (int i = 0; i < n_images; ++i) {
    if (rows[i] == 0) continue;
    // cudaMemcpyDefault: desc[i] may be host memory (pageable or pinned) or device memory (e.g. a replica that
    // arrived over NVLink) -- unified addressing sorts it out
    CU_CHECK(ctx, cudaMemcpyAsync(A.desc.p + (size_t)A.row0[i] * kDim, desc[i], (size_t)rows[i] * kDim,
                                  cudaMemcpyDefault, st));
  }
  (void)pinned;
  CU_CHECK(ctx, cudaMemcpyAsync(A.img_row0.p, A.row0.data(), (n_images + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  if (n_images)
    CU_CHECK(ctx, cudaMemcpyAsync(A.img_rows.p, A.rows.data(), n_images * sizeof(int), cudaMemcpyHostToDevice, st));
  row_consts_kernel<<<A.arena_rows / kK1Rows, 8 * kK1Rows, 0, st>>>(A.desc.p, A.img_row0.p, A.img_rows.p, n_images, A.arena_rows,
                                                       A.ccol.p);
  CU_CHECK(ctx, cudaGetLastError());
  int rc = make_tmap(ctx, &A.tmap_q, A.desc.p, A.arena_rows, kBlockQ);
  if (rc) return rc;
  rc = make_tmap(ctx, &A.tmap_db, A.desc.p, A.arena_rows, 64);
  if (rc) return rc;
  CU_CHECK(ctx, cudaStreamSynchronize(st));  // row0/rows vectors and caller buffers are free again
  return MVGCUDA_OK;
}

struct BatchPlan {
  int n_jobs = 0;
  int n_items = 0;
  long long n_records = 0;
  long long mean_db_rows = 0;  // item-weighted mean db rows of the batch
};

// Fill h_jobs / h_item_start for pairs [p0, p1) of the list; images from arena A.
static void plan_batch(mvgcuda_ctx* ctx, const Arena& A, const int32_t* pairs, int64_t p0, int64_t p1, BatchPlan& bp) {
  bp.n_jobs = (int)(p1 - p0);
  long long rec = 0, db_sum = 0;
  int items = 0;
  for (int64_t p = p0; p < p1; ++p) {
    const int I = pairs[2 * p], J = pairs[2 * p + 1];
    PairJob& j = ctx->h_jobs.p[p - p0];
    j.db_row0 = A.row0[I];
    j.db_rows = A.rows[I];
    j.q_row0 = A.row0[J];
    j.q_rows = A.rows[J];
    j.out_off = (int)rec;
    j.valid = (j.db_rows >= 2 && j.q_rows >= 1) ? 1 : 0;
    ctx->h_item_start.p[p - p0] = items;
    if (j.valid) {
      const int ni = (j.q_rows + 2 * kBlockQ - 1) / (2 * kBlockQ);  // a CTA pair takes two query blocks per item
      items += ni;
      db_sum += (long long)ni * j.db_rows;
    }
    rec += j.q_rows;
  }
  bp.mean_db_rows = items ? db_sum / items : 0;
  ctx->h_item_start.p[p1 - p0] = items;
  bp.n_items = items;
  bp.n_records = rec;
}

// Work list + one persistent launch of K2.  tmap_q: boxes of 128 query rows; A: the arena the db rows live in.
static int launch_knn_raw(mvgcuda_ctx* ctx, const CUtensorMap& tmap_q, const Arena& A, KnnParams kp, const PairJob* d_jobs,
                          const int* d_item_start, int n_jobs, DevBuf<KnnItem>& d_items) {
  if (kp.n_items == 0) return MVGCUDA_OK;
  CU_CHECK(ctx, d_items.reserve(kp.n_items));
  build_items_kernel<<<(kp.n_items + 255) / 256, 256, 0, ctx->stream>>>(d_jobs, d_item_start, n_jobs, kp.n_items, d_items.p);
  CU_CHECK(ctx, cudaGetLastError());
  kp.items = d_items.p;
  kp.ccol = A.ccol.p;
  kp.hmin = A.ccol.p + (size_t)(A.arena_rows / kTileDb) * kTileC;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kKnnThreads);
  cfg.gridDim = dim3(2 * std::min(kp.n_items, ctx->prop.multiProcessorCount / 2));  // CTA pairs
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

// prune_ratio: the Lowe ratio the records will be tested against, or FLT_MAX for the exact 2-NN of every query.
static int launch_knn(mvgcuda_ctx* ctx, const Arena& A, const BatchPlan& bp, float prune_ratio, float prune_rho) {
  KnnParams kp = {};
  kp.qcol = A.ccol.p;
  kp.n_items = bp.n_items;
  kp.out = ctx->d_knn.p;
  kp.prune_ratio = prune_ratio;
  kp.prune_rho = prune_rho;
  return launch_knn_raw(ctx, A.tmap_q, A, kp, ctx->d_jobs.p, ctx->d_item_start.p, bp.n_jobs, ctx->d_items);
}

constexpr long long kBatchRecords = 24ll << 20;  // queries per batch (384 MB of KnnRecord)
constexpr int kBatchPairs = 1 << 16;

static int reserve_batch(mvgcuda_ctx* ctx, int n_jobs, long long n_records) {
  CU_CHECK(ctx, ctx->d_jobs.reserve(n_jobs));
  CU_CHECK(ctx, ctx->d_item_start.reserve(n_jobs + 1));
  CU_CHECK(ctx, ctx->d_npass.reserve(n_jobs));
  CU_CHECK(ctx, ctx->d_counts.reserve(n_jobs));
  CU_CHECK(ctx, ctx->d_offsets.reserve(n_jobs + 1));
  CU_CHECK(ctx, ctx->d_total.reserve(1));
  CU_CHECK(ctx, ctx->d_knn.reserve(std::max<long long>(n_records, 1)));
  CU_CHECK(ctx, ctx->d_tmp.reserve(std::max<long long>(n_records, 1)));
  CU_CHECK(ctx, ctx->d_matches.reserve(std::max<long long>(n_records, 1)));
  return MVGCUDA_OK;
}

static int validate_pairs(mvgcuda_ctx* ctx, const Arena& A, int64_t n_pairs, const int32_t* pairs) {
  if (n_pairs < 0 || (n_pairs > 0 && !pairs)) { ctx->set_error("bad pair list"); return MVGCUDA_ERR_INVALID; }
  const int n = (int)A.rows.size();
  for (int64_t p = 0; p < n_pairs; ++p) {
    const int I = pairs[2 * p], J = pairs[2 * p + 1];
    if (I < 0 || I >= n || J < 0 || J >= n) { ctx->set_error("pair %lld: image id out of range", (long long)p); return MVGCUDA_ERR_INVALID; }
  }
  return MVGCUDA_OK;
}

// The admission factor actually used for a ratio: never below the ratio itself (a row with d < ratio * d(best) could be
// a passing nearest neighbour and must always be admitted), 1 = plain best-distance bound, no rescans.
static float effective_rho(const mvgcuda_ctx* ctx, float ratio_sq) {
  float rho = ctx->prune_rho;
  if (!(rho < 1.0f)) return 1.0f;
  return std::min(1.0f, std::max(rho, ratio_sq * 1.002f));
}

// After the pruned K2 pass of a batch of nb pairs: find the ambiguous queries, match them again exactly (same kernel,
// gathered query rows, prune_ratio = FLT_MAX) and put the exact records in place.  One small D2H + sync per batch; the
// gathered rows go through a bounded buffer in as many rounds as needed.
static int rescan_ambiguous(mvgcuda_ctx* ctx, const Arena& A, int nb, long long n_records, float ratio_sq, float rho,
                            int& launches) {
  cudaStream_t st = ctx->stream;
  CU_CHECK(ctx, ctx->d_resc_idx.reserve(std::max<long long>(n_records, 1)));
  CU_CHECK(ctx, ctx->d_resc_cnt.reserve(nb));
  CU_CHECK(ctx, ctx->h_resc_cnt.reserve(nb));
  flag_ambiguous_kernel<<<nb, kCompactThreads, 0, st>>>(ctx->d_jobs.p, ctx->d_knn.p, ratio_sq, rho, ctx->d_resc_idx.p,
                                                        ctx->d_resc_cnt.p);
  CU_CHECK(ctx, cudaGetLastError());
  ++launches;
  CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_resc_cnt.p, ctx->d_resc_cnt.p, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_CHECK(ctx, cudaStreamSynchronize(st));
  long long total = 0;
  for (int k = 0; k < nb; ++k) total += ctx->h_resc_cnt.p[k];
  if (total == 0) return MVGCUDA_OK;
  ctx->rescanned += total;
  const int cap = (int)std::min<long long>(std::max(ctx->rescan_cap_rows, 1), round_up((int)std::min<long long>(total, 1 << 30), 256));
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
  CU_CHECK(ctx, ctx->h_rjobs.reserve(nb + 1));
  CU_CHECK(ctx, ctx->h_rsrc.reserve(nb + 1));
  CU_CHECK(ctx, ctx->h_ritem_start.reserve(nb + 2));
  CU_CHECK(ctx, ctx->d_rjobs.reserve(nb + 1));
  CU_CHECK(ctx, ctx->d_rsrc.reserve(nb + 1));
  CU_CHECK(ctx, ctx->d_ritem_start.reserve(nb + 2));
  int j = 0, first = 0;
  while (j < nb) {
    // n_rs gather/scatter sources (one per original pair with flagged queries); n_rj second-pass jobs: consecutive
    // pairs against the same db image (the normal case, the pair list is (i, j)-ordered) share ONE job, so their few
    // flagged queries fill 128-query blocks together instead of one nearly empty block per pair
    int used = 0, n_rs = 0, n_rj = 0, items = 0;
    while (j < nb && used < cap) {
      const int left = ctx->h_resc_cnt.p[j] - first;
      if (left <= 0) { ++j; first = 0; continue; }
      const int take = std::min(left, cap - used);
      const PairJob& J = ctx->h_jobs.p[j];
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
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->d_ritem_start.p, ctx->h_ritem_start.p, (n_rj + 1) * sizeof(int),
                                  cudaMemcpyHostToDevice, st));
    rescan_gather_kernel<<<n_rs, 256, 0, st>>>(ctx->d_rsrc.p, ctx->d_resc_idx.p, A.desc.p, A.ccol.p, ctx->d_resc_desc.p,
                                               ctx->d_resc_ccol.p);
    CU_CHECK(ctx, cudaGetLastError());
    KnnParams kp = {};
    kp.qcol = ctx->d_resc_ccol.p;
    kp.n_items = items;
    kp.out = ctx->d_resc_knn.p;
    kp.prune_ratio = FLT_MAX;
    kp.prune_rho = 1.0f;
    int rc = launch_knn_raw(ctx, ctx->tmap_resc, A, kp, ctx->d_rjobs.p, ctx->d_ritem_start.p, n_rj, ctx->d_ritems);
    if (rc) return rc;
    rescan_scatter_kernel<<<n_rs, 256, 0, st>>>(ctx->d_rsrc.p, ctx->d_resc_idx.p, ctx->d_resc_knn.p, ctx->d_knn.p);
    CU_CHECK(ctx, cudaGetLastError());
    launches += 3;
    CU_CHECK(ctx, cudaStreamSynchronize(st));  // the pinned job lists are rewritten by the next round
  }
  return MVGCUDA_OK;
}

// Optional consumer of finished batches (collection level): `on_batch(p1)` is called as soon as the raw matches of
// pairs [0, p1) are on the host, so that the host-side coordinate de-dup overlaps the GPU work of the next batch.
// `results_mtx` is held exclusively while the pinned result buffer is re-allocated.
struct BatchSink {
  std::function<void(int64_t)> on_batch;
  std::shared_mutex* results_mtx = nullptr;
  long long batch_records = 0;  // smaller batches than the default, for a finer pipeline
};

static int match_pairs_impl(mvgcuda_ctx* ctx, int64_t n_pairs, const int32_t* pairs, float ratio_sq,
                            mvgcuda_pair_matches* out, BatchSink* sink = nullptr) {
  const Arena& A = ctx->images;
  int rc = validate_pairs(ctx, A, n_pairs, pairs);
  if (rc) return rc;
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  CU_CHECK(ctx, ctx->r_counts.reserve(std::max<int64_t>(n_pairs, 1)));
  CU_CHECK(ctx, ctx->r_offsets.reserve(n_pairs + 1));
  CU_CHECK(ctx, ctx->h_total.reserve(1));
  ctx->r_pairs = n_pairs;
  ctx->last_was_collection = false;
  float gpu_ms = 0.f, knn_ms = 0.f;
  int knn_launches = 0, launches = 0;
  ctx->rescanned = 0;
  long long match_base = 0;
  cudaStream_t st = ctx->stream;

  int64_t p0 = 0;
  while (p0 < n_pairs) {
    // batch = as many consecutive pairs as fit the record budget
    int64_t p1 = p0;
    long long rec = 0;
    while (p1 < n_pairs && (p1 - p0) < kBatchPairs) {
      const long long qr = A.rows[pairs[2 * p1 + 1]];
      if (p1 > p0 && rec + qr > (sink && sink->batch_records ? sink->batch_records : kBatchRecords)) break;
      rec += qr;
      ++p1;
    }
    const int nb = (int)(p1 - p0);
    CU_CHECK(ctx, ctx->h_jobs.reserve(nb));
    CU_CHECK(ctx, ctx->h_item_start.reserve(nb + 1));
    BatchPlan bp;
    plan_batch(ctx, A, pairs, p0, p1, bp);
    rc = reserve_batch(ctx, nb, bp.n_records);
    if (rc) return rc;
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->d_jobs.p, ctx->h_jobs.p, nb * sizeof(PairJob), cudaMemcpyHostToDevice, st));
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->d_item_start.p, ctx->h_item_start.p, (nb + 1) * sizeof(int),
                                  cudaMemcpyHostToDevice, st));
    CU_CHECK(ctx, cudaEventRecord(ctx->ev[0], st));
    // ratio <= 1: records only have to carry what the ratio test and the match list need (ratio-aware pruning);
    // ratio > 1 keeps the exact 2nd neighbour of every query for the tie fix-up below
    const bool prune = ratio_sq <= 1.0f;
    const float rho = prune ? effective_rho(ctx, ratio_sq) : 1.0f;
    rc = launch_knn(ctx, A, bp, prune ? ratio_sq : FLT_MAX, rho);
    if (rc) return rc;
    if (prune && rho < 1.0f && bp.n_items) {
      rc = rescan_ambiguous(ctx, A, nb, bp.n_records, ratio_sq, rho, launches);
      if (rc) return rc;
    }
    CU_CHECK(ctx, cudaEventRecord(ctx->ev[1], st));  // K2 incl. the exact second pass over ambiguous queries
    if (ratio_sq > 1.0f) {
      // For ratio > 1 a tie d1 == d2 passes the test, so WHICH of the tied rows is reported matters: reproduce the
      // reference's std::partial_sort choice exactly (one extra CUDA-core pass per pair; the usual ratios <= 1 skip it).
      for (int k = 0; k < nb; ++k) {
        const PairJob& J = ctx->h_jobs.p[k];
        if (!J.valid) continue;
        tie_fixup_kernel<<<(J.q_rows + 7) / 8, 256, 0, st>>>(A.desc.p, J, ctx->d_knn.p);
        ++launches;
      }
      CU_CHECK(ctx, cudaGetLastError());
    }
    ratio_filter_kernel<<<nb, kCompactThreads, 0, st>>>(ctx->d_jobs.p, ctx->d_knn.p, ratio_sq, ctx->d_tmp.p,
                                                        ctx->d_npass.p, ctx->d_counts.p);
    CU_CHECK(ctx, cudaGetLastError());
    scan_counts_kernel<<<1, 1024, 0, st>>>(ctx->d_counts.p, nb, match_base, ctx->d_offsets.p, ctx->d_total.p);
    CU_CHECK(ctx, cudaGetLastError());
    dedup_scatter_kernel<<<nb, kCompactThreads, 0, st>>>(ctx->d_jobs.p, ctx->d_tmp.p, ctx->d_npass.p, ctx->d_offsets.p,
                                                         match_base, ctx->d_matches.p);
    CU_CHECK(ctx, cudaGetLastError());
    CU_CHECK(ctx, cudaEventRecord(ctx->ev[2], st));
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_total.p, ctx->d_total.p, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->r_counts.p + p0, ctx->d_counts.p, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->r_offsets.p + p0, ctx->d_offsets.p, (nb + 1) * sizeof(long long),
                                  cudaMemcpyDeviceToHost, st));
    CU_CHECK(ctx, cudaStreamSynchronize(st));
    const long long new_total = ctx->h_total.p[0];
    const long long nm = new_total - match_base;
    {
      std::unique_lock<std::shared_mutex> lk;
      if (sink && sink->results_mtx && (size_t)std::max<long long>(new_total, 1) * 2 > ctx->r_matches.cap)
        lk = std::unique_lock<std::shared_mutex>(*sink->results_mtx);  // readers of the old buffer finish first
      CU_CHECK(ctx, ctx->r_matches.reserve((size_t)std::max<long long>(new_total, 1) * 2, (size_t)match_base * 2));
    }
    if (nm > 0) {
      CU_CHECK(ctx, cudaMemcpyAsync(ctx->r_matches.p + match_base * 2, ctx->d_matches.p, nm * sizeof(int2),
                                    cudaMemcpyDeviceToHost, st));
      CU_CHECK(ctx, cudaStreamSynchronize(st));
    }
    float ms = 0.f;
    CU_CHECK(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[2]));
    gpu_ms += ms;
    if (bp.n_items) {
      CU_CHECK(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
      knn_ms += ms;
      ++knn_launches;
      ++launches;
    }
    launches += 3;
    match_base = new_total;
    p0 = p1;
    if (sink && sink->on_batch) sink->on_batch(p1);
  }
  ctx->r_offsets.p[n_pairs] = match_base;
  if (out) {
    out->n_pairs = n_pairs;
    out->counts = ctx->r_counts.p;
    out->offsets = reinterpret_cast<const int64_t*>(ctx->r_offsets.p);
    out->matches = ctx->r_matches.p;
    out->gpu_ms = gpu_ms;
    out->knn_kernel_ms = knn_ms;
    out->knn_kernel_launches = knn_launches;
    out->total_launches = launches;
    out->rescanned_queries = ctx->rescanned;
  }
  return MVGCUDA_OK;
}

// One pair of arena A -> KnnRecords on the host.
static int knn2_impl(mvgcuda_ctx* ctx, const Arena& A, int db_img, int q_img, int tie_mode, int32_t* idx, float* dist) {
  if (!idx || !dist) { ctx->set_error("null output"); return MVGCUDA_ERR_INVALID; }
  if (tie_mode != MVGCUDA_TIE_LOWEST_INDEX && tie_mode != MVGCUDA_TIE_REFERENCE) { ctx->set_error("bad tie_mode"); return MVGCUDA_ERR_INVALID; }
  const int n = (int)A.rows.size();
  if (db_img < 0 || db_img >= n || q_img < 0 || q_img >= n) { ctx->set_error("image id out of range"); return MVGCUDA_ERR_INVALID; }
  const int nI = A.rows[db_img], nq = A.rows[q_img];
  if (nI < 2 || nq < 1) {
    // matcher_brute_force.h:107-110
    ctx->set_error("Too much asked nearest neighbors");
    return MVGCUDA_ERR_INVALID;
  }
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  CU_CHECK(ctx, ctx->h_jobs.reserve(1));
  CU_CHECK(ctx, ctx->h_item_start.reserve(2));
  const int32_t pr[2] = {db_img, q_img};
  BatchPlan bp;
  plan_batch(ctx, A, pr, 0, 1, bp);
  int rc = reserve_batch(ctx, 1, bp.n_records);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  CU_CHECK(ctx, cudaMemcpyAsync(ctx->d_jobs.p, ctx->h_jobs.p, sizeof(PairJob), cudaMemcpyHostToDevice, st));
  CU_CHECK(ctx, cudaMemcpyAsync(ctx->d_item_start.p, ctx->h_item_start.p, 2 * sizeof(int), cudaMemcpyHostToDevice, st));
  rc = launch_knn(ctx, A, bp, FLT_MAX, 1.0f);  // exact 2-NN of every query
  if (rc) return rc;
  if (tie_mode == MVGCUDA_TIE_REFERENCE) {
    const int warps_per_block = 8;
    tie_fixup_kernel<<<(nq + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(
        A.desc.p, ctx->h_jobs.p[0], ctx->d_knn.p);
    CU_CHECK(ctx, cudaGetLastError());
  }
  std::vector<KnnRecord> rec(nq);
  CU_CHECK(ctx, cudaMemcpyAsync(rec.data(), ctx->d_knn.p, nq * sizeof(KnnRecord), cudaMemcpyDeviceToHost, st));
  CU_CHECK(ctx, cudaStreamSynchronize(st));
  for (int q = 0; q < nq; ++q) {
    idx[2 * q] = rec[q].idx1;
    idx[2 * q + 1] = rec[q].idx2;
    dist[2 * q] = (float)rec[q].d1;  // exact: < 2^24
    dist[2 * q + 1] = (float)rec[q].d2;
  }
  return MVGCUDA_OK;
}

// ---- host side of the collection level: IndexedMatchDecorator<float>::getDeduplicated -----------------
// Same ordering predicate as indexed_match_decorator.h:33-53 and the same container (libstdc++
// std::set built with the iterator-range constructor, :93-95): the predicate is not a strict weak
// order, so the result is defined by the red-black tree's insertion behaviour and must not be
// "simplified".
struct DecoratedMatch {
  float x1, y1, x2, y2;
  int i, j;
};
static inline bool decorated_equal(const DecoratedMatch& a, const DecoratedMatch& b) {
  return a.x1 == b.x1 && a.y1 == b.y1 && a.x2 == b.x2 && a.y2 == b.y2;
}
struct DecoratedLess {
  bool operator()(const DecoratedMatch& a, const DecoratedMatch& b) const {
    if (decorated_equal(a, b)) return false;
    if (a.x1 < b.x1) return a.y1 < b.y1;
    if (a.x1 > b.x1) return a.y1 < b.y1;
    return a.x1 < b.x1;  // equal x1 (or unordered): false
  }
};

// Bump allocator for the set's nodes: the container, its comparator and its insertion sequence are exactly the
// reference's (so is the resulting tree); only where the nodes live changes -- one malloc per pair instead of one per
// match.  deallocate() is a no-op, the arena is rewound by the owner after the set is gone.
struct NodeArena {
  std::vector<char> buf;
  size_t used = 0;
  void* take(size_t bytes, size_t align) {
    size_t p = (used + align - 1) / align * align;
    if (p + bytes > buf.size()) return nullptr;
    used = p + bytes;
    return buf.data() + p;
  }
};
template <typename T>
struct ArenaAlloc {
  typedef T value_type;
  NodeArena* arena;
  explicit ArenaAlloc(NodeArena* a) : arena(a) {}
  template <typename U>
  ArenaAlloc(const ArenaAlloc<U>& o) : arena(o.arena) {}
  T* allocate(size_t n) {
    void* p = arena->take(n * sizeof(T), alignof(T));
    if (!p) throw std::bad_alloc();
    return static_cast<T*>(p);
  }
  void deallocate(T*, size_t) {}
  template <typename U> bool operator==(const ArenaAlloc<U>& o) const { return arena == o.arena; }
  template <typename U> bool operator!=(const ArenaAlloc<U>& o) const { return arena != o.arena; }
};

static void dedup_xy(const float* fI, const float* fJ, const int* m, int n, std::vector<int>& out, NodeArena& arena,
                     std::vector<DecoratedMatch>& v) {
  v.resize(n);
  for (int k = 0; k < n; ++k) {
    const int I = m[2 * k], J = m[2 * k + 1];
    v[k] = DecoratedMatch{fI[2 * I], fI[2 * I + 1], fJ[2 * J], fJ[2 * J + 1], I, J};
  }
  if (arena.buf.size() < (size_t)n * 96 + 256) arena.buf.resize((size_t)n * 96 + 256);  // rb-tree node = 32 B header + 24 B payload
  arena.used = 0;
  out.clear();
  {
    typedef std::set<DecoratedMatch, DecoratedLess, ArenaAlloc<DecoratedMatch> > Set;
    Set s(v.begin(), v.end(), DecoratedLess(), ArenaAlloc<DecoratedMatch>(&arena));  // same range construction as the reference (:93-95)
    out.reserve(2 * s.size());
    for (const DecoratedMatch& d : s) { out.push_back(d.i); out.push_back(d.j); }
  }
}

}  // namespace mvgcuda

// ================================================================================ C ABI
extern "C" {

int mvgcuda_version(void) { return 100; }

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

const char* mvgcuda_last_error(const mvgcuda_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int mvgcuda_create(int device, mvgcuda_ctx** out) {
  if (!out) { g_create_error = "null out pointer"; return MVGCUDA_ERR_INVALID; }
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count is 0") +
                     " (libmvgcuda has no CPU fallback)";
    return MVGCUDA_ERR_CUDA;
  }
  if (device < 0 || device >= n) { g_create_error = "device index out of range"; return MVGCUDA_ERR_INVALID; }
  mvgcuda_ctx* ctx = new (std::nothrow) mvgcuda_ctx();
  if (!ctx) { g_create_error = "out of host memory"; return MVGCUDA_ERR_NOMEM; }
  ctx->device = device;
  auto fail = [&](const char* what, cudaError_t err) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
    delete ctx;
    return MVGCUDA_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
  if ((e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
  if (ctx->prop.major != 10) {
    g_create_error = std::string("device '") + ctx->prop.name + "' is sm_" + std::to_string(ctx->prop.major) +
                     std::to_string(ctx->prop.minor) + "; libmvgcuda is built for sm_100a only";
    delete ctx;
    return MVGCUDA_ERR_CUDA;
  }
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  ctx->stream = ctx->own_stream;
  for (auto& ev : ctx->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail("cudaEventCreate", e);
  ctx->knn_smem = sizeof(KnnSmem) + 1024;
  if ((e = cudaFuncSetAttribute(knn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->knn_smem)) != cudaSuccess)
    return fail("cudaFuncSetAttribute(knn2_kernel)", e);
  if ((e = cudaFuncSetAttribute(i8_peak_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(kBytesA + kBytesB + 1024))) != cudaSuccess)
    return fail("cudaFuncSetAttribute(probe)", e);
  *out = ctx;
  return MVGCUDA_OK;
}

void mvgcuda_destroy(mvgcuda_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ctx->images.release();
  ctx->scratch.release();
  ctx->d_jobs.release(); ctx->d_item_start.release(); ctx->d_items.release(); ctx->d_ritems.release(); ctx->d_knn.release(); ctx->d_tmp.release();
  ctx->d_npass.release(); ctx->d_counts.release(); ctx->d_offsets.release(); ctx->d_total.release();
  ctx->d_matches.release();
  ctx->h_jobs.release(); ctx->h_item_start.release(); ctx->h_total.release();
  ctx->d_resc_idx.release(); ctx->d_resc_cnt.release(); ctx->d_resc_ccol.release(); ctx->d_ritem_start.release();
  ctx->d_resc_desc.release(); ctx->d_resc_knn.release(); ctx->d_rjobs.release(); ctx->d_rsrc.release();
  ctx->h_resc_cnt.release(); ctx->h_ritem_start.release(); ctx->h_rjobs.release(); ctx->h_rsrc.release();
  ctx->r_counts.release(); ctx->r_offsets.release(); ctx->r_matches.release();
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
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

int mvgcuda_upload_images(mvgcuda_ctx* ctx, int n_images, const uint8_t* const* desc, const int32_t* rows, int pinned) {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  if (n_images < 0 || (n_images > 0 && (!desc || !rows))) { ctx->set_error("bad image list"); return MVGCUDA_ERR_INVALID; }
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  ctx->feats.clear();
  ctx->r_pairs = 0;
  return fill_arena(ctx, ctx->images, n_images, desc, rows, pinned);
}

int mvgcuda_num_images(const mvgcuda_ctx* ctx) { return ctx ? (int)ctx->images.rows.size() : -1; }
int mvgcuda_image_rows(const mvgcuda_ctx* ctx, int image) {
  if (!ctx || image < 0 || image >= (int)ctx->images.rows.size()) return -1;
  return ctx->images.rows[image];
}

int mvgcuda_knn2(mvgcuda_ctx* ctx, int db_img, int q_img, int tie_mode, int32_t* idx, float* dist) {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  return knn2_impl(ctx, ctx->images, db_img, q_img, tie_mode, idx, dist);
}

int mvgcuda_knn2_arrays(mvgcuda_ctx* ctx, const uint8_t* db, int db_rows, const uint8_t* query, int q_rows,
                        int tie_mode, int32_t* idx, float* dist) {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  if (db_rows < 2 || q_rows < 1) { ctx->set_error("Too much asked nearest neighbors"); return MVGCUDA_ERR_INVALID; }
  if (!db || !query) { ctx->set_error("null descriptor pointer"); return MVGCUDA_ERR_INVALID; }
  CU_CHECK(ctx, cudaSetDevice(ctx->device));
  const uint8_t* d[2] = {db, query};
  const int32_t r[2] = {db_rows, q_rows};
  int rc = fill_arena(ctx, ctx->scratch, 2, d, r, 0);
  if (rc) return rc;
  return knn2_impl(ctx, ctx->scratch, 0, 1, tie_mode, idx, dist);
}

int mvgcuda_match_pairs(mvgcuda_ctx* ctx, int64_t n_pairs, const int32_t* pairs, float ratio_sq,
                        mvgcuda_pair_matches* out) {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  return match_pairs_impl(ctx, n_pairs, pairs, ratio_sq, out);
}

int mvgcuda_clone_images(mvgcuda_ctx* ctx, const mvgcuda_ctx* src) {
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
  ctx->feats = src->feats;
  ctx->r_pairs = 0;
  A.row0 = S.row0;
  A.rows = S.rows;
  A.arena_rows = S.arena_rows;
  const int n_images = (int)A.rows.size();
  const size_t ccol_len = ccol_ints(A.arena_rows);
  CU_CHECK(ctx, A.desc.reserve((size_t)A.arena_rows * kDim));
  CU_CHECK(ctx, A.ccol.reserve(ccol_len));
  CU_CHECK(ctx, A.img_row0.reserve(n_images + 1));
  CU_CHECK(ctx, A.img_rows.reserve(std::max(n_images, 1)));
  cudaStream_t st = ctx->stream;
  CU_CHECK(ctx, cudaMemcpyPeerAsync(A.desc.p, ctx->device, S.desc.p, src->device, (size_t)A.arena_rows * kDim, st));
  CU_CHECK(ctx, cudaMemcpyPeerAsync(A.ccol.p, ctx->device, S.ccol.p, src->device, ccol_len * sizeof(int), st));
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

int mvgcuda_set_features(mvgcuda_ctx* ctx, int n_images, const float* const* feats_xy, const int32_t* rows) {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  if (n_images != (int)ctx->images.rows.size()) { ctx->set_error("set_features: image count differs from uploaded set"); return MVGCUDA_ERR_INVALID; }
  for (int i = 0; i < n_images; ++i)
    if (rows[i] != ctx->images.rows[i] || (rows[i] > 0 && !feats_xy[i])) { ctx->set_error("set_features: image %d rows mismatch", i); return MVGCUDA_ERR_INVALID; }
  ctx->feats.resize(n_images);
  for (int i = 0; i < n_images; ++i) ctx->feats[i].assign(feats_xy[i], feats_xy[i] + 2 * (size_t)rows[i]);
  return MVGCUDA_OK;
}

int mvgcuda_match_collection(mvgcuda_ctx* ctx, int64_t n_pairs, const int32_t* pairs, float ratio_sq,
                             int host_threads, mvgcuda_pair_matches* out) {
  if (!ctx) return MVGCUDA_ERR_INVALID;
  if (ctx->feats.size() != ctx->images.rows.size()) { ctx->set_error("match_collection: call mvgcuda_set_features first"); return MVGCUDA_ERR_INVALID; }
  if (host_threads <= 0) host_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  host_threads = (int)std::min<int64_t>(host_threads, std::max<int64_t>(n_pairs, 1));
  // Host pool: pair p is de-duplicated as soon as its batch has landed, while the GPU works on the next batch.
  std::vector<std::vector<int>> per_pair(n_pairs);
  std::atomic<int64_t> cursor{0}, avail{0};
  std::atomic<bool> failed{false};
  std::mutex cv_m;
  std::condition_variable cv;
  std::shared_mutex results_mtx;
  auto work = [&]() {
    std::vector<int> tmp;
    NodeArena arena;
    std::vector<DecoratedMatch> deco;
    for (;;) {
      const int64_t p = cursor.fetch_add(1);
      if (p >= n_pairs) break;
      if (avail.load(std::memory_order_acquire) <= p) {
        std::unique_lock<std::mutex> lk(cv_m);
        cv.wait(lk, [&] { return avail.load(std::memory_order_acquire) > p || failed.load(); });
      }
      if (failed.load()) break;
      const int I = pairs[2 * p], J = pairs[2 * p + 1];
      {
        std::shared_lock<std::shared_mutex> lk(results_mtx);
        dedup_xy(ctx->feats[I].data(), ctx->feats[J].data(), ctx->r_matches.p + 2 * ctx->r_offsets.p[p], ctx->r_counts.p[p], tmp,
                 arena, deco);
      }
      per_pair[p] = tmp;
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < host_threads; ++t) th.emplace_back(work);
  BatchSink sink;
  sink.results_mtx = &results_mtx;
  sink.batch_records = 6ll << 20;
  sink.on_batch = [&](int64_t p1) {
    { std::lock_guard<std::mutex> lk(cv_m); avail.store(p1, std::memory_order_release); }
    cv.notify_all();
  };
  mvgcuda_pair_matches raw;
  int rc = match_pairs_impl(ctx, n_pairs, pairs, ratio_sq, &raw, &sink);
  if (rc) {
    { std::lock_guard<std::mutex> lk(cv_m); failed.store(true); }
    cv.notify_all();
  }
  for (auto& t : th) t.join();
  if (rc) return rc;
  ctx->c_counts.resize(n_pairs);
  ctx->c_offsets.resize(n_pairs + 1);
  long long tot = 0;
  for (int64_t p = 0; p < n_pairs; ++p) {
    ctx->c_counts[p] = (int)per_pair[p].size() / 2;
    ctx->c_offsets[p] = tot;
    tot += ctx->c_counts[p];
  }
  ctx->c_offsets[n_pairs] = tot;
  ctx->c_matches.resize((size_t)tot * 2);
  for (int64_t p = 0; p < n_pairs; ++p)
    if (!per_pair[p].empty())
      memcpy(&ctx->c_matches[2 * ctx->c_offsets[p]], per_pair[p].data(), per_pair[p].size() * sizeof(int));
  ctx->last_was_collection = true;
  if (out) {
    *out = raw;
    out->counts = ctx->c_counts.data();
    out->offsets = reinterpret_cast<const int64_t*>(ctx->c_offsets.data());
    out->matches = ctx->c_matches.data();
  }
  return MVGCUDA_OK;
}

int mvgcuda_export_matches(mvgcuda_ctx* ctx, const int32_t* pairs, const char* path) {
  if (!ctx || !pairs || !path) return MVGCUDA_ERR_INVALID;
  const int64_t n = ctx->r_pairs;
  const int* counts = ctx->last_was_collection ? ctx->c_counts.data() : ctx->r_counts.p;
  const long long* offs = ctx->last_was_collection ? ctx->c_offsets.data() : ctx->r_offsets.p;
  const int* m = ctx->last_was_collection ? ctx->c_matches.data() : ctx->r_matches.p;
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

int mvgcuda_probe_i8_peak(mvgcuda_ctx* ctx, int iters, double* ops_per_sec, float* ms_out) {
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

#if 0
// developer probe builds only (not part of include/mvgcuda.h)
int mvgcuda_debug_counters(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_dbg, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_dbg, z, sizeof z); }
  return 0;
}
#endif

}  // extern "C"
