// kernels.cuh -- the sm_100a kernels of libmvgcuda.
//
//   K1 row_consts_kernel   per arena row:  ccol = (||d||^2 << 8) | (row & 255), plus min ||d||^2 per 16 and per 128 rows
//   K2 knn2_kernel         fused u8xu8->s32 tcgen05 GEMM + ||d||^2 - 2 q.d + running top-2
//                          (replaces matcher_brute_force.h:117-131 + metric.h:57-81 +
//                          indexed_sort.h:52-66; the distance matrix lives only in TMEM/registers)
//   K3 ratio/compaction    DistanceRatioFilter (matching_filters.h:27-47), drop-last loop
//                          (matcher_all_in_memory.h:117-122), unique-on-_i (indexed_match.h:49-55)
//
// All arithmetic on the path is exact int32; the only fp32 operation is the ratio test
// float(d1) < ratio_sq * float(d2) (one __fmul_rn, strict <), as in the reference.
#pragma once
#include <cstddef>

#include "ptx.cuh"
#include "rbtree_dedup.cuh"

namespace mvgcuda {

constexpr int kDim = 128;      // descriptor bytes == GEMM K
constexpr int kBlockQ = 128;   // query rows per block   (MMA M per CTA, one TMEM lane per query)
constexpr int kTileDb = 256;   // db rows per tile       (one TMEM column per db row)
constexpr int kHalfCols = kTileDb / 2;  // columns of one MMA group == one accumulator buffer
constexpr int kSliceCols = 64; // columns one epilogue warp takes of every tile
constexpr int kSlotsA = 2;     // query block double buffer
constexpr int kStagesB = 8;    // db tile ring (TMA -> MMA); a CTA stages half a tile (16 KB) per stage
constexpr int kSlotsC = 8;     // per-column constant ring (TMA -> epilogue): one slot per tile PAIR, outlives the B stages
constexpr int kAccBufs = 4;    // TMEM accumulator buffers of 128 columns (all 512 columns)
constexpr int kRowAlign = 256; // every image starts at a multiple of this in the arena
constexpr int kPadNorm = 0x7FFFFF;  // "norm" of padding rows: > 128*255^2, so they never win

constexpr uint32_t kBytesA = kBlockQ * kDim;        // 16 KB
constexpr uint32_t kBytesB = kTileDb * kDim;        // 32 KB of db rows per tile, half of it in each CTA of a pair
constexpr uint32_t kStageBytes = kBytesB / 2;       // per CTA
constexpr int kChunk = 16;                     // db rows per TMEM load / exact step in the epilogue
constexpr int kTileC = kTileDb + kTileDb / kChunk;  // per-tile constants: 256 packed (norm<<8|col) + 16 chunk-min norms
constexpr uint32_t kBytesC = kTileC * sizeof(int);  // 1088 B

constexpr int kEpiParts = 4;     // warps per TMEM lane quadrant; they share 32 queries, 64 columns of every tile each
constexpr int kNumEpiWarps = 4 * kEpiParts;
constexpr int kFirstEpiWarp = 2;                       // warp 0: TMA producer + TMEM allocator, warp 1: MMA issuer of group 0
constexpr int kMmaWarp1 = kFirstEpiWarp + kNumEpiWarps;  // warp 18: MMA issuer of group 1
constexpr int kKnnThreads = 32 * (kMmaWarp1 + 1);      // 608 threads; three schedulers host 5 warps -> 96 registers per thread

struct PairJob {
  int db_row0;  // arena row of image I (db), multiple of kRowAlign
  int db_rows;
  int q_row0;   // arena row of image J (query)
  int q_rows;
  int out_off;  // first record of this pair in the knn output buffer
  int valid;    // db_rows >= 2 && q_rows >= 1
};

struct KnnRecord {  // one per query
  int idx1, idx2;   // db rows of nearest / second nearest
  int d1, d2;       // exact squared distances
};

// One work item of K2 = (image pair, two consecutive blocks of 128 queries): one CTA pair (cluster of 2, one TPC).  The
// CTA of cluster rank r owns the queries [128 r, 128 r + 128) of the item.
struct alignas(16) KnnItem {
  int db_row0, db_rows;  // db image: first arena row (multiple of 256) and row count
  int q_row0;            // first query row of the item (row of the QUERY tensor map)
  int q_valid;           // valid query rows from there (>= 1; only the first 256 belong to this item)
  int out_off;           // record index of the item's first query
  int pad0, pad1, pad2;
};

struct RescanMeta {  // written by plan_rescan_kernel (second pass planned on the device)
  int n_items;   // work items of the second knn2_kernel launch
  int total;     // ambiguous queries of the batch
  int overflow;  // total > cap: nothing was planned, the host re-runs the batch through its bounded multi-round path
  int n_jobs;
};

struct KnnSmem {
  alignas(1024) uint8_t a[kSlotsA][kBytesA];
  alignas(1024) uint8_t b[kStagesB][kStageBytes];
  alignas(16) int c[kSlotsC][2 * kTileC];
  uint64_t a_full[kSlotsA], a_empty[kSlotsA];
  uint64_t b_full[kStagesB], b_empty[kStagesB];
  uint64_t c_full[kSlotsC], c_empty[kSlotsC];  // (the epilogue relies on c_empty following c_full)
  uint64_t acc_full[kAccBufs], acc_empty[kAccBufs];
  uint32_t tmem_base;
  int bound[2][kBlockQ];  // per (item parity, query): best-known admission bound t, atomically tightened by all parts
  alignas(16) int4 xchg[2][kEpiParts - 1][kBlockQ];  // parts 1.. hand their top-2 to part 0 at the end of an item
};

struct KnnParams {
  const int* __restrict__ ccol;        // K1 output, [arena_rows / 256][kTileC]: per-row constants of the db operand
  const int* __restrict__ hmin;        // tail of K1's output: min ||d||^2 of every 128 db rows
  const int* __restrict__ qcol;        // K1 layout for the rows the QUERY tensor map addresses (== ccol, except rescans)
  const KnnItem* __restrict__ items;   // [n_items]
  int n_items;
  const RescanMeta* __restrict__ meta; // non-null: the item count lives on the device (meta->n_items), n_items is ignored
  KnnRecord* __restrict__ out;
  // Ratio-aware pruning (see slice_commit): the fp32 squared ratio of the Lowe test the records feed, or FLT_MAX when the
  // caller needs the exact 2nd neighbour of EVERY query (array-level API, ratio > 1 with the tie fix-up).
  float prune_ratio;
  float prune_rho;  // in (0, 1]: a failing query admits only rows with d <= rho * d(best); 1 = no rescans ever needed
};

// ------------------------------------------------------------------------------------------ K1
// Per-row constants, laid out per 256-row tile as [256 x ((||d||^2 << 8) | col)] [16 x min ||d||^2 of each 16-row chunk],
// followed (after the last tile) by one min ||d||^2 per 128 rows -- the filter constant of one epilogue warp's slice of a
// tile, which the warp keeps in a register (the host presets those to 0x7F7F7F7F; atomicMin here).
// 8 threads per 128-byte row (one 16-B load each), __dp4a squares, 3 shuffles; ONE chunk of 16 rows per block of 128
// threads: small enough (128 threads x 40 registers, no shared memory to speak of) to run next to a resident knn2_kernel
// CTA, so the constants of an image that arrives while earlier pairs are being matched do not wait for that kernel.
__device__ __forceinline__ int ccol_index(int row) { return (row >> 8) * kTileC + (row & 255); }
__host__ __device__ constexpr size_t ccol_ints(int arena_rows) {
  return static_cast<size_t>(arena_rows / kTileDb) * kTileC + static_cast<size_t>(arena_rows / kHalfCols) + 2;  // + 2: the epilogue prefetches one tile past an item's last
}
constexpr int kK1Rows = kChunk;  // rows per block of K1

__global__ void __launch_bounds__(8 * kK1Rows)
row_consts_kernel(const uint8_t* __restrict__ arena, const int* __restrict__ img_row0, const int* __restrict__ img_rows,
                  int n_images, int arena_rows, int row_begin /* multiple of 256: first arena row of this launch */,
                  int* __restrict__ ccol) {
  __shared__ int norms[kK1Rows];
  const int block_row0 = row_begin + blockIdx.x * kK1Rows;
  const int row = block_row0 + (threadIdx.x >> 3);
  const int part = threadIdx.x & 7;
  const uint4 v = *reinterpret_cast<const uint4*>(arena + (size_t)row * kDim + part * 16);
  unsigned s = 0;
  s = __dp4a(v.x, v.x, s);
  s = __dp4a(v.y, v.y, s);
  s = __dp4a(v.z, v.z, s);
  s = __dp4a(v.w, v.w, s);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (part == 0) {
    // image that owns this arena row: last i with img_row0[i] <= row
    int lo = 0, hi = n_images - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (img_row0[mid] <= row) lo = mid; else hi = mid - 1;
    }
    const bool real = n_images > 0 && (row - img_row0[lo]) < img_rows[lo];
    const int norm = real ? static_cast<int>(s) : kPadNorm;
    norms[threadIdx.x >> 3] = norm;
    ccol[ccol_index(row)] = (norm << 8) | (row & 255);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = norms[0];
#pragma unroll
    for (int k = 1; k < kK1Rows; ++k) m = min(m, norms[k]);
    ccol[(block_row0 >> 8) * kTileC + kTileDb + ((block_row0 & 255) >> 4)] = m;
    atomicMin(&ccol[static_cast<size_t>(arena_rows / kTileDb) * kTileC + block_row0 / kHalfCols], m);
  }
}

// ------------------------------------------------------------------------------------------ K2
// Work list: item -> (db image, 256 query rows).  One thread per item (binary search over the per-job item prefix), so the
// three roles of K2 read ONE 32-byte descriptor per item instead of each running the search (16 dependent L2 loads,
// longer than the whole item for images of ~2,000 rows).
__global__ void __launch_bounds__(256)
build_items_kernel(const PairJob* __restrict__ jobs, const int* __restrict__ item_start, int n_jobs, int n_items,
                   const RescanMeta* __restrict__ meta, KnnItem* __restrict__ items) {
  if (meta) { n_jobs = meta->n_jobs; n_items = meta->n_items; }  // counts that only the device knows
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_items) return;
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {  // last job with item_start[job] <= item
    const int mid = (lo + hi + 1) >> 1;
    if (item_start[mid] <= item) lo = mid; else hi = mid - 1;
  }
  const PairJob J = jobs[lo];
  const int q0 = 2 * kBlockQ * (item - item_start[lo]);
  KnnItem it;
  it.db_row0 = J.db_row0;
  it.db_rows = J.db_rows;
  it.q_row0 = J.q_row0 + q0;
  it.q_valid = J.q_rows - q0;
  it.out_off = J.out_off + q0;
  it.pad0 = it.pad1 = it.pad2 = 0;
  items[item] = it;
}

// Sorted pair (lo <= hi) helpers for the exact top-2 of a group of rows: a merge tree has plenty of instruction-level
// parallelism, where serial insertions form one long dependency chain.
struct Pair2 { int lo, hi; };
__device__ __forceinline__ Pair2 sort2(int a, int b) { return Pair2{min(a, b), max(a, b)}; }
__device__ __forceinline__ Pair2 merge2(Pair2 a, Pair2 b) {
  return Pair2{min(a.lo, b.lo), __vimin3_s32(max(a.lo, b.lo), a.hi, b.hi)};
}

// max of 16 in 8 three-input ops
__device__ __forceinline__ int max16(const int32_t* __restrict__ x) {
  const int a = __vimax3_s32(__vimax3_s32(x[0], x[1], x[2]), __vimax3_s32(x[3], x[4], x[5]), __vimax3_s32(x[6], x[7], x[8]));
  const int b = __vimax3_s32(__vimax3_s32(x[9], x[10], x[11]), __vimax3_s32(x[12], x[13], x[14]), x[15]);
  return max(a, b);
}

constexpr int kTInit = 0x3FFFFFFF;  // "no bound yet": hm - kTInit cannot overflow and passes every row

// (t, index) lexicographic order: smaller distance first, lower db row on ties
__device__ __forceinline__ bool cand_less(int ta, int ia, int tb, int ib) { return ta < tb || (ta == tb && ia < ib); }

// Exact packed keys ((||d||^2 - 2 q.d) << 8 | column) of four rows (x = raw dot products, cc = their packed constants
// (||d||^2 << 8 | column)), merged into the tile's top-2 (l1 <= l2).  Ties are decided by the packed compare: lowest column.
__device__ __forceinline__ void top2_of4(const int32_t* __restrict__ x, const int4 cc, int& l1, int& l2) {
  const int p0 = static_cast<int>(static_cast<uint32_t>(cc.x) - 512u * static_cast<uint32_t>(x[0]));
  const int p1 = static_cast<int>(static_cast<uint32_t>(cc.y) - 512u * static_cast<uint32_t>(x[1]));
  const int p2 = static_cast<int>(static_cast<uint32_t>(cc.z) - 512u * static_cast<uint32_t>(x[2]));
  const int p3 = static_cast<int>(static_cast<uint32_t>(cc.w) - 512u * static_cast<uint32_t>(x[3]));
  const Pair2 c = merge2(sort2(p0, p1), sort2(p2, p3));
  const int nl2 = __vimin3_s32(max(l1, c.lo), l2, c.hi);
  l1 = min(l1, c.lo);
  l2 = nl2;
}
// all 16 rows of a chunk (constants: warp-uniform shared-memory address, i.e. a broadcast load)
__device__ __forceinline__ void exact_chunk(const int32_t* __restrict__ x, const uint32_t cs_saddr, int& l1, int& l2) {
#pragma unroll
  for (int k = 0; k < 4; ++k) top2_of4(x + 4 * k, ptx::lds128(cs_saddr + 16 * k), l1, l2);
}

// End of a tile in which some chunk was examined exactly: merge the tile's top-2 (packed keys l1 <= l2) into the running
// top-2 -- ties keep the earlier (lower) index; branch-free selects; an untouched l1/l2 (0x7FFFFFFF) decodes to
// t = 0x7FFFFF, larger than any real t -- and tighten the admission bound T (admit t <= T):
//   * g1t <= g2t are the two smallest t = ||d||^2 - 2 q.d of rows this warp has really seen; rows with t > g2t can never
//     enter the top-2;
//   * ratio-aware pruning: when (g1t, g2t) already FAILS the Lowe test (d1 >= ratio * d2, the reference's fp32
//     expression), rows with g1t < t <= g2t cannot matter either -- such a row could only become the final 2nd neighbour,
//     and then d2 <= g2t makes the query fail the test whether it is recorded or not; if a better 1st neighbour turns up
//     later, the 2nd neighbour is at most g1t.  So the bound drops to g1t: idx1/d1 stay exact for every query, d2/idx2 stay
//     exact for every query that passes, a failing query keeps a d2 that still fails (DESIGN.md section 4 has the proof);
//   * admission factor rho < 1: a failing query admits only d <= rho * d(g1), rounded up.  It may then even miss its
//     true nearest row (one within rho..1 of the recorded one cannot pass the test either); what stays exact is every
//     pass/fail decision and the nearest row of every passing query -- after the second pass over the records that
//     flag_ambiguous_kernel cannot decide (tests/test_pruning_model.py checks the rule).
__device__ __forceinline__ void slice_commit(const int l1, const int l2, const int t, const int qn, const float prune_ratio,
                                             const float prune_rho, const uint32_t bound_saddr, int& g1t, int& g1i, int& g2t,
                                             int& g2i, int& T) {
  const int base = t * kTileDb;
  const int t1 = l1 >> 8, i1 = base + (l1 & 255);
  const int t2 = l2 >> 8, i2 = base + (l2 & 255);
  const bool a = t1 < g1t, b = t2 < g1t, c = t1 < g2t;
  const int n2t = a ? (b ? t2 : g1t) : (c ? t1 : g2t);
  const int n2i = a ? (b ? i2 : g1i) : (c ? i1 : g2i);
  g1t = a ? t1 : g1t;
  g1i = a ? i1 : g1i;
  g2t = n2t;
  g2i = n2i;
  const bool passes = __int2float_rn(qn + g1t) < __fmul_rn(prune_ratio, __int2float_rn(qn + g2t));
  const int tf = __float2int_ru(__fmul_ru(prune_rho, __int2float_rn(qn + g1t))) - qn;
  T = min(T, passes ? g2t : tf);
  ptx::red_min_shared(bound_saddr, T);  // the warps scanning the other columns of these queries tighten their filter now
}

// Values the compiler must keep in a register instead of re-deriving them (it otherwise rematerialises the aligned base
// of the dynamic shared memory -- a dozen uniform-datapath instructions -- in front of every barrier access of a loop).
__device__ __forceinline__ uint32_t keep_reg(uint32_t x) {
  asm volatile("" : "+r"(x));
  return x;
}

// K2: persistent, warp-specialised, one CTA PAIR (cluster of 2 = one TPC) per work item; grid = 2 x (#SMs / 2).
//
//   pipeline   One tcgen05.mma.cta_group::2 spans both SMs (M = 256: 128 queries from each CTA's shared memory); each CTA
//              stages only HALF of every db tile (B is split along N), which halves the L2 -> shared traffic per SM and
//              keeps the shared-memory read rate of the tensor pipe at 96 B/clk for N = 128 MMAs (a single CTA needs
//              128 B/clk for that shape and starves, profiles/r02a).  A tile of 256 db rows is issued as two N = 128
//              groups (4 K-steps of 32 bytes each) into FOUR accumulator buffers of 128 TMEM columns, each with its own
//              full/empty barrier pair.
//   warp 0     TMA producer (both CTAs): query block (2 slots) per item, per tile two boxes of 64 db rows (8-stage ring)
//              and, per tile pair, the per-column constants (2 x 1088 B bulk copy, 8-slot ring).  All loads of a pair
//              signal the LEADER's barriers.
//   warps 1,18 MMA issuers (leader CTA only), one per MMA group of a tile.  Commits are multicast to the barriers of both CTAs.
//   warps 2-17 epilogue, 4 per TMEM lane quadrant (= per scheduler).  Warp (quad, g, sub) owns lanes 32 quad.. and, of EVERY
//              tile, the 64 columns [128 g + 64 sub, +64) = 64 registers, loaded at once; the accumulator buffer goes
//              back to the MMA warp as soon as those loads have landed -- before any arithmetic -- so a buffer is away
//              from the tensor pipe for one TMEM-load latency only.
//   filter     A row can only matter if ||d||^2 - 2 q.d <= T, hence only if 2 q.d >= hm - T with hm = min ||d||^2 of the
//              tile half (one prefetched register): 35 max ops over the 64 raw dot products, ONE compare, one vote, one
//              branch per tile -- no shared memory, no per-chunk control flow.  Only when some lane passes: which chunks
//              (four votes), exact packed keys of all 16 rows of those chunks straight from the registers, top-2 merge,
//              one bound update (slice_commit).  The first tile of an item skips the filter (no useful bound yet).
//   bookkeeping Every role keeps ring positions / barrier phases as counters that step (no div/mod per tile), shared-memory
//              addresses as 32-bit registers, and every item starts in accumulator buffer 0; the uniform-datapath
//              instructions ptxas would otherwise re-derive per tile were the limit of the MMA issue rate.
__global__ void __launch_bounds__(kKnnThreads, 1)
knn2_kernel(const __grid_constant__ CUtensorMap tmap_q,   // box 128 rows x 128 B (query rows)
            const __grid_constant__ CUtensorMap tmap_db,  // box 64 rows x 128 B (db rows)
            const KnnParams p) {
  extern __shared__ uint8_t smem_raw[];
  KnnSmem& s = *reinterpret_cast<KnnSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // warp / lane through a shuffle: values ptxas cannot rematerialise.  Derived directly from %tid they are re-read with
  // S2R (~25 clocks, and a dependent shift) inside the tile loop instead of being kept in a register.
  const int lane = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x & 31), static_cast<int>(threadIdx.x & 31));
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const uint32_t rank = ptx::cluster_ctarank();  // 0 = leader: issues the MMAs of the pair
  const int worker = static_cast<int>(blockIdx.x >> 1);
  const int n_workers = static_cast<int>(gridDim.x >> 1);
  const int n_items = p.meta ? p.meta->n_items : p.n_items;
  const uint32_t sbase = ptx::smem_u32(&s);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_db);
  }
  if (warp == 1 && lane == 0) {
    // a query slot / db stage is free again when BOTH MMA issuers' instructions have read it (one commit each)
    for (int i = 0; i < kSlotsA; ++i) { ptx::mbar_init(&s.a_full[i], 1); ptx::mbar_init(&s.a_empty[i], 2); }
    for (int i = 0; i < kStagesB; ++i) { ptx::mbar_init(&s.b_full[i], 1); ptx::mbar_init(&s.b_empty[i], 2); }
    // a tile pair's constants are used by all 16 epilogue warps of this CTA
    for (int i = 0; i < kSlotsC; ++i) { ptx::mbar_init(&s.c_full[i], 1); ptx::mbar_init(&s.c_empty[i], kNumEpiWarps); }
    // an accumulator buffer goes back to the leader's MMA warp when the 8 warps of its column half in BOTH CTAs have read it
    for (int i = 0; i < kAccBufs; ++i) { ptx::mbar_init(&s.acc_full[i], 1); ptx::mbar_init(&s.acc_empty[i], 2 * kNumEpiWarps / 2); }
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc_pair<512>(&s.tmem_base);
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    const uint32_t sb_ = keep_reg(sbase);
    const uint32_t a_full_l = keep_reg(ptx::mapa_shared(sb_ + offsetof(KnnSmem, a_full), 0));  // the LEADER's barriers
    const uint32_t b_full_l = keep_reg(ptx::mapa_shared(sb_ + offsetof(KnnSmem, b_full), 0));
    uint32_t a_it = 0;
    uint32_t sb = 0, b_par = 1;  // B ring position; parity to wait for on b_empty (a fresh barrier passes a wait on parity 1)
    uint32_t sc = 0, c_par = 1;  // constants ring, same
    for (int item = worker; item < n_items; item += n_workers, ++a_it) {
      const int4 it = *reinterpret_cast<const int4*>(&p.items[item]);  // db_row0, db_rows, q_row0, q_valid
      const uint32_t sa = a_it & 1u;
      ptx::mbar_wait_addr(sb_ + offsetof(KnnSmem, a_empty) + 8u * sa, ((a_it >> 1) & 1u) ^ 1u);
      if (ptx::elect_one()) {
        // both CTAs load their own query block; the bytes of both are counted on the leader's barrier
        if (rank == 0) ptx::mbar_arrive_expect_tx_addr(sb_ + offsetof(KnnSmem, a_full) + 8u * sa, 2 * kBytesA);
        ptx::tma_load_2d_pair_addr(sb_ + offsetof(KnnSmem, a) + sa * kBytesA, &tmap_q, 0, it.z + kBlockQ * static_cast<int>(rank),
                                   a_full_l + 8u * sa);
      }
      __syncwarp();
      const int ntiles = (it.y + kTileDb - 1) / kTileDb;
      const int* ctile = p.ccol + static_cast<size_t>(it.x >> 8) * kTileC;
      int row0 = it.x + 64 * static_cast<int>(rank);
      for (int t = 0; t < ntiles; ++t) {
        if ((t & 1) == 0) {  // one ring slot per tile PAIR (the constants of consecutive tiles are contiguous)
          ptx::mbar_wait_addr(sb_ + offsetof(KnnSmem, c_empty) + 8u * sc, c_par);
          if (ptx::elect_one()) {
            const uint32_t bytes = (t + 1 < ntiles ? 2u : 1u) * kBytesC;
            ptx::mbar_arrive_expect_tx_addr(sb_ + offsetof(KnnSmem, c_full) + 8u * sc, bytes);
            ptx::bulk_load_1d_addr(sb_ + offsetof(KnnSmem, c) + sc * (2u * kBytesC), ctile, bytes, sb_ + offsetof(KnnSmem, c_full) + 8u * sc);
          }
          __syncwarp();
          ctile += 2 * kTileC;
          if (++sc == kSlotsC) { sc = 0; c_par ^= 1u; }
        }
        ptx::mbar_wait_addr(sb_ + offsetof(KnnSmem, b_empty) + 8u * sb, b_par);
        if (ptx::elect_one()) {
          // this CTA stages HALF of the tile: MMA group h covers tile rows [128 h, 128 h + 128) and the CTA of rank r
          // supplies its columns [64 r, 64 r + 64): two boxes of 64 rows
          if (rank == 0) ptx::mbar_arrive_expect_tx_addr(sb_ + offsetof(KnnSmem, b_full) + 8u * sb, kBytesB);
          const uint32_t dst = sb_ + offsetof(KnnSmem, b) + sb * kStageBytes;
          ptx::tma_load_2d_pair_addr(dst, &tmap_db, 0, row0, b_full_l + 8u * sb);
          ptx::tma_load_2d_pair_addr(dst + 64 * kDim, &tmap_db, 0, row0 + kHalfCols, b_full_l + 8u * sb);
        }
        __syncwarp();
        row0 += kTileDb;
        if (++sb == kStagesB) { sb = 0; b_par ^= 1u; }
      }
    }
    // Do not leave (and let the CTA's shared memory go) while commits of the leader can still arrive here: the last
    // a_empty commit of the pair follows every other commit, wait for it.
    if (a_it > 0) ptx::mbar_wait_addr(sb_ + offsetof(KnnSmem, a_empty) + 8u * ((a_it - 1) & 1u), ((a_it - 1) >> 1) & 1u);
  } else if (warp == 1 || warp == kMmaWarp1) {
    // ===================== MMA issuers (leader CTA): ONE elected lane of each runs the whole loop =====================
    // Two warps on two schedulers: warp 1 issues MMA group 0 of every tile (tile rows [0, 128) -> accumulator buffers 0, 2),
    // warp 18 group 1 (rows [128, 256) -> buffers 1, 3).  One issuer was the limit of the pipeline: ~100 instructions per tile,
    // each a dependent step of a single thread that shares its scheduler with four epilogue warps (profiles/r02j).
    // (elect.sync rather than `lane == 0`: the compiler then knows a single thread is active and feeds the uniform-datapath
    // operands of UTCIMMA / UTCBAR with one R2UR each instead of a loop over lanes.)
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_u8u8s32(2 * kBlockQ, kHalfCols);
      const uint32_t h = warp == 1 ? 0u : 1u;
      const uint32_t sb_ = keep_reg(sbase);
      const uint32_t a_lo0 = keep_reg(ptx::kmajor_sw128_desc_lo(sb_ + offsetof(KnnSmem, a)));
      const uint32_t b_lo0 = keep_reg(ptx::kmajor_sw128_desc_lo(sb_ + offsetof(KnnSmem, b)) + h * (64 * kDim / 16));  // group h reads this CTA's rows [64 h, +64) of a stage
      uint32_t a_it = 0;
      uint32_t sb = 0, b_par = 0;  // B ring position; parity to wait for on b_full
      uint32_t eph = 0x3u;         // bit (t & 1): parity to wait for on acc_empty of buffer 2 (t & 1) + h (fresh barriers pass parity 1)
      for (int item = worker; item < n_items; item += n_workers, ++a_it) {
        const int db_rows = p.items[item].db_rows;
        const uint32_t sa = a_it & 1u;
        ptx::mbar_wait_addr(sb_ + offsetof(KnnSmem, a_full) + 8u * sa, (a_it >> 1) & 1u);
        const uint32_t a_lo = a_lo0 + sa * (kBytesA >> 4);
        const int ntiles = (db_rows + kTileDb - 1) / kTileDb;
        for (int t = 0; t < ntiles; ++t) {
          const uint32_t tb = static_cast<uint32_t>(t & 1);  // every item starts in buffers 0 / 1
          const uint32_t buf = 2u * tb + h;
          ptx::mbar_wait_addr(sb_ + offsetof(KnnSmem, b_full) + 8u * sb, b_par);
          ptx::mbar_wait_addr(sb_ + offsetof(KnnSmem, acc_empty) + 8u * buf, (eph >> tb) & 1u);
          eph ^= 1u << tb;
          ptx::tc_fence_after();
          const uint32_t b_lo = b_lo0 + sb * (kStageBytes >> 4);
#pragma unroll
          for (int k = 0; k < kDim / 32; ++k)  // K = 32 bytes per kind::i8 instruction
            ptx::mma_i8_pair_lo(tmem_base + buf * kHalfCols, a_lo + 2 * k, b_lo + 2 * k, idesc, k > 0);
          ptx::mma_commit_pair_addr(sb_ + offsetof(KnnSmem, acc_full) + 8u * buf);  // accumulator ready in both CTAs
          ptx::mma_commit_pair_addr(sb_ + offsetof(KnnSmem, b_empty) + 8u * sb);    // this issuer is done with the db stage (both CTAs)
          if (t == ntiles - 1) ptx::mma_commit_pair_addr(sb_ + offsetof(KnnSmem, a_empty) + 8u * sa);  // ... and with the query slots
          if (++sb == kStagesB) { sb = 0; b_par ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;  // a warp may only touch its own TMEM lane quadrant
    const int part = (warp - kFirstEpiWarp) >> 2;
    const int g = part >> 1;    // column half of the tile == MMA group == accumulator buffers g and 2 + g
    const int sub = part & 1;   // 64-column slice of that half
    const int row = quad * 32 + lane;
    const uint32_t tslice = keep_reg(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + g * kHalfCols + sub * kSliceCols);
    // loop-invariant shared-memory addresses (32-bit, shared space)
    const uint32_t full0 = keep_reg(sbase + offsetof(KnnSmem, acc_full) + 8u * g);
    const uint32_t full1 = keep_reg(sbase + offsetof(KnnSmem, acc_full) + 8u * (2 + g));
    const uint32_t empty0 = keep_reg(ptx::mapa_shared(sbase + offsetof(KnnSmem, acc_empty) + 8u * g, 0));        // the leader's
    const uint32_t empty1 = keep_reg(ptx::mapa_shared(sbase + offsetof(KnnSmem, acc_empty) + 8u * (2 + g), 0));
    const uint32_t c_ring = keep_reg(sbase + offsetof(KnnSmem, c) + static_cast<uint32_t>(g * kHalfCols + sub * kSliceCols) * 4u);  // the slice's constants in slot 0
    const uint32_t c_bars = keep_reg(sbase + offsetof(KnnSmem, c_full));
    const uint32_t bound0 = keep_reg(sbase + offsetof(KnnSmem, bound) + 4u * row);
    uint32_t ph = 0;              // bit b: parity of the next acc_full phase of buffer b (of this column half)
    uint32_t sc = 0, c_par = 0;   // constants ring position; parity to wait for on c_full
    uint32_t item_it = 0;
    ptx::sts32(bound0, kTInit);
    ptx::sts32(bound0 + 4u * kBlockQ, kTInit);
    asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * kEpiParts) : "memory");
    for (int item = worker; item < n_items; item += n_workers, ++item_it) {
      const int4 it = *reinterpret_cast<const int4*>(&p.items[item]);  // db_row0, db_rows, q_row0, q_valid
      const int out_off = p.items[item].out_off;
      const int q_local = kBlockQ * static_cast<int>(rank) + row;
      const bool q_ok = q_local < it.w;
      const int ntiles = (it.y + kTileDb - 1) / kTileDb;
      const int qn = p.qcol[ccol_index(it.z + min(q_local, it.w - 1))] >> 8;  // ||q||^2 (dist = qn + t)
      int g1t = 0x7FFFFFFF, g2t = 0x7FFFFFFF, g1i = -1, g2i = -1;  // running best two (t-domain) of this thread's columns
      int T = kTInit;                                              // admit t <= T
      const uint32_t bound_saddr = keep_reg(bound0 + (item_it & 1u) * (4u * kBlockQ));
      ptx::sts32(bound0 + ((item_it & 1u) ^ 1u) * (4u * kBlockQ), kTInit);  // idle slot (every part left the previous item at the barrier below)
      int hm_off = (it.x >> 7) + g;  // min ||d||^2 of tile t's half g: p.hmin[hm_off + 2 t]
      int hm_next = __ldg(p.hmin + hm_off);

      // One tile in accumulator buffer (kBuf, g); kBuf is a compile-time constant: everything it indexes is a loop invariant.
#define MVG_TILE_STEP(kBuf, t_expr)                                                                                       \
      {                                                                                                                    \
        const int t = (t_expr);                                                                                            \
        ptx::mbar_wait_addr(kBuf ? full1 : full0, (ph >> kBuf) & 1u);                                                      \
        ph ^= (1u << kBuf);                                                                                                \
        ptx::tc_fence_after();                                                                                             \
        int32_t va[64]; /* the whole 64-column slice in one load */                                                         \
        ptx::tmem_ld_32x32b_x64(tslice + kBuf * kTileDb, va);                                                              \
        const int32_t* const v0 = va; const int32_t* const v1 = va + 16;                                                   \
        const int32_t* const v2 = va + 32; const int32_t* const v3 = va + 48;                                              \
        const int hm = hm_next;                                                                                            \
        hm_off += 2;                                                                                                       \
        hm_next = __ldg(p.hmin + hm_off); /* unconditional: the array is padded by one tile */                             \
        T = min(T, ptx::lds32_volatile(bound_saddr));                                                                      \
        ptx::tmem_ld_wait_for64(va);                                                                                       \
        ptx::tc_fence_before();                                                                                            \
        __syncwarp();                                                                                                      \
        if (lane == 0) ptx::mbar_arrive_cluster(kBuf ? empty1 : empty0); /* the buffer goes back to the MMA warp now */    \
        const int m0 = max16(v0), m1 = max16(v1), m2 = max16(v2), m3 = max16(v3);                                          \
        const int thr = (hm - T + 1) >> 1; /* 2 m >= hm - T  <=>  m >= ceil((hm - T) / 2) */                               \
        if (__any_sync(0xffffffffu, max(__vimax3_s32(m0, m1, m2), m3) >= thr) || t == 0) {                                 \
          ptx::mbar_wait_addr(c_bars + 8u * sc, c_par); /* per-column constants: needed on this path only */               \
          const uint32_t cs = c_ring + sc * (2u * kBytesC) + kBuf * kBytesC;                                               \
          int l1 = 0x7FFFFFFF, l2 = 0x7FFFFFFF;                                                                            \
          const bool first = t == 0;                                                                                       \
          if (first || __any_sync(0xffffffffu, m0 >= thr)) exact_chunk(v0, cs, l1, l2);                                    \
          if (first || __any_sync(0xffffffffu, m1 >= thr)) exact_chunk(v1, cs + 64, l1, l2);                               \
          if (first || __any_sync(0xffffffffu, m2 >= thr)) exact_chunk(v2, cs + 128, l1, l2);                              \
          if (first || __any_sync(0xffffffffu, m3 >= thr)) exact_chunk(v3, cs + 192, l1, l2);                              \
          slice_commit(l1, l2, t, qn, p.prune_ratio, p.prune_rho, bound_saddr, g1t, g1i, g2t, g2i, T);                     \
        }                                                                                                                  \
      }
      for (int tp = 0; tp < ntiles; tp += 2) {
        MVG_TILE_STEP(0, tp)
        if (tp + 1 < ntiles) MVG_TILE_STEP(1, tp + 1)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_addr(c_bars + 8u * (kSlotsC + sc));  // c_empty follows c_full in KnnSmem
        if (++sc == kSlotsC) { sc = 0; c_par ^= 1u; }
      }
#undef MVG_TILE_STEP
      // parts 1.. hand their result to part 0, which merges by (t, row) and writes the record
      if (part > 0) s.xchg[item_it & 1][part - 1][row] = make_int4(g1t, g1i, g2t, g2i);
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * kEpiParts) : "memory");  // the warps sharing these 32 queries
      if (part == 0 && q_ok) {
        int b1t = g1t, b1i = g1i, b2t = g2t, b2i = g2i;
#pragma unroll
        for (int o_ = 0; o_ < kEpiParts - 1; ++o_) {
          const int4 o = s.xchg[item_it & 1][o_][row];
          if (cand_less(o.x, o.y, b1t, b1i)) {
            if (cand_less(o.z, o.w, b1t, b1i)) { b2t = o.z; b2i = o.w; } else { b2t = b1t; b2i = b1i; }
            b1t = o.x; b1i = o.y;
          } else if (cand_less(o.x, o.y, b2t, b2i)) {
            b2t = o.x; b2i = o.y;
          }
        }
        KnnRecord r;
        r.idx1 = b1i; r.idx2 = b2i; r.d1 = qn + b1t; r.d2 = qn + b2t;
        *reinterpret_cast<int4*>(&p.out[out_off + q_local]) = *reinterpret_cast<const int4*>(&r);
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();  // no CTA leaves while its peer may still signal it
  if (warp == 0) ptx::tmem_dealloc_pair<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------ probe
// Tensor-pipe ceiling: back-to-back kind::i8 M128xN256xK32 MMAs on whatever is in shared memory.
__global__ void __launch_bounds__(128, 1) i8_peak_probe_kernel(int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (int)(kBytesA + kBytesB) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(base)[i] = 0x01010101u * (i & 3);
  if (threadIdx.x == 0) { ptx::mbar_init(&done, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc<512>(&tmem_slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 32) {
    constexpr uint32_t idesc = ptx::make_idesc_u8u8s32(kBlockQ, kTileDb);
    const uint64_t adesc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(base));
    const uint64_t bdesc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(base + kBytesA));
    for (int i = 0; i < iters; ++i)
      ptx::mma_i8_ss(tmem_base + (i & 1) * kTileDb, adesc + 2 * (i & 3), bdesc + 2 * (i & 3), idesc, 1);
    ptx::mma_commit(&done);
    ptx::mbar_wait(&done, 0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------ tie fix-up
// Array-level API only (MVGCUDA_TIE_REFERENCE).  The reference's raw 2-NN indices under ties are
// what libstdc++'s std::partial_sort(first, first+2, last) leaves behind (indexed_sort.h:52-66),
// which equals this two-slot machine run over the db rows in index order:
//     (T,S) = d[1] < d[0] ? (0,1) : (1,0)
//     for v = 2..n-1:  if d[v] < d[T]:  if d[S] < d[v]: T = v   else: T = S, S = v
//     result = [S, T]
// Rows with d > D2 (the exact 2nd-smallest value, known from K2) can only occupy a slot
// transiently and never change which small rows end up in S and T, so the machine is run over
// rows {0,1} U {v : d[v] <= D2} only.  One warp per query; distances are recomputed on the CUDA
// cores with __dp4a, which also makes this an independent check of the tensor-core path.
// One launch for a whole batch: blockIdx.y = job, blockIdx.x = group of 8 queries of it.  q_arena is where the query rows
// live (the image arena, or the scratch arena of the array-level API).
__global__ void __launch_bounds__(256)
tie_fixup_kernel(const uint8_t* __restrict__ arena, const uint8_t* __restrict__ q_arena, const PairJob* __restrict__ jobs,
                 int job0, KnnRecord* __restrict__ knn) {
  const PairJob J = jobs[job0 + blockIdx.y];
  const int q = static_cast<int>((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (!J.valid || q >= J.q_rows) return;
  const uint4* qp = reinterpret_cast<const uint4*>(q_arena + (size_t)(J.q_row0 + q) * kDim);
  uint4 qv[8];
  unsigned qn = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    qv[k] = qp[k];
    qn = __dp4a(qv[k].x, qv[k].x, qn); qn = __dp4a(qv[k].y, qv[k].y, qn);
    qn = __dp4a(qv[k].z, qv[k].z, qn); qn = __dp4a(qv[k].w, qv[k].w, qn);
  }
  const int D2 = knn[J.out_off + q].d2;
  int S = -1, T = -1, dS = 0, dT = 0, d0 = 0;
  for (int base = 0; base < J.db_rows; base += 32) {
    const int row = base + lane;
    int d = 0x7FFFFFFF;
    if (row < J.db_rows) {
      const uint4* dp = reinterpret_cast<const uint4*>(arena + (size_t)(J.db_row0 + row) * kDim);
      unsigned dn = 0, dot = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint4 dv = dp[k];
        dn = __dp4a(dv.x, dv.x, dn); dn = __dp4a(dv.y, dv.y, dn); dn = __dp4a(dv.z, dv.z, dn); dn = __dp4a(dv.w, dv.w, dn);
        dot = __dp4a(dv.x, qv[k].x, dot); dot = __dp4a(dv.y, qv[k].y, dot);
        dot = __dp4a(dv.z, qv[k].z, dot); dot = __dp4a(dv.w, qv[k].w, dot);
      }
      d = static_cast<int>(qn + dn - 2u * dot);
    }
    unsigned m = __ballot_sync(0xffffffffu, row < J.db_rows && (d <= D2 || row < 2));
    while (m) {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const int dv = __shfl_sync(0xffffffffu, d, l);
      const int v = base + l;
      if (v == 0) {
        d0 = dv;
      } else if (v == 1) {
        if (dv < d0) { T = 0; dT = d0; S = 1; dS = dv; } else { T = 1; dT = dv; S = 0; dS = d0; }
      } else if (dv < dT) {
        if (dS < dv) { T = v; dT = dv; } else { T = S; dT = dS; S = v; dS = dv; }
      }
    }
  }
  if (lane == 0) {
    KnnRecord r;
    r.idx1 = S; r.idx2 = T; r.d1 = dS; r.d2 = dT;
    *reinterpret_cast<int4*>(&knn[J.out_off + q]) = *reinterpret_cast<const int4*>(&r);
  }
}

// ------------------------------------------------------------------------------------------ K3
// 128 threads x <= 40 registers and ~1 KB of shared memory: a CTA of these kernels fits NEXT TO a resident knn2_kernel CTA
// (608 threads x 96 registers, 197 KB), so the compaction of batch k runs under the fused kernel of batch k + 1.
constexpr int kCompactThreads = 128;

// Exclusive block-wide rank of `flag` among the threads of the block + block total (ordered by threadIdx).
__device__ __forceinline__ int block_rank(bool flag, int* warp_tot /*[8] smem*/, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  __syncthreads();  // previous use of warp_tot finished
  if (lane == 0) warp_tot[warp] = __popc(m);
  __syncthreads();
  int before = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kCompactThreads / 32; ++w) {
    const int c = warp_tot[w];
    if (w < warp) before += c;
    tot += c;
  }
  total = tot;
  return before + __popc(m & ((1u << lane) - 1u));
}

// ------------------------------------------------------------------------------------------ rescan of ambiguous queries
// With prune_rho < 1 a failing query admits only rows with d <= rho * d(best).  Every row x skipped that way satisfies
// d(x) > rho * d(B) (B = recorded 2nd neighbour), so a recorded pair (d1, d2) that passes the ratio
// test is FINAL when it also passes against floor(rho * d2): no skipped row can bring d2 low enough to fail it.
// Otherwise the query is ambiguous and is matched again, exactly, by a second knn2_kernel launch over the gathered rows.
struct RescanSrc {
  int src_out_off;  // first record / first entry of the flagged list of the original pair
  int src_q_row0;   // arena row of query 0 of the original pair
  int first;        // first entry of the flagged list this rescan job covers
  int count;
  int dst_row0;     // row in the rescan query buffer == record offset in the rescan record buffer
};

__device__ __forceinline__ bool ratio_pass(int d1, int d2, float ratio_sq) {
  return __int2float_rn(d1) < __fmul_rn(ratio_sq, __int2float_rn(d2));  // DistanceRatioFilter, matching_filters.h:44
}

// One CTA per pair: ascending list of the ambiguous queries of the pair -> resc_idx[out_off + k], their number -> resc_cnt.
__global__ void __launch_bounds__(kCompactThreads)
flag_ambiguous_kernel(const PairJob* __restrict__ jobs, const KnnRecord* __restrict__ knn, float ratio_sq, float rho,
                      int* __restrict__ resc_idx, int* __restrict__ resc_cnt) {
  __shared__ int warp_tot[kCompactThreads / 32];
  const PairJob J = jobs[blockIdx.x];
  int base = 0;
  if (J.valid) {
    for (int q0 = 0; q0 < J.q_rows; q0 += kCompactThreads) {
      const int q = q0 + threadIdx.x;
      bool flag = false;
      if (q < J.q_rows) {
        const int4 r = *reinterpret_cast<const int4*>(&knn[J.out_off + q]);
        if (ratio_pass(r.z, r.w, ratio_sq)) {
          const int d_floor = __float2int_rd(__fmul_rd(rho, __int2float_rn(r.w)));
          flag = !ratio_pass(r.z, d_floor, ratio_sq);
        }
      }
      int tot;
      const int rank = block_rank(flag, warp_tot, tot);
      if (flag) resc_idx[J.out_off + base + rank] = q;
      base += tot;
    }
  }
  if (threadIdx.x == 0) resc_cnt[blockIdx.x] = base;
}

// One CTA per rescan job: copy the flagged queries (128 B each, 8 threads per row) and their norms into the rescan buffer.
__global__ void __launch_bounds__(256)
rescan_gather_kernel(const RescanSrc* __restrict__ src, const int* __restrict__ resc_idx, const uint8_t* __restrict__ arena,
                     const int* __restrict__ ccol, uint8_t* __restrict__ dst, int* __restrict__ dst_ccol) {
  const RescanSrc S = src[blockIdx.x];
  const int part = threadIdx.x & 7;
  for (int i = threadIdx.x >> 3; i < S.count; i += 32) {
    const int row = S.src_q_row0 + resc_idx[S.src_out_off + S.first + i];
    const uint4 v = *reinterpret_cast<const uint4*>(arena + (size_t)row * kDim + part * 16);
    *reinterpret_cast<uint4*>(dst + (size_t)(S.dst_row0 + i) * kDim + part * 16) = v;
    if (part == 0) dst_ccol[ccol_index(S.dst_row0 + i)] = ccol[ccol_index(row)];  // only the norm (bits 8..) is read
  }
}

// One CTA per rescan job: exact records back to where the pruned ones were.
__global__ void __launch_bounds__(256)
rescan_scatter_kernel(const RescanSrc* __restrict__ src, const int* __restrict__ resc_idx,
                      const KnnRecord* __restrict__ resc_knn, KnnRecord* __restrict__ knn) {
  const RescanSrc S = src[blockIdx.x];
  for (int i = threadIdx.x; i < S.count; i += 256) {
    const int q = resc_idx[S.src_out_off + S.first + i];
    *reinterpret_cast<int4*>(&knn[S.src_out_off + q]) = *reinterpret_cast<const int4*>(&resc_knn[S.dst_row0 + i]);
  }
}

// K3a, one CTA per pair:
//   (1) ratio test, fp32 exactly as DistanceRatioFilter: float(d1) < ratio_sq * float(d2)
//   (2) ordered list of passing queries -> tmp[out_off + k] = (idx1, q)
//   (3) drop the LAST passing query, then count elements whose _i differs from the predecessor's.
__global__ void __launch_bounds__(kCompactThreads)
ratio_filter_kernel(const PairJob* __restrict__ jobs, const KnnRecord* __restrict__ knn, float ratio_sq,
                    int2* __restrict__ tmp, int* __restrict__ n_pass, int* __restrict__ counts) {
  __shared__ int warp_tot[kCompactThreads / 32];
  const PairJob J = jobs[blockIdx.x];
  int base_out = 0;
  if (J.valid) {
    for (int q0 = 0; q0 < J.q_rows; q0 += kCompactThreads) {
      const int q = q0 + threadIdx.x;
      bool pass = false;
      int idx1 = 0;
      if (q < J.q_rows) {
        const int4 r = *reinterpret_cast<const int4*>(&knn[J.out_off + q]);
        idx1 = r.x;
        pass = ratio_pass(r.z, r.w, ratio_sq);
      }
      int tot;
      const int rank = block_rank(pass, warp_tot, tot);
      if (pass) tmp[J.out_off + base_out + rank] = make_int2(idx1, q);
      base_out += tot;
    }
  }
  __syncthreads();  // tmp writes of this CTA visible to this CTA
  const int n = base_out > 0 ? base_out - 1 : 0;  // drop-last
  int kept = 0;
  for (int k0 = 0; k0 < n; k0 += kCompactThreads) {
    const int k = k0 + threadIdx.x;
    bool keep = false;
    if (k < n) keep = (k == 0) || (tmp[J.out_off + k].x != tmp[J.out_off + k - 1].x);
    kept += __syncthreads_count(keep);
  }
  if (threadIdx.x == 0) { n_pass[blockIdx.x] = base_out; counts[blockIdx.x] = kept; }
}

// Exclusive scan of counts over the batch (single CTA of kScanThreads; batches are a few thousand pairs).
constexpr int kScanThreads = 128;
__global__ void __launch_bounds__(kScanThreads) scan_counts_kernel(const int* __restrict__ counts, int n,
                                                                   long long base, long long* __restrict__ offsets,
                                                                   long long* __restrict__ total_out) {
  constexpr int kWarps = kScanThreads / 32;
  __shared__ long long warp_sum[kWarps];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = base;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i0 = 0; i0 < n; i0 += kScanThreads) {
    const int i = i0 + threadIdx.x;
    const long long v = i < n ? counts[i] : 0;
    long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    long long before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const long long c = warp_sum[w];
      if (w < warp) before += c;
      tot += c;
    }
    const long long carry = carry_s;
    if (i < n) offsets[i] = carry + before + (x - v);
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) { offsets[n] = carry_s; *total_out = carry_s; }
}

// K3b, one CTA per pair: ordered scatter of the kept (_i,_j) into the dense match arena.
__global__ void __launch_bounds__(kCompactThreads)
dedup_scatter_kernel(const PairJob* __restrict__ jobs, const int2* __restrict__ tmp, const int* __restrict__ n_pass,
                     const long long* __restrict__ offsets, long long arena_base, int2* __restrict__ matches) {
  __shared__ int warp_tot[kCompactThreads / 32];
  const PairJob J = jobs[blockIdx.x];
  const int np = n_pass[blockIdx.x];
  const int n = np > 0 ? np - 1 : 0;
  long long out = offsets[blockIdx.x] - arena_base;
  for (int k0 = 0; k0 < n; k0 += kCompactThreads) {
    const int k = k0 + threadIdx.x;
    bool keep = false;
    int2 m = make_int2(0, 0);
    if (k < n) {
      m = tmp[J.out_off + k];
      keep = (k == 0) || (m.x != tmp[J.out_off + k - 1].x);
    }
    int tot;
    const int rank = block_rank(keep, warp_tot, tot);
    if (keep) matches[out + rank] = m;
    out += tot;
  }
}

// ------------------------------------------------------------------------------------------ second pass, planned on the device
// The host never learns how many queries are ambiguous (that would be a device->host round trip and a stream
// synchronisation per batch): ONE CTA turns the per-pair counts of flag_ambiguous_kernel into the work lists of the second
// pass.  Consecutive pairs against the same db image (the normal case, the pair list is (i, j)-ordered) share ONE job, so
// their few flagged queries fill 256-query items together instead of one nearly empty item per pair.
// block-wide exclusive scan of one value per thread (1024 threads); returns the exclusive prefix, total via reference
__device__ __forceinline__ int block_scan_1024(int v, int* warp_sum /*[32] smem*/, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  __syncthreads();  // previous use of warp_sum finished
  if (lane == 31) warp_sum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sum[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += y;
    }
    warp_sum[lane] = w;  // inclusive
  }
  __syncthreads();
  total = warp_sum[31];
  return (warp ? warp_sum[warp - 1] : 0) + (x - v);
}

__global__ void __launch_bounds__(1024)
plan_rescan_kernel(const PairJob* __restrict__ jobs, const int* __restrict__ resc_cnt, int nb, int cap,
                   RescanSrc* __restrict__ rsrc, PairJob* __restrict__ rjobs, int* __restrict__ ritem_start,
                   int* __restrict__ used /*[nb + 1] scratch*/, int* __restrict__ job_of /*[nb] scratch*/,
                   RescanMeta* __restrict__ meta) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  // (A) position of every pair's flagged queries in the gathered buffer
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int i0 = 0; i0 < nb; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const int v = i < nb ? resc_cnt[i] : 0;
    int tot;
    const int ex = block_scan_1024(v, warp_sum, tot);
    const int carry = carry_s;
    if (i < nb) used[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  const int total = carry_s;
  const bool overflow = total > cap;
  if (threadIdx.x == 0) used[nb] = total;
  __syncthreads();
  // gather / scatter sources (one per pair; empty when the pass is skipped)
  for (int i = threadIdx.x; i < nb; i += 1024) {
    const PairJob J = jobs[i];
    RescanSrc r;
    r.src_out_off = J.out_off; r.src_q_row0 = J.q_row0; r.first = 0;
    r.count = overflow ? 0 : resc_cnt[i];
    r.dst_row0 = used[i];
    rsrc[i] = r;
  }
  if (overflow || total == 0) {
    if (threadIdx.x == 0) { meta->n_items = 0; meta->total = total; meta->overflow = overflow ? 1 : 0; meta->n_jobs = 0; }
    return;
  }
  // (B) jobs = maximal runs of adjacent pairs against the same db image
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int i0 = 0; i0 < nb; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    int head = 0;
    if (i < nb) head = (i == 0 || jobs[i].db_row0 != jobs[i - 1].db_row0 || jobs[i].db_rows != jobs[i - 1].db_rows) ? 1 : 0;
    int tot;
    const int ex = block_scan_1024(head, warp_sum, tot);
    const int carry = carry_s;
    if (i < nb) {
      const int j = carry + ex + head - 1;  // job of pair i
      job_of[i] = j;
      if (head) {
        PairJob R = jobs[i];
        R.q_row0 = used[i];
        R.out_off = used[i];
        R.q_rows = 0;  // filled in below
        R.valid = 1;
        rjobs[j] = R;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  const int n_rj = carry_s;
  __syncthreads();
  // rows of a job = start of the next job (or the total) - its own start: the LAST pair of every run knows the end
  for (int i = threadIdx.x; i < nb; i += 1024)
    if (i == nb - 1 || job_of[i + 1] != job_of[i]) rjobs[job_of[i]].q_rows = used[i + 1] - rjobs[job_of[i]].q_row0;
  __syncthreads();
  // (C) work items per job
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int j0 = 0; j0 < n_rj; j0 += 1024) {
    const int j = j0 + threadIdx.x;
    const int v = j < n_rj ? (rjobs[j].q_rows + 2 * kBlockQ - 1) / (2 * kBlockQ) : 0;
    int tot;
    const int ex = block_scan_1024(v, warp_sum, tot);
    const int carry = carry_s;
    if (j < n_rj) ritem_start[j] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ritem_start[n_rj] = carry_s;
    meta->n_items = carry_s; meta->total = total; meta->overflow = 0; meta->n_jobs = n_rj;
  }
}

// ------------------------------------------------------------------------------------------ row 13 on the GPU
// IndexedMatchDecorator<float>::getDeduplicated (indexed_match_decorator.h:90-104): one thread per pair builds the
// red-black tree the reference's std::set would build (rbtree_dedup.cuh) over the pair's matches, keyed by the (x, y) of
// the LEFT feature, and writes the survivors in set order.  feats: [arena rows] (x, y) of every descriptor row; nodes:
// scratch, [matches of the batch + pairs of the batch] (every pair needs one header node).
// Shared-memory form: one warp per pair; lane 0 builds the tree in 16-byte nodes held in shared memory (a dependent
// shared-memory load per level instead of an L2 round trip: ~40x less latency per insertion), the other lanes gather the keys
// and write the survivors.  Pairs with more than `cap` matches are left to dedup_xy_kernel below.
constexpr int kDedupSmemCap = 2047;       // 2048 nodes of 16 B + 2048 order entries of 2 B = 36 KB per CTA: 6 pairs in flight per SM
constexpr int kDedupSmemCapSmall = 1023;  // 18 KB: fits beside a resident knn2_kernel CTA (batches that have a successor)
template <int kCap>
__global__ void __launch_bounds__(32)
dedup_xy_smem_kernel(const PairJob* __restrict__ jobs, const int2* __restrict__ matches, const long long* __restrict__ offsets,
                     const int* __restrict__ counts, const float2* __restrict__ feats, int2* __restrict__ out,
                     int* __restrict__ counts2) {
  __shared__ RbNode16 nd[kCap + 1];
  __shared__ unsigned short ord[kCap + 1];
  __shared__ int kept_s;
  const int p = blockIdx.x, lane = threadIdx.x;
  const int n = counts[p];
  if (n > kCap) return;
  if (n == 0) { if (lane == 0) counts2[p] = 0; return; }
  const long long off = offsets[p];
  const int db_row0 = jobs[p].db_row0;
  const int2* m = matches + off;
  for (int k = lane; k < n; k += 32) {
    const float2 f = feats[db_row0 + m[k].x];
    nd[k].x = f.x;
    nd[k].y = f.y;
  }
  __syncwarp();
  if (lane == 0) kept_s = rbtree_dedup(nd, n, ord);
  __syncwarp();
  const int kept = kept_s;
  int2* o = out + off;
  for (int k = lane; k < kept; k += 32) o[k] = m[ord[k]];
  if (lane == 0) counts2[p] = kept;
}

// Global-memory form, one thread per pair, for the pairs the shared-memory kernel left (more than min_n matches).
__global__ void __launch_bounds__(64)
dedup_xy_kernel(const PairJob* __restrict__ jobs, int nb, int min_n, const int2* __restrict__ matches, const long long* __restrict__ offsets,
                const int* __restrict__ counts, const float2* __restrict__ feats, RbNode* __restrict__ nodes,
                int* __restrict__ order /*[matches of the batch]*/, int2* __restrict__ out, int* __restrict__ counts2) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nb) return;
  const int n = counts[p];
  if (n <= min_n) return;
  const long long off = offsets[p];
  const int db_row0 = jobs[p].db_row0;
  RbNode* nd = nodes + off + p;
  const int2* m = matches + off;
  for (int k = 0; k < n; ++k) {
    const float2 f = feats[db_row0 + m[k].x];
    nd[k].x = f.x;
    nd[k].y = f.y;
  }
  int* ord = order + off;
  const int kept = rbtree_dedup(nd, n, ord);
  int2* o = out + off;
  for (int k = 0; k < kept; ++k) o[k] = m[ord[k]];
  counts2[p] = kept;
}

// Close the gaps the de-dup left: pair p's kept matches move from src[old_off[p] ..] to dst[new_off[p] ..].
__global__ void __launch_bounds__(kCompactThreads)
compact_pairs_kernel(const int2* __restrict__ src, const long long* __restrict__ old_off, const long long* __restrict__ new_off,
                     const int* __restrict__ counts2, int2* __restrict__ dst) {
  const int n = counts2[blockIdx.x];
  const long long so = old_off[blockIdx.x], d0 = new_off[blockIdx.x];
  for (int k = threadIdx.x; k < n; k += kCompactThreads) dst[d0 + k] = src[so + k];
}

}  // namespace mvgcuda
