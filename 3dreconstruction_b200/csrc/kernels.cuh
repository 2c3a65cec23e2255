// kernels.cuh -- the sm_100a kernels of libmvgcuda.
//
//   K1 row_consts_kernel   per arena row:  ccol = (||d||^2 << 8) | (row & 255)
//   K2 knn2_kernel         fused u8xu8->s32 tcgen05 GEMM + ||d||^2 - 2 q.d + running top-2
//                          (replaces matcher_brute_force.h:117-131 + metric.h:57-81 +
//                          indexed_sort.h:52-66; the distance matrix lives only in TMEM/registers)
//   K3 ratio/compaction    DistanceRatioFilter (matching_filters.h:27-47), drop-last loop
//                          (matcher_all_in_memory.h:117-122), unique-on-_i (indexed_match.h:49-55)
//
// All arithmetic on the path is exact int32; the only fp32 operation is the ratio test
// float(d1) < ratio_sq * float(d2) (one __fmul_rn, strict <), as in the reference.
#pragma once
#include "ptx.cuh"

#ifndef MVGCUDA_PROBE
#define MVGCUDA_PROBE 0  // developer ceilings only (results wrong): 1 TMEM drain, 2 fast filter path, 8 TMA+MMA feed; 0 = the product
#endif

namespace mvgcuda {

constexpr int kDim = 128;      // descriptor bytes == GEMM K
constexpr int kBlockQ = 128;   // query rows per block   (MMA M per CTA, one TMEM lane per query)
constexpr int kTileDb = 256;   // db rows per tile       (one TMEM column per db row)
constexpr int kSlotsA = 2;     // query block double buffer
constexpr int kSlotsC = 8;     // per-column constant ring (TMA -> epilogue), outlives the B stage
constexpr int kAccBufs = 2;    // TMEM accumulator double buffer (2 x 256 columns = all 512)
constexpr int kRowAlign = 256; // every image starts at a multiple of this in the arena
constexpr int kPadNorm = 0x7FFFFF;  // "norm" of padding rows: > 128*255^2, so they never win

constexpr uint32_t kBytesA = kBlockQ * kDim;        // 16 KB
constexpr uint32_t kBytesB = kTileDb * kDim;        // 32 KB of db rows per tile (a CTA pair holds 16 KB each)
constexpr uint32_t kBytesBRing = 4 * kBytesB;       // db tile ring (TMA -> MMA): 4 stages of 32 KB, or 8 of 16 KB per CTA of a pair
constexpr int kChunk = 16;                     // db rows per filter decision in the epilogue
constexpr int kTileC = kTileDb + kTileDb / kChunk;  // per-tile constants: 256 packed (norm<<8|col) + 16 chunk-min norms
constexpr uint32_t kBytesC = kTileC * sizeof(int);  // 1088 B

constexpr int kEpiParts = 4;     // warps per TMEM lane quadrant; they share 32 queries: 2 column halves x 2 tile parities
constexpr int kNumEpiWarps = 4 * kEpiParts;
constexpr int kPartCols = kTileDb / 2;                // 128 columns of every other tile per warp
constexpr int kFirstEpiWarp = 2;                       // warp 0: TMA producer + TMEM allocator, warp 1: MMA issuer
constexpr int kKnnThreads = 32 * (kFirstEpiWarp + kNumEpiWarps);  // 576 threads -> 112 registers per thread

// Pipeline shapes (template parameters of knn2_kernel):
//   kPair   two CTAs of a cluster (one TPC) work on two query blocks of the same pair with ONE tcgen05.mma.cta_group::2
//           per K step (M = 256): each CTA stages only half of every db tile, so the shared-memory read rate of the tensor
//           pipe per SM drops from 96 to 64 B/clk (N = 256) -- the single-CTA N = 128 shape needs 128 B/clk and starves
//           (profiles/r02a_variants.log) -- and the L2 -> shared traffic per SM halves.
//   kSplit  (pair only) a tile is issued as two N = 128 MMA groups with their own barriers: four accumulator buffers of
//           128 columns instead of two of 256, so a buffer goes back to the MMA warp as soon as ITS four warps have
//           drained it, and the TMEM drain of one buffer no longer sits on the critical path of the next MMA.
template <bool kPair, bool kSplit>
struct KnnShape {
  static_assert(kPair || !kSplit, "N = 128 MMAs starve on shared-memory bandwidth without the CTA pair");
  static constexpr int kStagesB = kPair ? 8 : 4;
  static constexpr uint32_t kStageBytes = kBytesBRing / kStagesB;  // per CTA
  static constexpr int kAccBars = kSplit ? 2 * kAccBufs : kAccBufs;
  static constexpr int kMmaN = kSplit ? kPartCols : kTileDb;
  static constexpr int kMmaM = kPair ? 2 * kBlockQ : kBlockQ;
  static constexpr int kBoxDb = kPair ? (kSplit ? 64 : 128) : 256;  // rows of one TMA box of the db tensor map
};

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

template <bool kPair, bool kSplit>
struct KnnSmem {
  using Shape = KnnShape<kPair, kSplit>;
  alignas(1024) uint8_t a[kSlotsA][kBytesA];
  alignas(1024) uint8_t b[Shape::kStagesB][Shape::kStageBytes];
  alignas(16) int c[kSlotsC][2 * kTileC];  // schedules 0/1 use one tile per slot, schedule 2 a tile pair
  uint64_t a_full[kSlotsA], a_empty[kSlotsA];
  uint64_t b_full[Shape::kStagesB], b_empty[Shape::kStagesB];
  uint64_t c_full[kSlotsC], c_empty[kSlotsC];  // (epilogue_slices relies on c_empty following c_full)
  uint64_t acc_full[Shape::kAccBars], acc_empty[Shape::kAccBars];
  uint32_t tmem_base;
  int bound[2][kBlockQ];  // per (item parity, query): best-known 2nd-smallest t, atomically tightened by all parts
  alignas(16) int4 xchg[2][kEpiParts - 1][kBlockQ];  // parts 1.. hand their top-2 to part 0 at the end of an item
};

struct KnnParams {
  const int* __restrict__ ccol;        // K1 output, [arena_rows / 256][kTileC]: per-row constants of the db operand
  const int* __restrict__ hmin;        // tail of K1's output: min ||d||^2 of every 128 db rows
  const int* __restrict__ qcol;        // same layout for the rows the QUERY tensor map addresses (== ccol, except rescans)
  const PairJob* __restrict__ jobs;    // [n_jobs]
  const int* __restrict__ item_start;  // [n_jobs+1] prefix sum of query blocks per job
  int n_jobs;
  int n_items;
  KnnRecord* __restrict__ out;
  int two;  // always 2; a run-time value so that 2*x+T stays an IMAD (idle FMA pipe) instead of an IADD3 (ALU pipe, the bottleneck)
  // Ratio-aware pruning (see epi_chunk16): the fp32 squared ratio of the Lowe test the records feed, or FLT_MAX when the
  // caller needs the exact 2nd neighbour of EVERY query (array-level API, ratio > 1 with the tie fix-up).
  float prune_ratio;
  float prune_rho;  // in (0, 1]: a failing query admits only rows with d <= rho * d(best); 1 = no rescans ever needed
};

// ------------------------------------------------------------------------------------------ K1
// Per-row constants, laid out per 256-row tile as [256 x ((||d||^2 << 8) | col)] [16 x min ||d||^2 of each 16-row chunk],
// followed (after the last tile) by one min ||d||^2 per 128 rows -- the filter constant of one epilogue warp's slice of a
// tile, which the warp keeps in a register.  8 threads per 128-byte row (one 16-B load each), __dp4a squares, 3 shuffles;
// 128 rows per block.
__device__ __forceinline__ int ccol_index(int row) { return (row >> 8) * kTileC + (row & 255); }
__host__ __device__ constexpr size_t ccol_ints(int arena_rows) {
  return static_cast<size_t>(arena_rows / kTileDb) * kTileC + static_cast<size_t>(arena_rows / kPartCols);
}
constexpr int kK1Rows = kPartCols;  // rows per block of K1

__global__ void __launch_bounds__(8 * kK1Rows)
row_consts_kernel(const uint8_t* __restrict__ arena, const int* __restrict__ img_row0, const int* __restrict__ img_rows,
                  int n_images, int arena_rows, int* __restrict__ ccol) {
  __shared__ int norms[kK1Rows];
  const int row = blockIdx.x * kK1Rows + (threadIdx.x >> 3);  // arena_rows is a multiple of 256
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
  if (threadIdx.x < kK1Rows / kChunk) {
    int m = norms[threadIdx.x * kChunk];
#pragma unroll
    for (int k = 1; k < kChunk; ++k) m = min(m, norms[threadIdx.x * kChunk + k]);
    const int row0 = blockIdx.x * kK1Rows + threadIdx.x * kChunk;
    ccol[(row0 >> 8) * kTileC + kTileDb + ((row0 & 255) >> 4)] = m;
    // minimum of the block's 128 rows (8 chunk minima live in the first 8 lanes of warp 0)
    m = min(m, __shfl_xor_sync(0xffu, m, 1));
    m = min(m, __shfl_xor_sync(0xffu, m, 2));
    m = min(m, __shfl_xor_sync(0xffu, m, 4));
    if (threadIdx.x == 0) ccol[static_cast<size_t>(arena_rows / kTileDb) * kTileC + blockIdx.x] = m;
  }
}

// ------------------------------------------------------------------------------------------ K2
__device__ __forceinline__ void locate_item(const KnnParams& p, int item, int& job, int& qb) {
  int lo = 0, hi = p.n_jobs - 1;
  while (lo < hi) {  // last job with item_start[job] <= item
    const int mid = (lo + hi + 1) >> 1;
    if (p.item_start[mid] <= item) lo = mid; else hi = mid - 1;
  }
  job = lo;
  qb = item - p.item_start[lo];
}

// Sorted pair (lo <= hi) helpers for the exact top-2 of a chunk: a merge tree has depth ~10 and plenty of
// instruction-level parallelism, where 16 serial insertions form a 32-deep dependency chain.
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

// The exact step of a chunk: which groups of 4 rows can still matter, exact packed keys for those, top-2 merge, and the
// new bound (items 2-5 of DESIGN.md section 4).  g[k] = max of the raw dot products of group k.
__device__ __forceinline__ void epi_exact16(const int32_t* __restrict__ x, const int* __restrict__ g, const uint32_t cs_saddr,
                                            const int cmin, const int g1t, const int g2t, const int qn,
                                            const float prune_ratio, const float prune_rho, const uint32_t bound_saddr,
                                            const int two, int& l1, int& l2, int& T) {
    // which groups of 4 rows can still matter (all four votes issued back to back)
    const bool h0 = __any_sync(0xffffffffu, two * g[0] + T >= cmin);
    const bool h1 = __any_sync(0xffffffffu, two * g[1] + T >= cmin);
    const bool h2 = __any_sync(0xffffffffu, two * g[2] + T >= cmin);
    const bool h3 = __any_sync(0xffffffffu, two * g[3] + T >= cmin);
    const bool h[4] = {h0, h1, h2, h3};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (h[k]) {
        const int4 cc = ptx::lds128(cs_saddr + 16 * k);  // warp-uniform address: smem broadcast
        const int p0 = static_cast<int>(static_cast<uint32_t>(cc.x) - 512u * static_cast<uint32_t>(x[4 * k + 0]));
        const int p1 = static_cast<int>(static_cast<uint32_t>(cc.y) - 512u * static_cast<uint32_t>(x[4 * k + 1]));
        const int p2 = static_cast<int>(static_cast<uint32_t>(cc.z) - 512u * static_cast<uint32_t>(x[4 * k + 2]));
        const int p3 = static_cast<int>(static_cast<uint32_t>(cc.w) - 512u * static_cast<uint32_t>(x[4 * k + 3]));
        const Pair2 c = merge2(sort2(p0, p1), sort2(p2, p3));
        const int nl2 = __vimin3_s32(max(l1, c.lo), l2, c.hi);
        l1 = min(l1, c.lo);
        l2 = nl2;
      }
    }
    // Smallest (c1) and 2nd smallest (c2) t over {running top-2 of earlier tiles} U {this tile's top-2}: two rows this
    // warp has really seen.  Rows with t > c2 can never enter the top-2.  Ratio-aware pruning goes further: when c1
    // already fails the Lowe test against c2 (d1 >= ratio * d2, the reference's fp32 expression), rows with c1 < t <= c2
    // cannot matter either -- such a row could only become the FINAL 2nd neighbour, and then d2 <= c2 makes the query fail
    // the test whether it is recorded or not; and if a better 1st neighbour turns up later, the 2nd neighbour is at most
    // c1.  So the bound drops to c1: the record's idx1/d1 stay exact for every query, d2/idx2 stay exact for every query
    // that passes, and a failing query keeps a d2 that still fails (DESIGN.md "ratio-aware pruning" has the proof).
    const int l1t = l1 >> 8;
    const int c1 = min(g1t, l1t);
    const int c2 = __vimin3_s32(max(g1t, l1t), l2 >> 8, g2t);
    const bool passes = __int2float_rn(qn + c1) < __fmul_rn(prune_ratio, __int2float_rn(qn + c2));
    // failing: admit d <= rho * d(c1), rounded up (rho = 1 gives exactly c1).  With rho < 1 a failing query may even miss
    // its true nearest row (one within rho..1 of the recorded one cannot pass the test either); what stays exact is every
    // pass/fail decision and the nearest row of every passing query -- after the second pass over the records that
    // flag_ambiguous_kernel cannot decide (tests/test_pruning_model.py checks the rule).
    const int tf = __float2int_ru(__fmul_ru(prune_rho, __int2float_rn(qn + c1))) - qn;
    T = min(T, passes ? c2 : tf);
    ptx::red_min_shared(bound_saddr, T);  // the warps scanning the other columns of these queries tighten their filter now
}

// One filter step over a chunk of 16 db rows (TMEM columns), entirely on the raw dot products x = q.d:
//   some row of the chunk can still enter the top-2  =>  min_norm(chunk) - 2*max(x) <= T   (T: best-known 2nd-smallest
//   t = ||d||^2 - 2 q.d of this query, ties admitted).  8 max ops + 1 add + 1 compare + 1 vote per 16 rows and no
//   shared-memory traffic.  Only when some lane of the warp passes is the chunk examined in groups of 4, and only for the
//   groups that pass are exact packed keys ((||d||^2 - 2x) << 8 | col, one IMAD each) formed and merged (top-2 of 4 by
//   a small sorting network) into the tile's running top-2.
// Exactness: the test is necessary for membership in the final top-2, so no candidate is ever lost; ties are decided
// by the packed compare (same tile) and the strict merge (earlier tile wins), i.e. lowest db row.
__device__ __forceinline__ void epi_chunk16(const int32_t* __restrict__ x, const uint32_t cs_saddr, const int cmin,
                                            const int g1t, const int g2t, const int qn, const float prune_ratio, const float prune_rho,
                                            const uint32_t bound_saddr, const int two, int& l1, int& l2, int& T) {
#if MVGCUDA_PROBE == 1  // TMEM drain only: no filter work at all
  l1 = min(l1, x[0]);
  return;
#endif
  const int m = max16(x);  // 8 ops; the per-group maxima are only formed when the chunk goes to the exact step
#if MVGCUDA_PROBE == 2  // fast path only
  l1 = min(l1, m + T + cmin);
  return;
#endif
  if (__any_sync(0xffffffffu, two * m + T >= cmin)) {
    int g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = max(__vimax3_s32(x[4 * k], x[4 * k + 1], x[4 * k + 2]), x[4 * k + 3]);
    epi_exact16(x, g, cs_saddr, cmin, g1t, g2t, qn, prune_ratio, prune_rho, bound_saddr, two, l1, l2, T);
  }
}

constexpr int kTInit = 0x3FFFFFFF;  // "no bound yet": 2*x + kTInit cannot overflow and passes every chunk

// (t, index) lexicographic order: smaller distance first, lower db row on ties
__device__ __forceinline__ bool cand_less(int ta, int ia, int tb, int ib) { return ta < tb || (ta == tb && ia < ib); }

// End of a tile in which some chunk was examined exactly (schedule 2): merge the tile's top-2 (packed keys l1 <= l2) into
// the running top-2 -- ties keep the earlier (lower) index; branch-free selects; an untouched l1/l2 (0x7FFFFFFF) decodes
// to t = 0x7FFFFF, larger than any real t -- and tighten the bound (ratio-aware pruning, see epi_exact16).
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
  // g1t <= g2t are now the two smallest t of rows this warp has really seen
  const bool passes = __int2float_rn(qn + g1t) < __fmul_rn(prune_ratio, __int2float_rn(qn + g2t));
  const int tf = __float2int_ru(__fmul_ru(prune_rho, __int2float_rn(qn + g1t))) - qn;
  T = min(T, passes ? g2t : tf);
  ptx::red_min_shared(bound_saddr, T);  // the warps scanning the other columns of these queries tighten their filter now
}

// ===================== epilogue, schedule 2 ("slices") =====================
// Warp (quad, g, sub) owns TMEM lanes 32*quad.., and of EVERY tile the 64 columns [128 g + 64 sub, +64): four 16-column
// chunks = 64 registers, loaded at once.  The accumulator buffer is handed back to the MMA warp as soon as those loads
// have landed -- before any arithmetic -- so the time a buffer spends outside the tensor pipe is one TMEM-load latency
// instead of a filter pass (in the other schedules that hand-back sits on the critical path of the next MMA into the
// buffer; profiles/r02e: MMA warp 32 % of its time in acc_empty waits, epilogue warps 18 % in acc_full waits).
// The epilogue is bound by instruction issue (4 warps per scheduler), so the hot path is kept to the minimum:
//   * filter: a row of the slice can only matter if ||d||^2 - 2 q.d <= T, hence only if 2 q.d >= hm - T with
//     hm = min ||d||^2 of the tile half (one prefetched register): 35 max ops over the 64 raw dot products, ONE compare,
//     one vote, one branch per tile -- no shared memory, no per-chunk control flow;
//   * tiles are walked two at a time, so every barrier / TMEM address of a step is a loop-invariant register and the
//     barrier phases are two bits that flip (an item always starts in accumulator buffer 0);
//   * only when some lane passes: which chunks (four votes), exact packed keys of ALL 16 rows of those chunks straight
//     from the registers, top-2 merge, ONE bound update and ONE merge into the running top-2 per tile.
// The first tile of an item skips the filter (no useful bound yet).
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
// exact packed keys ((||d||^2 - 2 q.d) << 8 | column) of all 16 rows of a chunk, merged into the tile's top-2
__device__ __forceinline__ void exact_chunk_all(const int32_t* __restrict__ x, const uint32_t cs_saddr, int& l1, int& l2) {
#pragma unroll
  for (int k = 0; k < 4; ++k) top2_of4(x + 4 * k, ptx::lds128(cs_saddr + 16 * k), l1, l2);
}

// Values the compiler must keep in a register instead of re-deriving them (it otherwise rematerialises the aligned base
// of the dynamic shared memory -- a dozen uniform-datapath instructions -- in front of every barrier access of the loop).
__device__ __forceinline__ uint32_t keep_reg(uint32_t x) {
  asm volatile("" : "+r"(x));
  return x;
}

template <bool kPair, bool kSplit>
__device__ __forceinline__ void epilogue_slices(KnnSmem<kPair, kSplit>& s, const KnnParams& p, const uint32_t tmem_base,
                                                const int warp, const int lane, const uint32_t rank, const int worker,
                                                const int n_workers) {
  using Smem = KnnSmem<kPair, kSplit>;
  const int quad = warp & 3;  // a warp may only touch its own TMEM lane quadrant
  const int part = (warp - kFirstEpiWarp) >> 2;
  const int g = part >> 1;    // column half of the tile == its own accumulator buffer when kSplit
  const int sub = part & 1;   // 64-column slice of that half
  const int row = quad * 32 + lane;
  const uint32_t tslice = keep_reg(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + g * kPartCols + sub * 64);
  // loop-invariant shared-memory addresses (32-bit, shared space): the two accumulator buffers' barriers of half g,
  // the constants ring, this query's bound slots
  const uint32_t sbase = ptx::smem_u32(&s);
  const uint32_t full0 = keep_reg(sbase + offsetof(Smem, acc_full) + 8u * (kSplit ? g : 0));
  const uint32_t full1 = keep_reg(sbase + offsetof(Smem, acc_full) + 8u * (kSplit ? 2 + g : 1));
  uint32_t e0 = sbase + offsetof(Smem, acc_empty) + 8u * (kSplit ? g : 0);
  uint32_t e1 = sbase + offsetof(Smem, acc_empty) + 8u * (kSplit ? 2 + g : 1);
  if constexpr (kPair) { e0 = ptx::mapa_shared(e0, 0); e1 = ptx::mapa_shared(e1, 0); }
  const uint32_t empty0 = keep_reg(e0), empty1 = keep_reg(e1);
  const uint32_t c_ring = keep_reg(sbase + offsetof(Smem, c) + static_cast<uint32_t>(g * kPartCols + sub * 64) * 4u);  // the slice's per-column constants in slot 0
  const uint32_t c_bars = keep_reg(sbase + offsetof(Smem, c_full));
  const uint32_t bound0 = keep_reg(sbase + offsetof(Smem, bound) + 4u * row);
  uint32_t ph = 0;     // bit b: parity of the next acc_full phase of buffer b
  uint32_t c_cnt = 0;  // tile PAIRS seen so far == position in the constants ring (one slot holds two tiles)
  uint32_t item_it = 0;
  ptx::sts32(bound0, kTInit);
  ptx::sts32(bound0 + 4u * kBlockQ, kTInit);
  asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * kEpiParts) : "memory");
  for (int item = worker; item < p.n_items; item += n_workers, ++item_it) {
    int job, qb;
    locate_item(p, item, job, qb);
    if constexpr (kPair) qb = 2 * qb + static_cast<int>(rank);
    const PairJob J = p.jobs[job];
    const int q_local = qb * kBlockQ + row;
    const bool q_ok = q_local < J.q_rows;
    const int ntiles = (J.db_rows + kTileDb - 1) / kTileDb;
    const int qn = p.qcol[ccol_index(J.q_row0 + min(q_local, J.q_rows - 1))] >> 8;  // ||q||^2 (dist = qn + t)
    int g1t = 0x7FFFFFFF, g2t = 0x7FFFFFFF, g1i = -1, g2i = -1;  // running best two (t-domain) of this thread's columns
    int T = kTInit;                                              // admit t <= T
    const uint32_t bound_saddr = bound0 + (item_it & 1u) * (4u * kBlockQ);
    ptx::sts32(bound0 + ((item_it & 1u) ^ 1u) * (4u * kBlockQ), kTInit);  // idle slot (every part left the previous item at the barrier below)
    const int* hm_ptr = p.hmin + (J.db_row0 >> 7) + g;  // min ||d||^2 of tile t's half g: hm_ptr[2 t]
    int hm_next = ntiles > 0 ? __ldg(hm_ptr) : 0;

    // One tile in accumulator buffer kBuf (compile-time): everything indexed by the buffer is a loop invariant.
#define MVG_TILE_STEP(kBuf, t_expr)                                                                                       \
    {                                                                                                                      \
      const int t = (t_expr);                                                                                              \
      ptx::mbar_wait_addr(kBuf ? full1 : full0, (ph >> kBuf) & 1u);                                                        \
      ph ^= (1u << kBuf);                                                                                                  \
      ptx::tc_fence_after();                                                                                               \
      int32_t v0[16], v1[16], v2[16], v3[16];                                                                              \
      ptx::tmem_ld_32x32b_x16(tslice + kBuf * kTileDb, v0);                                                                \
      ptx::tmem_ld_32x32b_x16(tslice + kBuf * kTileDb + 16, v1);                                                           \
      ptx::tmem_ld_32x32b_x16(tslice + kBuf * kTileDb + 32, v2);                                                           \
      ptx::tmem_ld_32x32b_x16(tslice + kBuf * kTileDb + 48, v3);                                                           \
      const int hm = hm_next;                                                                                              \
      hm_ptr += 2;                                                                                                         \
      if (t + 1 < ntiles) hm_next = __ldg(hm_ptr);                                                                         \
      T = min(T, ptx::lds32_volatile(bound_saddr));                                                                        \
      ptx::tmem_ld_wait_for4(v0, v1, v2, v3);                                                                              \
      ptx::tc_fence_before();                                                                                              \
      __syncwarp();                                                                                                        \
      if (lane == 0) { /* the buffer goes back to the MMA warp now */                                                      \
        if constexpr (kPair) ptx::mbar_arrive_cluster(kBuf ? empty1 : empty0);                                             \
        else ptx::mbar_arrive_addr(kBuf ? empty1 : empty0);                                                                \
      }                                                                                                                    \
      const int m0 = max16(v0), m1 = max16(v1), m2 = max16(v2), m3 = max16(v3);                                            \
      const int thr = (hm - T + 1) >> 1; /* 2 m >= hm - T  <=>  m >= ceil((hm - T) / 2) */                                 \
      if (__any_sync(0xffffffffu, max(__vimax3_s32(m0, m1, m2), m3) >= thr) || t == 0) {                                   \
        const uint32_t sc = c_cnt & (kSlotsC - 1);                                                                         \
        ptx::mbar_wait_addr(c_bars + 8u * sc, (c_cnt / kSlotsC) & 1); /* per-column constants: on this path only */        \
        const uint32_t cs = c_ring + sc * (2u * kBytesC) + kBuf * kBytesC;                                                 \
        int l1 = 0x7FFFFFFF, l2 = 0x7FFFFFFF;                                                                              \
        const bool first = t == 0;                                                                                         \
        if (first || __any_sync(0xffffffffu, m0 >= thr)) exact_chunk_all(v0, cs, l1, l2);                                  \
        if (first || __any_sync(0xffffffffu, m1 >= thr)) exact_chunk_all(v1, cs + 64, l1, l2);                             \
        if (first || __any_sync(0xffffffffu, m2 >= thr)) exact_chunk_all(v2, cs + 128, l1, l2);                            \
        if (first || __any_sync(0xffffffffu, m3 >= thr)) exact_chunk_all(v3, cs + 192, l1, l2);                            \
        slice_commit(l1, l2, t, qn, p.prune_ratio, p.prune_rho, bound_saddr, g1t, g1i, g2t, g2i, T);                       \
      }                                                                                                                    \
    }
    for (int tp = 0; tp < ntiles; tp += 2) {
      MVG_TILE_STEP(0, tp)
      if (tp + 1 < ntiles) MVG_TILE_STEP(1, tp + 1)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_addr(c_bars + 8u * (kSlotsC + (c_cnt & (kSlotsC - 1))));  // c_empty follows c_full in KnnSmem
      ++c_cnt;
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
      *reinterpret_cast<int4*>(&p.out[J.out_off + q_local]) = *reinterpret_cast<const int4*>(&r);
    }
  }
}

// kSched selects the epilogue schedule of a tile (all are bit-identical in their results):
//   0  in place: each chunk is filtered and, if needed, examined exactly while it sits in registers; the TMEM loads
//      rotate through four register sets.
//   1  filter first: all eight chunks are reduced to their maxima, the few that still matter are loaded a second
//      time for the exact step (one rolled copy).  First tiles of an item cost more (every chunk matters and is
//      loaded twice), later tiles less.
//   2  slices: every warp takes 64 columns of EVERY tile and hands the accumulator back before any arithmetic
//      (epilogue_slices above).
// kPair / kSplit: see KnnShape.  In a pair the work item is (image pair, TWO consecutive query blocks): the CTA of cluster
// rank r owns block 2*item + r (an odd last block leaves rank 1 with no valid query; it still takes part in the MMAs).
template <int kSched, bool kPair, bool kSplit>
__global__ void __launch_bounds__(kKnnThreads, 1)  // 18 warps: two schedulers host 5 warps, so 16 K / 5 warps = 96 registers per thread
knn2_kernel(const __grid_constant__ CUtensorMap tmap_q,   // box 128 rows x 128 B
            const __grid_constant__ CUtensorMap tmap_db,  // box KnnShape::kBoxDb rows x 128 B
            const KnnParams p) {
  using Shape = KnnShape<kPair, kSplit>;
  using Smem = KnnSmem<kPair, kSplit>;
  constexpr int kStagesB = Shape::kStagesB;
  constexpr int kAccBars = Shape::kAccBars;
  extern __shared__ uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // warp / lane through a shuffle: values ptxas cannot rematerialise.  Derived directly from %tid they are re-read with
  // S2R (~25 clocks, and a dependent shift) inside the tile loop instead of being kept in a register.
  const int lane = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x & 31), static_cast<int>(threadIdx.x & 31));
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;        // 0 = leader: issues the MMAs of the pair
  const int worker = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_workers = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_db);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kSlotsA; ++i) { ptx::mbar_init(&s.a_full[i], 1); ptx::mbar_init(&s.a_empty[i], 1); }
    for (int i = 0; i < kStagesB; ++i) { ptx::mbar_init(&s.b_full[i], 1); ptx::mbar_init(&s.b_empty[i], 1); }
    // a tile's constants are consumed by the 8 epilogue warps of one tile-parity group (of this CTA)
    for (int i = 0; i < kSlotsC; ++i) { ptx::mbar_init(&s.c_full[i], 1); ptx::mbar_init(&s.c_empty[i], kSched == 2 ? kNumEpiWarps : kNumEpiWarps / 2); }
    // an accumulator buffer goes back to the (leader's) MMA warp when its warps of BOTH CTAs of a pair have drained it
    for (int i = 0; i < kAccBars; ++i) {
      ptx::mbar_init(&s.acc_full[i], 1);
      ptx::mbar_init(&s.acc_empty[i], (kPair ? 2 : 1) * (kSched == 2 ? 2 : 1) * kNumEpiWarps / kAccBars);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    if constexpr (kPair) ptx::tmem_alloc_pair<512>(&s.tmem_base); else ptx::tmem_alloc<512>(&s.tmem_base);
  }
  ptx::tc_fence_before();
  if constexpr (kPair) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    {
      uint32_t a_it = 0, b_it = 0, c_it = 0;
      for (int item = worker; item < p.n_items; item += n_workers, ++a_it) {
        int job, qb;
        locate_item(p, item, job, qb);
        if constexpr (kPair) qb = 2 * qb + static_cast<int>(rank);
        const PairJob J = p.jobs[job];
        const uint32_t sa = a_it % kSlotsA;
        ptx::mbar_wait(&s.a_empty[sa], ((a_it / kSlotsA) & 1) ^ 1);
        if (ptx::elect_one()) {
          if constexpr (kPair) {
            // both CTAs load their own query block; the bytes of both are counted on the leader's barrier
            if (rank == 0) ptx::mbar_arrive_expect_tx(&s.a_full[sa], 2 * kBytesA);
            ptx::tma_load_2d_pair(s.a[sa], &tmap_q, 0, J.q_row0 + qb * kBlockQ, ptx::mapa_shared(ptx::smem_u32(&s.a_full[sa]), 0));
          } else {
            ptx::mbar_arrive_expect_tx(&s.a_full[sa], kBytesA);
            ptx::tma_load_2d(s.a[sa], &tmap_q, 0, J.q_row0 + qb * kBlockQ, &s.a_full[sa]);
          }
        }
        __syncwarp();
        const int ntiles = (J.db_rows + kTileDb - 1) / kTileDb;
        for (int t = 0; t < ntiles; ++t, ++b_it) {
#if MVGCUDA_PROBE != 8
          if constexpr (kSched == 2) {
            if ((t & 1) == 0) {  // one ring slot per tile PAIR (the constants of consecutive tiles are contiguous)
              const uint32_t sc = c_it % kSlotsC;
              ptx::mbar_wait(&s.c_empty[sc], ((c_it / kSlotsC) & 1) ^ 1);
              if (ptx::elect_one()) {
                const uint32_t bytes = (t + 1 < ntiles ? 2u : 1u) * kBytesC;
                ptx::mbar_arrive_expect_tx(&s.c_full[sc], bytes);
                ptx::bulk_load_1d(s.c[sc], p.ccol + (size_t)((J.db_row0 >> 8) + t) * kTileC, bytes, &s.c_full[sc]);
              }
              __syncwarp();
              ++c_it;
            }
          } else {
            const uint32_t sc = c_it % kSlotsC;
            ptx::mbar_wait(&s.c_empty[sc], ((c_it / kSlotsC) & 1) ^ 1);
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(&s.c_full[sc], kBytesC);
              ptx::bulk_load_1d(s.c[sc], p.ccol + (size_t)((J.db_row0 >> 8) + t) * kTileC, kBytesC, &s.c_full[sc]);
            }
            __syncwarp();
            ++c_it;
          }
#endif
          const uint32_t sb = b_it % kStagesB;
          ptx::mbar_wait(&s.b_empty[sb], ((b_it / kStagesB) & 1) ^ 1);
          const int row0 = J.db_row0 + t * kTileDb;
          if (ptx::elect_one()) {
            if constexpr (kPair) {
              // this CTA stages HALF of the tile: the rows its SM feeds to the pair's MMAs (B is split along N).
              // N = 256: rank r supplies columns [128 r, 128 r + 128).  N = 128 (kSplit): MMA group h covers tile rows
              // [128 h, 128 h + 128) and rank r supplies its columns [64 r, 64 r + 64): two boxes of 64 rows.
              if (rank == 0) ptx::mbar_arrive_expect_tx(&s.b_full[sb], kBytesB);
              const uint32_t bar = ptx::mapa_shared(ptx::smem_u32(&s.b_full[sb]), 0);
              if constexpr (kSplit) {
                ptx::tma_load_2d_pair(s.b[sb], &tmap_db, 0, row0 + 64 * static_cast<int>(rank), bar);
                ptx::tma_load_2d_pair(s.b[sb] + 64 * kDim, &tmap_db, 0, row0 + kPartCols + 64 * static_cast<int>(rank), bar);
              } else {
                ptx::tma_load_2d_pair(s.b[sb], &tmap_db, 0, row0 + kPartCols * static_cast<int>(rank), bar);
              }
            } else {
              ptx::mbar_arrive_expect_tx(&s.b_full[sb], kBytesB);
              ptx::tma_load_2d(s.b[sb], &tmap_db, 0, row0, &s.b_full[sb]);
            }
          }
          __syncwarp();
        }
      }
      if constexpr (kPair) {
        // Do not leave (and let the CTA's shared memory go) while commits of the leader can still arrive here: the last
        // a_empty commit of the pair follows every other commit, wait for it.
        if (a_it > 0) ptx::mbar_wait(&s.a_empty[(a_it - 1) % kSlotsA], ((a_it - 1) / kSlotsA) & 1);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (in a pair: of the leader CTA) =====================
    // The whole warp walks the loop (uniform control flow, uniform addresses); one elected lane issues.
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_u8u8s32(Shape::kMmaM, Shape::kMmaN);
      uint32_t a_it = 0, b_it = 0, acc_it = 0;
      uint32_t eph = 0;  // schedule 2: bit `bar` = uses of that accumulator barrier so far, mod 2
      for (int item = worker; item < p.n_items; item += n_workers, ++a_it) {
        int job, qb;
        locate_item(p, item, job, qb);
        const int db_rows = p.jobs[job].db_rows;
        const uint32_t sa = a_it % kSlotsA;
        ptx::mbar_wait(&s.a_full[sa], (a_it / kSlotsA) & 1);
        const uint64_t adesc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s.a[sa]));
        const int ntiles = (db_rows + kTileDb - 1) / kTileDb;
        for (int t = 0; t < ntiles; ++t, ++b_it, ++acc_it) {
          const uint32_t sb = b_it % kStagesB;
          // schedule 2 starts every item in buffer 0 (its epilogue walks tiles in pairs with fixed buffers)
          const uint32_t buf = kSched == 2 ? static_cast<uint32_t>(t & 1) : acc_it % kAccBufs;
          ptx::mbar_wait(&s.b_full[sb], (b_it / kStagesB) & 1);
          const uint64_t bdesc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s.b[sb]));
          const uint32_t tmem_d = tmem_base + buf * kTileDb;
          constexpr int kGroups = kSplit ? 2 : 1;
#pragma unroll
          for (int h = 0; h < kGroups; ++h) {
            const uint32_t bar = kSplit ? 2 * buf + h : buf;
#if MVGCUDA_PROBE != 8
            if constexpr (kSched == 2) {
              ptx::mbar_wait(&s.acc_empty[bar], ((eph >> bar) & 1u) ^ 1u);
              eph ^= 1u << bar;
            } else {
              ptx::mbar_wait(&s.acc_empty[bar], ((acc_it / kAccBufs) & 1) ^ 1);
            }
#endif
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < kDim / 32; ++k) {  // K = 32 bytes per kind::i8 instruction
                if constexpr (kPair)  // group h reads this CTA's rows [64 h, 64 h + 64) of the stage (8 KB further)
                  ptx::mma_i8_ss_pair(tmem_d + h * kPartCols, adesc + 2 * k, bdesc + h * (64 * kDim / 16) + 2 * k, idesc, k > 0);
                else
                  ptx::mma_i8_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0);
              }
              if constexpr (kPair) {
                if (h == kGroups - 1) ptx::mma_commit_pair(&s.b_empty[sb]);  // db stage of BOTH CTAs reusable
                ptx::mma_commit_pair(&s.acc_full[bar]);                      // accumulator ready in both CTAs
              } else {
                ptx::mma_commit(&s.b_empty[sb]);    // db stage reusable once these MMAs have read it
                ptx::mma_commit(&s.acc_full[bar]);  // accumulator ready for the epilogue
              }
              if (t == ntiles - 1 && h == kGroups - 1) {  // query slot reusable
                if constexpr (kPair) ptx::mma_commit_pair(&s.a_empty[sa]); else ptx::mma_commit(&s.a_empty[sa]);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if constexpr (kSched == 2 && MVGCUDA_PROBE == 0) {
    epilogue_slices<kPair, kSplit>(s, p, tmem_base, warp, lane, rank, worker, n_workers);
  } else {
    // ===================== epilogue: kEpiParts threads per query row =====================
    // Warp (quad, half, par) owns TMEM lanes 32*quad.., columns [128*half, 128*half+128) of the tiles whose running
    // accumulator index has parity `par` (those tiles always land in TMEM buffer `par`).  Two parities x two halves:
    // per-tile fixed costs are paid once per 128 columns, and the two parity groups drift independently.
    const int quad = warp & 3;            // a warp may only touch its own TMEM lane quadrant
    const int part = (warp - kFirstEpiWarp) >> 2;
    const int half = part & 1;
    const uint32_t par = static_cast<uint32_t>(part >> 1);
    const int row = quad * 32 + lane;     // query row within the block
    const int two = p.two;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + par * kTileDb + half * kPartCols;
    uint32_t acc_it = 0, c_it = 0, item_it = 0;
    const uint32_t acc_bar = kSplit ? 2 * par + static_cast<uint32_t>(half) : par;
    // where "this buffer is drained" is reported: the barrier of the CTA that issues the MMAs
    const uint32_t acc_empty_addr = kPair ? ptx::mapa_shared(ptx::smem_u32(&s.acc_empty[acc_bar]), 0) : ptx::smem_u32(&s.acc_empty[acc_bar]);
    s.bound[0][row] = kTInit;
    s.bound[1][row] = kTInit;
    asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * kEpiParts) : "memory");
    for (int item = worker; item < p.n_items; item += n_workers, ++item_it) {
      int job, qb;
      locate_item(p, item, job, qb);
      if constexpr (kPair) qb = 2 * qb + static_cast<int>(rank);
      const PairJob J = p.jobs[job];
      const int q_local = qb * kBlockQ + row;
      const bool q_ok = q_local < J.q_rows;
#if MVGCUDA_PROBE == 8  // TMA + MMA feed ceiling: the epilogue never touches a tile
      const int ntiles = 0;
#else
      const int ntiles = (J.db_rows + kTileDb - 1) / kTileDb;
#endif
      const int qn = p.qcol[ccol_index(J.q_row0 + min(q_local, J.q_rows - 1))] >> 8;  // ||q||^2 (dist = qn + t)
      // running best two of this thread's columns in the t-domain, t = ||d||^2 - 2 q.d  (dist = ||q||^2 + t)
      int g1t = 0x7FFFFFFF, g2t = 0x7FFFFFFF, g1i = -1, g2i = -1;
      const uint32_t bound_saddr = ptx::smem_u32(&s.bound[item_it & 1][row]);
      // the other parity's slot is idle (every part left the previous item at the barrier below): reset it for the next item
      s.bound[(item_it & 1) ^ 1][row] = kTInit;
      // this group's tiles of the item: running accumulator index of parity `par`
      for (int t = static_cast<int>((par ^ acc_it) & 1u); t < ntiles; t += 2) {
        const uint32_t acc = acc_it + t, cc_it = c_it + t;
        const uint32_t sc = cc_it % kSlotsC;
        ptx::mbar_wait(&s.c_full[sc], (cc_it / kSlotsC) & 1);
        ptx::mbar_wait(&s.acc_full[acc_bar], (acc >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t cs = ptx::smem_u32(s.c[sc] + half * kPartCols);
        int l1 = 0x7FFFFFFF, l2 = 0x7FFFFFFF;
        int T = min(min(g2t, kTInit), ptx::lds32_volatile(bound_saddr));  // admit t <= T
        const uint32_t cm_saddr = ptx::smem_u32(s.c[sc] + kTileDb + half * (kPartCols / kChunk));
        if constexpr (kSched == 1 && MVGCUDA_PROBE == 0) {
        const int4 cm0 = ptx::lds128(cm_saddr);
        const int4 cm1 = ptx::lds128(cm_saddr + 16);
        // Filter first, exact step later.  The filter needs only the maximum of a chunk, so the 16 registers of a chunk
        // are dead eight max ops after they arrive and the whole 128-column slice drains in two batches of four loads.
        // Per lane one bit per chunk records "some row of this chunk may still matter" (bound as of the tile start:
        // conservative, the bound only shrinks); the OR over the warp (one REDUX) says which chunks need the exact
        // step.  Those few (3 % of the chunks) are loaded from TMEM a second time, one after the other, by ONE rolled
        // copy of the exact step -- the accumulator is only handed back after that.
        int32_t v0[16], v1[16], v2[16], v3[16];
        uint32_t hits = 0;
        ptx::tmem_ld_32x32b_x16(taddr, v0);
        ptx::tmem_ld_32x32b_x16(taddr + 16, v1);
        ptx::tmem_ld_32x32b_x16(taddr + 32, v2);
        ptx::tmem_ld_32x32b_x16(taddr + 48, v3);
        ptx::tmem_ld_wait_for(v0);
        ptx::tmem_ld_wait_for(v1);
        ptx::tmem_ld_wait_for(v2);
        ptx::tmem_ld_wait_for(v3);
        hits |= (two * max16(v0) + T >= cm0.x) ? 1u : 0u;
        hits |= (two * max16(v1) + T >= cm0.y) ? 2u : 0u;
        hits |= (two * max16(v2) + T >= cm0.z) ? 4u : 0u;
        hits |= (two * max16(v3) + T >= cm0.w) ? 8u : 0u;
        ptx::tmem_ld_32x32b_x16(taddr + 64, v0);
        ptx::tmem_ld_32x32b_x16(taddr + 80, v1);
        ptx::tmem_ld_32x32b_x16(taddr + 96, v2);
        ptx::tmem_ld_32x32b_x16(taddr + 112, v3);
        ptx::tmem_ld_wait_for(v0);
        ptx::tmem_ld_wait_for(v1);
        ptx::tmem_ld_wait_for(v2);
        ptx::tmem_ld_wait_for(v3);
        hits |= (two * max16(v0) + T >= cm1.x) ? 16u : 0u;
        hits |= (two * max16(v1) + T >= cm1.y) ? 32u : 0u;
        hits |= (two * max16(v2) + T >= cm1.z) ? 64u : 0u;
        hits |= (two * max16(v3) + T >= cm1.w) ? 128u : 0u;
        uint32_t todo = __reduce_or_sync(0xffffffffu, hits);
#pragma unroll 1
        while (todo) {
          const int k = __ffs(todo) - 1;
          todo &= todo - 1;
          ptx::tmem_ld_32x32b_x16(taddr + 16 * k, v0);
          ptx::tmem_ld_wait_for(v0);
          int g[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) g[q4] = max(__vimax3_s32(v0[4 * q4], v0[4 * q4 + 1], v0[4 * q4 + 2]), v0[4 * q4 + 3]);
          T = min(T, ptx::lds32_volatile(bound_saddr));
          epi_exact16(v0, g, cs + 64 * k, ptx::lds32(cm_saddr + 4 * k), g1t, g2t, qn, p.prune_ratio, p.prune_rho, bound_saddr,
                      two, l1, l2, T);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) { if constexpr (kPair) ptx::mbar_arrive_cluster(acc_empty_addr); else ptx::mbar_arrive(&s.acc_empty[acc_bar]); }
        } else {
        const int4 cm0 = ptx::lds128(cm_saddr);
        const int4 cm1 = ptx::lds128(cm_saddr + 16);
        // Eight chunks of 16 columns through four register sets.  tcgen05.wait::ld waits for EVERY outstanding load, so
        // each load is issued one chunk ahead of the wait that covers it: only the first wait of a tile sees the TMEM
        // latency.  The accumulator goes back to the MMA warp as soon as the last chunk is in registers.
        int32_t v0[16], v1[16], v2[16], v3[16];
#define MVG_CHUNK(v, k, cmv) epi_chunk16(v, cs + 64 * (k), cmv, g1t, g2t, qn, p.prune_ratio, p.prune_rho, bound_saddr, two, l1, l2, T)
        ptx::tmem_ld_32x32b_x16(taddr, v0);
        ptx::tmem_ld_32x32b_x16(taddr + 16, v1);
        ptx::tmem_ld_32x32b_x16(taddr + 32, v2);
        ptx::tmem_ld_wait_for(v0);
        ptx::tmem_ld_wait_for(v1);
        ptx::tmem_ld_wait_for(v2);
        ptx::tmem_ld_32x32b_x16(taddr + 48, v3);
        MVG_CHUNK(v0, 0, cm0.x);
        ptx::tmem_ld_wait_for(v3);
        ptx::tmem_ld_32x32b_x16(taddr + 64, v0);
        MVG_CHUNK(v1, 1, cm0.y);
        ptx::tmem_ld_wait_for(v0);
        ptx::tmem_ld_32x32b_x16(taddr + 80, v1);
        T = min(T, ptx::lds32_volatile(bound_saddr));
        MVG_CHUNK(v2, 2, cm0.z);
        ptx::tmem_ld_wait_for(v1);
        ptx::tmem_ld_32x32b_x16(taddr + 96, v2);
        MVG_CHUNK(v3, 3, cm0.w);
        ptx::tmem_ld_wait_for(v2);
        ptx::tmem_ld_32x32b_x16(taddr + 112, v3);
        T = min(T, ptx::lds32_volatile(bound_saddr));
        MVG_CHUNK(v0, 4, cm1.x);
        ptx::tmem_ld_wait_for(v3);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) { if constexpr (kPair) ptx::mbar_arrive_cluster(acc_empty_addr); else ptx::mbar_arrive(&s.acc_empty[acc_bar]); }
        MVG_CHUNK(v1, 5, cm1.y);
        T = min(T, ptx::lds32_volatile(bound_saddr));
        MVG_CHUNK(v2, 6, cm1.z);
        MVG_CHUNK(v3, 7, cm1.w);
#undef MVG_CHUNK
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&s.c_empty[sc]);
        // merge the tile's top-2 into the running top-2; ties keep the earlier (lower) index.  Branch-free (selects):
        // a lane-divergent branch here costs more than the dozen ALU ops.  An untouched l1/l2 (0x7FFFFFFF) decodes to
        // t = 0x7FFFFF, larger than any real t, so it can only land in a slot that a real row later replaces.
        {
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
        }
      }
#if MVGCUDA_PROBE == 8
      acc_it += (J.db_rows + kTileDb - 1) / kTileDb;
#else
      acc_it += ntiles;
      c_it += ntiles;
#endif
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
        *reinterpret_cast<int4*>(&p.out[J.out_off + q_local]) = *reinterpret_cast<const int4*>(&r);
      }
    }
  }

  ptx::tc_fence_before();
  if constexpr (kPair) ptx::cluster_sync_all(); else __syncthreads();  // pair: no CTA leaves while its peer may still signal it
  if (warp == 0) {
    if constexpr (kPair) ptx::tmem_dealloc_pair<512>(tmem_base); else ptx::tmem_dealloc<512>(tmem_base);
  }
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
__global__ void __launch_bounds__(256)
tie_fixup_kernel(const uint8_t* __restrict__ arena, const PairJob J, KnnRecord* __restrict__ knn) {
  const int q = static_cast<int>((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= J.q_rows) return;
  const uint4* qp = reinterpret_cast<const uint4*>(arena + (size_t)(J.q_row0 + q) * kDim);
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
constexpr int kCompactThreads = 256;

// Exclusive block-wide rank of `flag` among the 256 threads + block total (ordered by threadIdx).
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

// Exclusive scan of counts over the batch (single CTA; batches are a few thousand pairs).
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int* __restrict__ counts, int n,
                                                           long long base, long long* __restrict__ offsets,
                                                           long long* __restrict__ total_out) {
  __shared__ long long warp_sum[32];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = base;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i0 = 0; i0 < n; i0 += 1024) {
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
    if (warp == 0) {
      long long w = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      warp_sum[lane] = w;  // inclusive
    }
    __syncthreads();
    const long long carry = carry_s;
    const long long excl = carry + (warp ? warp_sum[warp - 1] : 0) + (x - v);
    if (i < n) offsets[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_sum[31];
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

}  // namespace mvgcuda
