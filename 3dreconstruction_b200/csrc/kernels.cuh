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

namespace mvgcuda {

constexpr int kDim = 128;      // descriptor bytes == GEMM K
constexpr int kBlockQ = 128;   // query rows per block   (MMA M, one TMEM lane per query)
constexpr int kTileDb = 256;   // db rows per tile       (MMA N, one TMEM column per db row)
constexpr int kStagesB = 4;    // db tile ring (TMA -> MMA)
constexpr int kSlotsA = 2;     // query block double buffer
constexpr int kSlotsC = 8;     // per-column constant ring (TMA -> epilogue), outlives the B stage
constexpr int kAccBufs = 2;    // TMEM accumulator double buffer (2 x 256 columns = all 512)
constexpr int kRowAlign = 256; // every image starts at a multiple of this in the arena
constexpr int kPadNorm = 0x7FFFFF;  // "norm" of padding rows: > 128*255^2, so they never win

constexpr uint32_t kBytesA = kBlockQ * kDim;        // 16 KB
constexpr uint32_t kBytesB = kTileDb * kDim;        // 32 KB
constexpr uint32_t kBytesC = kTileDb * sizeof(int); // 1 KB

constexpr int kNumEpiWarps = 4;
constexpr int kKnnThreads = 128 + 32 * kNumEpiWarps;  // warps 0..3: TMA, MMA, TMEM-alloc, spare; 4..7 epilogue

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

struct KnnSmem {
  alignas(1024) uint8_t a[kSlotsA][kBytesA];
  alignas(1024) uint8_t b[kStagesB][kBytesB];
  alignas(16) int c[kSlotsC][kTileDb];
  uint64_t a_full[kSlotsA], a_empty[kSlotsA];
  uint64_t b_full[kStagesB], b_empty[kStagesB];
  uint64_t c_full[kSlotsC], c_empty[kSlotsC];
  uint64_t acc_full[kAccBufs], acc_empty[kAccBufs];
  uint32_t tmem_base;
};

struct KnnParams {
  const int* __restrict__ ccol;        // K1 output, [arena_rows]
  const PairJob* __restrict__ jobs;    // [n_jobs]
  const int* __restrict__ item_start;  // [n_jobs+1] prefix sum of query blocks per job
  int n_jobs;
  int n_items;
  KnnRecord* __restrict__ out;
};

// ------------------------------------------------------------------------------------------ K1
// 8 threads per 128-byte row (one 16-B load each), __dp4a squares, 3 shuffles.
__global__ void row_consts_kernel(const uint8_t* __restrict__ arena, const int* __restrict__ img_row0,
                                  const int* __restrict__ img_rows, int n_images, int arena_rows,
                                  int* __restrict__ ccol) {
  const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int row = static_cast<int>(gtid >> 3);
  const int part = static_cast<int>(gtid & 7);
  if (row >= arena_rows) return;  // whole 8-lane groups leave together (arena_rows*8 % 8 == 0)
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
    ccol[row] = (norm << 8) | (row & 255);
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

// Streaming exact top-2 of packed keys ((||d||^2 - 2 q.d) << 8 | column): 3 min/max per element.
__device__ __forceinline__ void top2_insert(int& l1, int& l2, int p) {
  l2 = min(l2, max(l1, p));
  l1 = min(l1, p);
}

__device__ __forceinline__ void epi_chunk(const int32_t (&v)[32], const int* __restrict__ cs, int& l1, int& l2) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const int4 cc = *reinterpret_cast<const int4*>(cs + j);  // warp-uniform address: smem broadcast
    top2_insert(l1, l2, static_cast<int>(static_cast<uint32_t>(cc.x) - 512u * static_cast<uint32_t>(v[j + 0])));
    top2_insert(l1, l2, static_cast<int>(static_cast<uint32_t>(cc.y) - 512u * static_cast<uint32_t>(v[j + 1])));
    top2_insert(l1, l2, static_cast<int>(static_cast<uint32_t>(cc.z) - 512u * static_cast<uint32_t>(v[j + 2])));
    top2_insert(l1, l2, static_cast<int>(static_cast<uint32_t>(cc.w) - 512u * static_cast<uint32_t>(v[j + 3])));
  }
}

__global__ void __launch_bounds__(kKnnThreads, 1)
knn2_kernel(const __grid_constant__ CUtensorMap tmap_q,   // box 128 rows x 128 B
            const __grid_constant__ CUtensorMap tmap_db,  // box 256 rows x 128 B
            const KnnParams p) {
  extern __shared__ uint8_t smem_raw[];
  KnnSmem& s = *reinterpret_cast<KnnSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_db);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kSlotsA; ++i) { ptx::mbar_init(&s.a_full[i], 1); ptx::mbar_init(&s.a_empty[i], 1); }
    for (int i = 0; i < kStagesB; ++i) { ptx::mbar_init(&s.b_full[i], 1); ptx::mbar_init(&s.b_empty[i], 1); }
    for (int i = 0; i < kSlotsC; ++i) { ptx::mbar_init(&s.c_full[i], 1); ptx::mbar_init(&s.c_empty[i], kNumEpiWarps); }
    for (int i = 0; i < kAccBufs; ++i) { ptx::mbar_init(&s.acc_full[i], 1); ptx::mbar_init(&s.acc_empty[i], kNumEpiWarps); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<512>(&s.tmem_base);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0) {
      uint32_t a_it = 0, b_it = 0, c_it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++a_it) {
        int job, qb;
        locate_item(p, item, job, qb);
        const PairJob J = p.jobs[job];
        const uint32_t sa = a_it % kSlotsA;
        ptx::mbar_wait(&s.a_empty[sa], ((a_it / kSlotsA) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&s.a_full[sa], kBytesA);
        ptx::tma_load_2d(s.a[sa], &tmap_q, 0, J.q_row0 + qb * kBlockQ, &s.a_full[sa]);
        const int ntiles = (J.db_rows + kTileDb - 1) / kTileDb;
        for (int t = 0; t < ntiles; ++t, ++b_it, ++c_it) {
          const uint32_t sc = c_it % kSlotsC;
          ptx::mbar_wait(&s.c_empty[sc], ((c_it / kSlotsC) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&s.c_full[sc], kBytesC);
          ptx::bulk_load_1d(s.c[sc], p.ccol + J.db_row0 + t * kTileDb, kBytesC, &s.c_full[sc]);
          const uint32_t sb = b_it % kStagesB;
          ptx::mbar_wait(&s.b_empty[sb], ((b_it / kStagesB) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&s.b_full[sb], kBytesB);
          ptx::tma_load_2d(s.b[sb], &tmap_db, 0, J.db_row0 + t * kTileDb, &s.b_full[sb]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_u8u8s32(kBlockQ, kTileDb);
      uint32_t a_it = 0, b_it = 0, acc_it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++a_it) {
        int job, qb;
        locate_item(p, item, job, qb);
        const int db_rows = p.jobs[job].db_rows;
        const uint32_t sa = a_it % kSlotsA;
        ptx::mbar_wait(&s.a_full[sa], (a_it / kSlotsA) & 1);
        const uint64_t adesc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s.a[sa]));
        const int ntiles = (db_rows + kTileDb - 1) / kTileDb;
        for (int t = 0; t < ntiles; ++t, ++b_it, ++acc_it) {
          const uint32_t sb = b_it % kStagesB;
          const uint32_t buf = acc_it % kAccBufs;
          ptx::mbar_wait(&s.b_full[sb], (b_it / kStagesB) & 1);
          ptx::mbar_wait(&s.acc_empty[buf], ((acc_it / kAccBufs) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint64_t bdesc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s.b[sb]));
          const uint32_t tmem_d = tmem_base + buf * kTileDb;
#pragma unroll
          for (int k = 0; k < kDim / 32; ++k)  // K = 32 bytes per kind::i8 instruction
            ptx::mma_i8_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0);
          ptx::mma_commit(&s.b_empty[sb]);    // db stage reusable once these MMAs have read it
          ptx::mma_commit(&s.acc_full[buf]);  // accumulator ready for the epilogue
        }
        ptx::mma_commit(&s.a_empty[sa]);  // query slot reusable
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: one thread per query row =====================
    const int quad = warp & 3;  // TMEM lanes 32*quad .. 32*quad+31
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    uint32_t acc_it = 0, c_it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      int job, qb;
      locate_item(p, item, job, qb);
      const PairJob J = p.jobs[job];
      const int q_local = qb * kBlockQ + quad * 32 + lane;
      const bool q_ok = q_local < J.q_rows;
      const int ntiles = (J.db_rows + kTileDb - 1) / kTileDb;
      // global best two in the t-domain, t = ||d||^2 - 2 q.d  (dist = ||q||^2 + t)
      int g1t = 0x7FFFFFFF, g2t = 0x7FFFFFFF, g1i = -1, g2i = -1;
      for (int t = 0; t < ntiles; ++t, ++acc_it, ++c_it) {
        const uint32_t buf = acc_it % kAccBufs;
        const uint32_t sc = c_it % kSlotsC;
        ptx::mbar_wait(&s.c_full[sc], (c_it / kSlotsC) & 1);
        ptx::mbar_wait(&s.acc_full[buf], (acc_it / kAccBufs) & 1);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + lane_sel + buf * kTileDb;
        const int* cs = s.c[sc];
        int l1 = 0x7FFFFFFF, l2 = 0x7FFFFFFF;
        int32_t va[32], vb[32];
        ptx::tmem_ld_32x32b_x32(taddr, va);
#pragma unroll
        for (int c = 0; c < kTileDb / 32; c += 2) {
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32b_x32(taddr + (c + 1) * 32, vb);
          epi_chunk(va, cs + c * 32, l1, l2);
          ptx::tmem_ld_wait();
          if (c + 2 < kTileDb / 32) {
            ptx::tmem_ld_32x32b_x32(taddr + (c + 2) * 32, va);
          } else {
            // every column of this accumulator is in registers: hand the buffer back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&s.acc_empty[buf]);
          }
          epi_chunk(vb, cs + (c + 1) * 32, l1, l2);
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&s.c_empty[sc]);
        // merge the tile's top-2 into the running top-2; ties keep the earlier (lower) index
        const int base = t * kTileDb;
        const int t1 = l1 >> 8, i1 = base + (l1 & 255);
        const int t2 = l2 >> 8, i2 = base + (l2 & 255);
        if (t1 < g1t) {
          if (t2 < g1t) { g2t = t2; g2i = i2; } else { g2t = g1t; g2i = g1i; }
          g1t = t1; g1i = i1;
        } else if (t1 < g2t) {
          g2t = t1; g2i = i1;
        }
      }
      if (q_ok) {
        const int qn = p.ccol[J.q_row0 + q_local] >> 8;
        KnnRecord r;
        r.idx1 = g1i; r.idx2 = g2i; r.d1 = qn + g1t; r.d2 = qn + g2t;
        *reinterpret_cast<int4*>(&p.out[J.out_off + q_local]) = *reinterpret_cast<const int4*>(&r);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
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
        pass = __int2float_rn(r.z) < __fmul_rn(ratio_sq, __int2float_rn(r.w));
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
