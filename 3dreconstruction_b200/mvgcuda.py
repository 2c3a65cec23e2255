"""ctypes binding of libmvgcuda (include/mvgcuda.h) plus the Python mirror of the reference's matcher
interfaces for this path.

Mirrors (reference file:line, relative to the reference tree):
  * ``ArrayMatcherCuda``       <-> ArrayMatcher<uchar,Metric>   libs/feature/include/mvg/feature/matching_interface.h:16-64
                                    (BF implementation: matcher_brute_force.h:42-50,102-134)
  * ``MatcherCudaAllInMemory`` <-> MatcherAllInMemory            matcher_all_in_memory.h:19-147
  * ``square_f32``             <-> Square(float)                 libs/base/include/mvg/math/numeric.h:108-111

There is NO CPU fallback here: if the shared library is missing, or no sm_100 device is present,
every compute call raises ``MvgCudaError``.  (The CPU oracle lives under ``oracle/`` and is test
infrastructure only.)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

DIM = 128
TIE_LOWEST_INDEX = 0
TIE_REFERENCE = 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MVGCUDA_LIB", os.path.join(_HERE, "libmvgcuda.so"))  # override: developer probes


class MvgCudaError(RuntimeError):
    pass


class _PairMatches(C.Structure):
    _fields_ = [
        ("n_pairs", C.c_int64),
        ("counts", C.POINTER(C.c_int32)),
        ("offsets", C.POINTER(C.c_int64)),
        ("matches", C.POINTER(C.c_int32)),
        ("gpu_ms", C.c_float),
        ("knn_kernel_ms", C.c_float),
        ("knn_kernel_launches", C.c_int32),
        ("total_launches", C.c_int32),
        ("rescanned_queries", C.c_int64),
    ]


class _DeviceInfo(C.Structure):
    _fields_ = [
        ("name", C.c_char * 128),
        ("sm_count", C.c_int),
        ("cc_major", C.c_int),
        ("cc_minor", C.c_int),
        ("clock_khz", C.c_int),
        ("hbm_bytes", C.c_int64),
    ]


# every symbol include/mvgcuda.h declares: name -> (restype, argtypes)
_u8pp = C.POINTER(C.POINTER(C.c_uint8))
_f32pp = C.POINTER(C.POINTER(C.c_float))
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_ctx = C.c_void_p
ABI: Dict[str, Tuple[object, list]] = {
    "mvgcuda_version": (C.c_int, []),
    "mvgcuda_device_count": (C.c_int, []),
    "mvgcuda_device_ordinal": (C.c_int, [C.c_int]),
    "mvgcuda_create": (C.c_int, [C.c_int, C.POINTER(_ctx)]),
    "mvgcuda_destroy": (None, [_ctx]),
    "mvgcuda_last_error": (C.c_char_p, [_ctx]),
    "mvgcuda_set_stream": (C.c_int, [_ctx, C.c_void_p]),
    "mvgcuda_set_tuning": (C.c_int, [_ctx, C.c_float, C.c_int]),
    "mvgcuda_upload_images": (C.c_int, [_ctx, C.c_int, _u8pp, _i32p, C.c_int]),
    "mvgcuda_clone_images": (C.c_int, [_ctx, _ctx]),
    "mvgcuda_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "mvgcuda_host_free": (None, [C.c_void_p]),
    "mvgcuda_stream_begin": (C.c_int, [_ctx, C.c_int, _i32p]),
    "mvgcuda_stream_image": (C.c_int, [_ctx, C.c_int, C.POINTER(C.c_uint8), _f32p]),
    "mvgcuda_stream_images": (C.c_int, [_ctx, C.c_int, _i32p, _u8pp, _f32pp]),
    "mvgcuda_stream_end": (C.c_int, [_ctx]),
    "mvgcuda_db_create": (C.c_int, [_ctx, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_void_p)]),
    "mvgcuda_db_destroy": (None, [_ctx, C.c_void_p]),
    "mvgcuda_db_rows": (C.c_int, [C.c_void_p]),
    "mvgcuda_db_knn2": (C.c_int, [_ctx, C.c_void_p, C.POINTER(C.c_uint8), C.c_int, C.c_int, _i32p, _f32p]),
    "mvgcuda_num_images": (C.c_int, [_ctx]),
    "mvgcuda_image_rows": (C.c_int, [_ctx, C.c_int]),
    "mvgcuda_knn2": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, _i32p, _f32p]),
    "mvgcuda_knn2_arrays": (C.c_int, [_ctx, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int, _i32p, _f32p]),
    "mvgcuda_match_pairs": (C.c_int, [_ctx, C.c_int64, _i32p, C.c_float, C.POINTER(_PairMatches)]),
    "mvgcuda_set_features": (C.c_int, [_ctx, C.c_int, _f32pp, _i32p]),
    "mvgcuda_match_collection": (C.c_int, [_ctx, C.c_int64, _i32p, C.c_float, C.c_int, C.POINTER(_PairMatches)]),
    "mvgcuda_export_matches": (C.c_int, [_ctx, _i32p, C.c_char_p]),
    "mvgcuda_geometric_filter": (C.c_int, [_ctx, C.c_char, C.c_double, C.c_int, C.c_uint, C.c_int64, _i32p, _i32p, C.POINTER(C.c_int64), _i32p,
                                           _i32p, C.POINTER(_PairMatches)]),
    "mvgcuda_geo_selftest": (C.c_int, [_ctx, C.c_int] + [C.POINTER(C.c_double)] * 4 + [_i32p] + [C.POINTER(C.c_double)] * 2),
    "mvgcuda_geo_selftest_h": (C.c_int, [_ctx, C.c_int] + [C.POINTER(C.c_double)] * 5),
    "mvgcuda_write_matches": (C.c_int, [C.c_char_p, C.c_int64, _i32p, _i32p, C.POINTER(C.c_int64), _i32p, C.c_int]),
    "mvgcuda_get_device_info": (C.c_int, [_ctx, C.POINTER(_DeviceInfo)]),
    "mvgcuda_probe_i8_peak": (C.c_int, [_ctx, C.c_int, C.POINTER(C.c_double), _f32p]),
}

_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """dlopen libmvgcuda.so and bind every ABI symbol.  Raises MvgCudaError if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise MvgCudaError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def square_f32(r: float) -> np.float32:
    """Square(float) of the reference (numeric.h:108-111): the product is rounded in fp32.
    0.8 -> 0x3f23d70b, which is NOT float32(0.64)."""
    f = np.float32(r)
    return np.float32(f * f)


def pairs_exhaustive(n_images: int) -> np.ndarray:
    """(i, j), i < j, in the order of MatcherAllInMemory::Match's double loop (matcher_all_in_memory.h:71-90)."""
    i, j = np.triu_indices(n_images, k=1)
    return np.stack([i, j], axis=1).astype(np.int32)


def upload_friendly_order(pairs: np.ndarray, head_images: Optional[int] = None) -> np.ndarray:
    """The same pairs in an order that lets matching overlap a streamed upload AND keeps the fused kernel fed:
      * head: the pairs among the first `head_images` images (default: a third of them, at least 8), by their LARGER image
        id first -- with images streamed in index order these batches only touch images that have already arrived;
      * tail: everything else in (i, j) order, as the reference visits them.  Consecutive pairs then share their db image
        I, whose tiles stay in L2; in larger-id-first order the db image changes with every pair and every pair starts
        with a cold fetch of its first db tiles (measured on a resident collection of 100 x 10k: 45.2 instead of 42.4 ms
        in the fused kernel).  By the time the head has been matched the rest of the upload has arrived (end to end, 100
        images x 10k: head 16 / 25 / 34 / 50 / all = 105 / 106 / 110 / 108 / 104k pairs/s).
    Results are keyed by pair, so the order is free (the export sorts by (i, j) itself)."""
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    if len(pairs) == 0:
        return pairs
    hi, lo = pairs.max(axis=1), pairs.min(axis=1)
    n = int(hi.max()) + 1
    m = int(head_images) if head_images is not None else max(8, (n + 2) // 3)
    head = hi < m
    hp, tp = pairs[head], pairs[~head]
    hp = hp[np.lexsort((lo[head], hi[head]))]
    tp = tp[np.lexsort((tp[:, 1], tp[:, 0]))]
    return np.ascontiguousarray(np.concatenate([hp, tp]))


def write_matches(path: str, res: "PairMatches", skip_empty: bool = False) -> None:
    """PairedIndexedMatchToStream (indexed_match_utils.h:22-38) of a result whose pairs are in (i, j) order; skip_empty drops
    pairs without matches, as the reference's geometric map never receives them (geometric_filter.h:85-98)."""
    lib = load_library()
    pairs = np.ascontiguousarray(res.pairs, dtype=np.int32).reshape(-1, 2)
    counts = np.ascontiguousarray(res.counts, dtype=np.int32)
    offsets = np.ascontiguousarray(res.offsets, dtype=np.int64)
    matches = np.ascontiguousarray(res.matches, dtype=np.int32).reshape(-1, 2)
    if len(matches) == 0:
        matches = np.zeros((1, 2), np.int32)
    rc = lib.mvgcuda_write_matches(path.encode(), len(pairs), pairs.ctypes.data_as(_i32p), counts.ctypes.data_as(_i32p),
                                   offsets.ctypes.data_as(C.POINTER(C.c_int64)), matches.ctypes.data_as(_i32p), int(skip_empty))
    if rc != 0:
        raise MvgCudaError(f"mvgcuda_write_matches({path}) failed [{rc}]")


def _as_u8_matrix(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim != 2 or a.shape[1] != DIM:
        raise ValueError(f"descriptors must be [rows][{DIM}] uint8, got {a.shape}")
    return a


class PairMatches:
    """Result of a match call (copied out of the context's buffers)."""

    def __init__(self, pairs: np.ndarray, counts: np.ndarray, offsets: np.ndarray, matches: np.ndarray, timing: dict):
        self.pairs, self.counts, self.offsets, self.matches, self.timing = pairs, counts, offsets, matches, timing

    def __len__(self) -> int:
        return len(self.counts)

    def pair(self, p: int) -> np.ndarray:
        """[count][2] (_i, _j) of pair p."""
        return self.matches[self.offsets[p]:self.offsets[p] + self.counts[p]]

    @staticmethod
    def from_dict(pairwise: Dict[Tuple[int, int], np.ndarray]) -> "PairMatches":
        """From a PairWiseMatches-like dict (e.g. io.matches_from_text of an imported matches.putative.txt), in std::map order."""
        keys = sorted(pairwise)
        pairs = np.array(keys, np.int32).reshape(-1, 2)
        counts = np.array([len(pairwise[k]) for k in keys], np.int32)
        offsets = np.zeros(len(keys) + 1, np.int64)
        np.cumsum(counts, out=offsets[1:])
        matches = (np.concatenate([np.asarray(pairwise[k], np.int32).reshape(-1, 2) for k in keys]) if keys else np.zeros((0, 2), np.int32))
        return PairMatches(pairs, counts, offsets, matches.astype(np.int32), {})

    def as_dict(self) -> Dict[Tuple[int, int], np.ndarray]:
        """PairWiseMatches (indexed_match.h:69): first insertion wins for duplicate keys, as std::map::insert."""
        out: Dict[Tuple[int, int], np.ndarray] = {}
        for p, (i, j) in enumerate(self.pairs):
            out.setdefault((int(i), int(j)), self.pair(p))
        return out


class StagedImages:
    """Argument tables of a streamed upload for a fixed set of host buffers (Context.stage_images)."""
    __slots__ = ("mats", "fm", "n", "cnt", "rows", "idx", "dps", "fps")


class Context:
    """One libmvgcuda context == one GPU."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = _ctx()
        rc = self._lib.mvgcuda_create(device, C.byref(h))
        if rc != 0:
            raise MvgCudaError(f"mvgcuda_create({device}) failed [{rc}]: {self._lib.mvgcuda_last_error(None).decode()}")
        self._h = h
        self.device = device

    # -- plumbing
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.mvgcuda_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise MvgCudaError(f"{what} failed [{rc}]: {self._lib.mvgcuda_last_error(self._h).decode()}")

    def set_stream(self, cuda_stream: int) -> None:
        self._check(self._lib.mvgcuda_set_stream(self._h, C.c_void_p(cuda_stream)), "mvgcuda_set_stream")

    def set_tuning(self, prune_rho: float = 0.72, rescan_rows: int = 0) -> None:
        """Ratio-aware pruning knobs (results never depend on them): admission factor and rescan buffer rows."""
        self._check(self._lib.mvgcuda_set_tuning(self._h, C.c_float(prune_rho), int(rescan_rows)), "mvgcuda_set_tuning")

    def device_info(self) -> dict:
        di = _DeviceInfo()
        self._check(self._lib.mvgcuda_get_device_info(self._h, C.byref(di)), "mvgcuda_get_device_info")
        return {"name": di.name.decode(), "sm_count": di.sm_count, "cc": (di.cc_major, di.cc_minor),
                "clock_khz": di.clock_khz, "hbm_bytes": di.hbm_bytes}

    def probe_i8_peak(self, iters: int = 20000) -> Tuple[float, float]:
        ops = C.c_double()
        ms = C.c_float()
        self._check(self._lib.mvgcuda_probe_i8_peak(self._h, iters, C.byref(ops), C.byref(ms)), "mvgcuda_probe_i8_peak")
        return ops.value, ms.value

    # -- residency
    def upload_images(self, descs: Sequence[np.ndarray], pinned: bool = False) -> None:
        mats = [_as_u8_matrix(d) if len(d) else np.zeros((0, DIM), np.uint8) for d in descs]
        n = len(mats)
        ptrs = (C.POINTER(C.c_uint8) * max(n, 1))()
        rows = (C.c_int32 * max(n, 1))()
        for k, m in enumerate(mats):
            ptrs[k] = m.ctypes.data_as(C.POINTER(C.c_uint8)) if m.shape[0] else None
            rows[k] = m.shape[0]
        self._check(self._lib.mvgcuda_upload_images(self._h, n, ptrs, rows, int(pinned)), "mvgcuda_upload_images")

    def upload_images_device(self, device_ptrs: Sequence[int], rows: Sequence[int]) -> None:
        """Same as upload_images, from descriptor arrays that already live in device memory (raw CUDA pointers, e.g.
        ``tensor.data_ptr()`` of a [rows][128] uint8 tensor on any GPU of the box)."""
        n = len(rows)
        ptrs = (C.POINTER(C.c_uint8) * max(n, 1))()
        rws = (C.c_int32 * max(n, 1))()
        for k in range(n):
            ptrs[k] = C.cast(C.c_void_p(int(device_ptrs[k])), C.POINTER(C.c_uint8)) if rows[k] else None
            rws[k] = int(rows[k])
        self._check(self._lib.mvgcuda_upload_images(self._h, n, ptrs, rws, 0), "mvgcuda_upload_images")

    def stage_images(self, descs: Sequence[np.ndarray], feats_xy: Optional[Sequence[np.ndarray]] = None,
                     order: Optional[Sequence[int]] = None) -> "StagedImages":
        """The argument tables of a streamed upload (row counts, image order, buffer addresses), built once for a set of
        host buffers that will be sent again and again -- e.g. page-locked staging buffers a loader refills.  The handle
        keeps the arrays alive; their CONTENTS are read at every stream_images(staged=...) call."""
        mats = [_as_u8_matrix(d) if len(d) else np.zeros((0, DIM), np.uint8) for d in descs]
        n = len(mats)
        fm = [np.ascontiguousarray(f, dtype=np.float32).reshape(-1, 2) for f in feats_xy] if feats_xy is not None else None
        ids = [int(k) for k in (order if order is not None else range(n))]
        cnt = len(ids)
        st = StagedImages()
        st.mats, st.fm, st.n, st.cnt = mats, fm, n, cnt
        st.rows = (C.c_int32 * max(n, 1))(*[m.shape[0] for m in mats])
        st.idx = (C.c_int32 * max(cnt, 1))(*ids)
        # raw addresses (cheaper than a ctypes pointer object per array: this loop is host time in front of the match call)
        st.dps = (C.c_void_p * max(cnt, 1))(*[mats[k].__array_interface__["data"][0] if mats[k].shape[0] else None for k in ids])
        st.fps = None
        if fm is not None:
            st.fps = (C.c_void_p * max(cnt, 1))(*[fm[k].__array_interface__["data"][0] if mats[k].shape[0] else None for k in ids])
        return st

    def stream_images(self, descs: Optional[Sequence[np.ndarray]] = None, feats_xy: Optional[Sequence[np.ndarray]] = None,
                      order: Optional[Sequence[int]] = None, wait: bool = True, staged: Optional["StagedImages"] = None) -> None:
        """Same residency as upload_images (+ set_features), through the streaming entry points: one asynchronous copy per
        image on the context's upload stream, in `order` (default 0..n-1).  With wait=False the call returns while the
        copies are in flight -- a following match call starts on the pairs whose images have arrived -- and the arrays must
        stay untouched until stream_end() (the context keeps references to them until then).  `staged`: a handle from
        stage_images() instead of the arrays (two library calls, no per-image Python work)."""
        st = staged if staged is not None else self.stage_images(descs, feats_xy, order)
        self._check(self._lib.mvgcuda_stream_begin(self._h, st.n, st.rows), "mvgcuda_stream_begin")
        self._staged = st
        self._check(self._lib.mvgcuda_stream_images(self._h, st.cnt, st.idx, C.cast(st.dps, _u8pp),
                                                    C.cast(st.fps, _f32pp) if st.fps is not None else None), "mvgcuda_stream_images")
        if wait:
            self.stream_end()

    def stream_end(self) -> None:
        self._check(self._lib.mvgcuda_stream_end(self._h), "mvgcuda_stream_end")
        self._staged = None

    def db_create(self, db: np.ndarray) -> "ResidentDb":
        return ResidentDb(self, db)

    def clone_images_from(self, src: "Context") -> None:
        """Replica of another context's uploaded collection (and features), copied device to device."""
        self._check(self._lib.mvgcuda_clone_images(self._h, src._h), "mvgcuda_clone_images")

    def set_features(self, feats_xy: Sequence[np.ndarray]) -> None:
        mats = [np.ascontiguousarray(f, dtype=np.float32).reshape(-1, 2) for f in feats_xy]
        n = len(mats)
        ptrs = (C.POINTER(C.c_float) * max(n, 1))()
        rows = (C.c_int32 * max(n, 1))()
        for k, m in enumerate(mats):
            ptrs[k] = m.ctypes.data_as(C.POINTER(C.c_float)) if m.shape[0] else None
            rows[k] = m.shape[0]
        self._check(self._lib.mvgcuda_set_features(self._h, n, ptrs, rows), "mvgcuda_set_features")

    # -- array level
    def knn2(self, db_img: int, q_img: int, tie_mode: int = TIE_REFERENCE) -> Tuple[np.ndarray, np.ndarray]:
        nq = max(self._lib.mvgcuda_image_rows(self._h, q_img), 0)  # the library's own count: valid for clones and streamed sets too
        idx = np.empty((max(nq, 1), 2), np.int32)
        dist = np.empty((max(nq, 1), 2), np.float32)
        self._check(self._lib.mvgcuda_knn2(self._h, db_img, q_img, tie_mode, idx.ctypes.data_as(_i32p),
                                           dist.ctypes.data_as(_f32p)), "mvgcuda_knn2")
        return idx[:nq], dist[:nq]

    def knn2_arrays(self, db: np.ndarray, query: np.ndarray, tie_mode: int = TIE_REFERENCE) -> Tuple[np.ndarray, np.ndarray]:
        db, query = _as_u8_matrix(db), _as_u8_matrix(query)
        nq = query.shape[0]
        idx = np.empty((max(nq, 1), 2), np.int32)
        dist = np.empty((max(nq, 1), 2), np.float32)
        u8p = C.POINTER(C.c_uint8)
        self._check(self._lib.mvgcuda_knn2_arrays(self._h, db.ctypes.data_as(u8p), db.shape[0], query.ctypes.data_as(u8p), nq,
                                                  tie_mode, idx.ctypes.data_as(_i32p), dist.ctypes.data_as(_f32p)),
                    "mvgcuda_knn2_arrays")
        return idx[:nq], dist[:nq]

    # -- pair / collection level
    def _collect(self, pairs: np.ndarray, pm: _PairMatches) -> PairMatches:
        n = int(pm.n_pairs)
        counts = np.ctypeslib.as_array(pm.counts, shape=(max(n, 1),))[:n].copy()
        offsets = np.ctypeslib.as_array(pm.offsets, shape=(n + 1,)).copy()
        total = int(offsets[n])
        matches = (np.ctypeslib.as_array(pm.matches, shape=(max(total, 1) * 2,))[: total * 2].copy().reshape(total, 2))
        timing = {"gpu_ms": pm.gpu_ms, "knn_kernel_ms": pm.knn_kernel_ms, "knn_kernel_launches": pm.knn_kernel_launches,
                  "total_launches": pm.total_launches, "rescanned_queries": pm.rescanned_queries}
        return PairMatches(pairs, counts, offsets, matches, timing)

    def match_pairs(self, pairs: np.ndarray, ratio_sq: float, collect: bool = True):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        pm = _PairMatches()
        self._check(self._lib.mvgcuda_match_pairs(self._h, len(pairs), pairs.ctypes.data_as(_i32p), C.c_float(ratio_sq),
                                                  C.byref(pm)), "mvgcuda_match_pairs")
        return self._collect(pairs, pm) if collect else pm

    def match_collection(self, pairs: np.ndarray, ratio_sq: float, host_threads: int = 0, collect: bool = True):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        pm = _PairMatches()
        self._check(self._lib.mvgcuda_match_collection(self._h, len(pairs), pairs.ctypes.data_as(_i32p), C.c_float(ratio_sq),
                                                       host_threads, C.byref(pm)), "mvgcuda_match_collection")
        return self._collect(pairs, pm) if collect else pm

    # -- the step after the path: AC-RANSAC geometric filter (geometric_filter.h:37-102)
    def geometric_filter(self, putative: "PairMatches", image_sizes: Sequence[Tuple[int, int]], model: str = "f", precision: float = 4.0,
                         iterations: int = 4096, seed: int = 1) -> "PairMatches":
        """ImageCollectionGeometricFilter::Filter(GeometricFilter_FMatrix_AC(precision, iterations)) -- model "f" -- or
        GeometricFilter_HMatrix_AC -- model "h" -- over `putative` (pairs in std::map order; features already set).  seed=1
        is the reference's never-seeded rand() stream.  Returns the kept matches per pair in ascending-residual order;
        timing['rand_consumed'] = rand() values drawn.  The essential-matrix functor ("e") is not built: error."""
        pairs = np.ascontiguousarray(putative.pairs, dtype=np.int32).reshape(-1, 2)
        counts = np.ascontiguousarray(putative.counts, dtype=np.int32)
        offsets = np.ascontiguousarray(putative.offsets, dtype=np.int64)
        matches = np.ascontiguousarray(putative.matches, dtype=np.int32).reshape(-1, 2)
        if len(matches) == 0:
            matches = np.zeros((1, 2), np.int32)
        sizes = np.ascontiguousarray(image_sizes, dtype=np.int32).reshape(-1, 2)
        pm = _PairMatches()
        self._check(self._lib.mvgcuda_geometric_filter(self._h, model.encode()[:1], C.c_double(precision), int(iterations), C.c_uint(seed),
                                                       len(pairs), pairs.ctypes.data_as(_i32p), counts.ctypes.data_as(_i32p),
                                                       offsets.ctypes.data_as(C.POINTER(C.c_int64)), matches.ctypes.data_as(_i32p),
                                                       sizes.ctypes.data_as(_i32p), C.byref(pm)), "mvgcuda_geometric_filter")
        res = self._collect(pairs, pm)
        res.timing["rand_consumed"] = res.timing.pop("rescanned_queries")
        return res

    def geo_selftest_h(self, x1: np.ndarray, x2: np.ndarray, probe: np.ndarray):
        """Device bits of the homography path: (H [n][9], transfer error [n])."""
        x1 = np.ascontiguousarray(x1, np.float64).reshape(-1, 8)
        x2 = np.ascontiguousarray(x2, np.float64).reshape(-1, 8)
        probe = np.ascontiguousarray(probe, np.float64).reshape(-1, 4)
        n = len(x1)
        H = np.zeros((n, 9)); err = np.zeros(n)
        dp = C.POINTER(C.c_double)
        self._check(self._lib.mvgcuda_geo_selftest_h(self._h, n, x1.ctypes.data_as(dp), x2.ctypes.data_as(dp), probe.ctypes.data_as(dp),
                                                     H.ctypes.data_as(dp), err.ctypes.data_as(dp)), "mvgcuda_geo_selftest_h")
        return H, err

    def geo_selftest(self, x1: np.ndarray, x2: np.ndarray, probe: np.ndarray):
        """Device bits of the solver core: (F [n][27], n_models [n], residual [n], nfa term [n])."""
        x1 = np.ascontiguousarray(x1, np.float64).reshape(-1, 14)
        x2 = np.ascontiguousarray(x2, np.float64).reshape(-1, 14)
        probe = np.ascontiguousarray(probe, np.float64).reshape(-1, 4)
        n = len(x1)
        F = np.zeros((n, 27)); nm = np.zeros(n, np.int32); err = np.zeros(n); nfa = np.zeros(n)
        dp = C.POINTER(C.c_double)
        self._check(self._lib.mvgcuda_geo_selftest(self._h, n, x1.ctypes.data_as(dp), x2.ctypes.data_as(dp), probe.ctypes.data_as(dp),
                                                   F.ctypes.data_as(dp), nm.ctypes.data_as(_i32p), err.ctypes.data_as(dp), nfa.ctypes.data_as(dp)),
                    "mvgcuda_geo_selftest")
        return F, nm, err, nfa

    def export_matches(self, pairs: np.ndarray, path: str) -> None:
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        self._check(self._lib.mvgcuda_export_matches(self._h, pairs.ctypes.data_as(_i32p), path.encode()), "mvgcuda_export_matches")


class ResidentDb:
    """A database image that stays in HBM (mvgcuda_db_*): Build once, search many times, only the queries travel."""

    def __init__(self, ctx: Context, db: np.ndarray):
        self._ctx = ctx
        db = _as_u8_matrix(db)
        h = C.c_void_p()
        ctx._check(ctx._lib.mvgcuda_db_create(ctx._h, db.ctypes.data_as(C.POINTER(C.c_uint8)), db.shape[0], C.byref(h)), "mvgcuda_db_create")
        self._h = h
        self.rows = db.shape[0]

    def knn2(self, query: np.ndarray, tie_mode: int = TIE_REFERENCE) -> Tuple[np.ndarray, np.ndarray]:
        query = _as_u8_matrix(query)
        nq = query.shape[0]
        idx = np.empty((max(nq, 1), 2), np.int32)
        dist = np.empty((max(nq, 1), 2), np.float32)
        self._ctx._check(self._ctx._lib.mvgcuda_db_knn2(self._ctx._h, self._h, query.ctypes.data_as(C.POINTER(C.c_uint8)), nq, tie_mode,
                                                        idx.ctypes.data_as(_i32p), dist.ctypes.data_as(_f32p)), "mvgcuda_db_knn2")
        return idx[:nq], dist[:nq]

    def close(self) -> None:
        if getattr(self, "_h", None) and getattr(self._ctx, "_h", None):
            self._ctx._lib.mvgcuda_db_destroy(self._ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# Mirrors of the reference's operator interfaces (same names / argument meaning / error behaviour)


class ArrayMatcherCuda:
    """ArrayMatcher<unsigned char, SquaredEuclideanDistanceVectorized<unsigned char>> on the GPU.

    ``Build`` copies the db to HBM (the reference's BF matcher borrows the pointer,
    matcher_brute_force.h:47-48); ``SearchNeighbours`` APPENDS k (idx, dist) per query to the output
    lists like the BF matcher (``push_back``, :128-131), returns False and reports
    "Too much asked nearest neighbors" when k > rows or nq < 1 (:107-110).  Only k in {1, 2}.
    """

    def __init__(self, ctx: Optional[Context] = None, tie_mode: int = TIE_REFERENCE):
        self._ctx = ctx or Context(0)
        self._db: Optional[ResidentDb] = None
        self._rows = 0
        self._tie = tie_mode

    def Build(self, dataset: np.ndarray, rows_num: int, dimension: int = DIM) -> bool:
        if self._db is not None:
            self._db.close()
            self._db = None
        self._rows = 0
        if rows_num < 1:  # matcher_brute_force.h:43-46
            return False
        if dimension != DIM:
            raise ValueError("ArrayMatcherCuda handles 128-byte descriptors only")
        db = _as_u8_matrix(np.asarray(dataset).reshape(-1, DIM)[:rows_num])
        self._rows = db.shape[0]
        if self._rows == 1:  # k == 1 against a single row: duplicate it so the 2-NN kernel has two candidates
            db = np.concatenate([db, db], axis=0)
        self._db = self._ctx.db_create(db)  # the rows go to HBM once; searches upload only their queries
        return True

    def SearchNeighbours(self, query: np.ndarray, query_num: int, vec_indice: list, vec_distance: list,
                         nearest_neighbor_num: int = 2) -> bool:
        if nearest_neighbor_num > self._rows or query_num < 1:
            import sys
            print("Too much asked nearest neighbors", file=sys.stderr)
            return False
        if nearest_neighbor_num not in (1, 2):
            raise ValueError("ArrayMatcherCuda accelerates k = 1 or 2 only")
        q = _as_u8_matrix(np.asarray(query).reshape(-1, DIM)[:query_num])
        k = nearest_neighbor_num
        # k == 1: std::partial_sort(first, first + 1, last) keeps the FIRST minimum (strict < in __heap_select,
        # indexed_sort.h:52-66), i.e. the lowest index; only k == 2 needs the two-slot tie machine
        idx, dist = self._db.knn2(q, self._tie if k == 2 else TIE_LOWEST_INDEX)
        if self._rows == 1:
            idx = np.zeros_like(idx)
        vec_indice.extend(idx[:, :k].reshape(-1).tolist())
        vec_distance.extend(dist[:, :k].reshape(-1).tolist())
        return True

    def SearchNeighbour(self, query: np.ndarray) -> Tuple[bool, int, float]:
        """Single nearest neighbour (matcher_brute_force.h:61-89: std::min_element => lowest index on ties)."""
        if self._db is None:
            return False, -1, 0.0
        idx, dist = self._db.knn2(_as_u8_matrix(np.asarray(query).reshape(1, DIM)), TIE_LOWEST_INDEX)
        return True, (int(idx[0, 0]) if self._rows > 1 else 0), float(dist[0, 0])


class MatcherCudaAllInMemory:
    """MatcherAllInMemory<KeypointSet<ScalePointFeature, Descriptor<uchar,128>>, ArrayMatcherBruteForce<...>>.

    ``LoadData(file_names, match_dir)`` reads ``<match_dir>/<basename>.feat`` (text, feature.h:117-133) and
    ``.desc`` (binary, descriptor.h:160-181; 8-byte count as written on Linux, or the 4-byte count of the
    shipped data/et files) and makes them resident in HBM; ``Match(file_names)`` returns the PairWiseMatches
    dict {(i, j): [count][2]} for every i < j, empty pairs included (matcher_all_in_memory.h:135).
    """

    def __init__(self, distRatio: float, ctx: Optional[Context] = None, host_threads: int = 0):
        self.distance_ratio = np.float32(distRatio)
        self._ctx = ctx or Context(0)
        self._n = 0  # (host_threads: accepted and ignored -- the coordinate de-dup runs on the GPU)

    def LoadArrays(self, descs: Sequence[np.ndarray], feats_xy: Sequence[np.ndarray], wait: bool = True,
                   order: Optional[Sequence[int]] = None) -> bool:
        # row counts come from the features, as in the reference (matcher_all_in_memory.h:80,85,107)
        rows = [np.asarray(f).reshape(-1, 2).shape[0] for f in feats_xy]
        d2 = []
        for d, r in zip(descs, rows):
            d = np.asarray(d, dtype=np.uint8).reshape(-1, DIM)
            if d.shape[0] < r:
                raise MvgCudaError(".feat has more rows than .desc (the reference over-reads here; refused)")
            d2.append(d[:r])
        # one asynchronous copy per image (descriptors + coordinates).  wait=False: Match() starts while images still travel
        self._ctx.stream_images(d2, feats_xy, order=order, wait=wait)
        self._n = len(d2)
        return True

    def StageArrays(self, descs: Sequence[np.ndarray], feats_xy: Sequence[np.ndarray], order: Optional[Sequence[int]] = None) -> StagedImages:
        """LoadArrays' checks and argument tables once, for buffers that are sent repeatedly (LoadStaged)."""
        rows = [np.asarray(f).reshape(-1, 2).shape[0] for f in feats_xy]
        d2 = []
        for d, r in zip(descs, rows):
            d = np.asarray(d, dtype=np.uint8).reshape(-1, DIM)
            if d.shape[0] < r:
                raise MvgCudaError(".feat has more rows than .desc (the reference over-reads here; refused)")
            d2.append(d[:r])
        return self._ctx.stage_images(d2, feats_xy, order=order)

    def LoadStaged(self, staged: StagedImages, wait: bool = True) -> bool:
        self._ctx.stream_images(staged=staged, wait=wait)
        self._n = staged.n
        return True

    def LoadData(self, file_names: Sequence[str], match_dir: str) -> bool:
        from .io import load_descs_bin, load_feats
        descs, feats = [], []
        for name in file_names:
            base = os.path.splitext(os.path.basename(name))[0]
            feats.append(load_feats(os.path.join(match_dir, base + ".feat"))[:, :2])
            descs.append(load_descs_bin(os.path.join(match_dir, base + ".desc")))
        return self.LoadArrays(descs, feats)

    def Match(self, file_names: Optional[Sequence[str]] = None, pairs: Optional[np.ndarray] = None) -> Dict[Tuple[int, int], np.ndarray]:
        n = self._n if file_names is None else len(file_names)
        if pairs is None:
            pairs = pairs_exhaustive(n)
        pairs = upload_friendly_order(pairs)
        self.last = self._ctx.match_collection(pairs, float(square_f32(self.distance_ratio)))
        self._ctx.stream_end()  # images handed over with wait=False: their buffers are free again
        return self.last.as_dict()

    def Export(self, path: str) -> None:
        """matches.putative.txt, byte-identical to PairedIndexedMatchToStream (indexed_match_utils.h:22-38)."""
        self._ctx.export_matches(self.last.pairs, path)
