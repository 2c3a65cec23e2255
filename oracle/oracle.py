"""ctypes access to the two CPU oracles.  TEST INFRASTRUCTURE ONLY.

  L1  oracle/_build/liboracle_l1.so  <- oracle/oracle_l1.cpp   independent restatement (always available; built
                                                                 on demand with g++)
  L0  oracle/_ref/libmvgref[_omp].so <- oracle/ref_driver.cpp   the reference's own headers, compiled by
                                                                 oracle/build_ref.sh where /root/reference exists;
                                                                 the built .so travels to the GPU box

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg, --impl reference) may import this
module.  The product (3dreconstruction_b200/) never does, and has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
L1_PATH = os.path.join(_HERE, "_build", "liboracle_l1.so")
L0_PATH = os.path.join(_HERE, "_ref", "libmvgref.so")
L0_OMP_PATH = os.path.join(_HERE, "_ref", "libmvgref_omp.so")
REF_ROOT = os.environ.get("MVG_REF", "/root/reference")

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
DIM = 128


def build_l1(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle_l1.cpp")
    if force or not os.path.exists(L1_PATH) or os.path.getmtime(L1_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s", "_build/liboracle_l1.so"])
    return L1_PATH


def build_l0() -> Optional[str]:
    """Build L0 if the reference tree is visible; returns the path or None."""
    if os.path.isdir(os.path.join(REF_ROOT, "libs", "feature", "include")):
        drv = os.path.join(_HERE, "ref_driver.cpp")
        if not os.path.exists(L0_PATH) or os.path.getmtime(L0_PATH) < os.path.getmtime(drv):
            subprocess.check_call(["bash", os.path.join(_HERE, "build_ref.sh")])
    return L0_PATH if os.path.exists(L0_PATH) else None


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, DIM)


def _ptr_array(mats: Sequence[np.ndarray], ctype):
    arr = (C.POINTER(ctype) * max(len(mats), 1))()
    for k, m in enumerate(mats):
        arr[k] = m.ctypes.data_as(C.POINTER(ctype)) if m is not None and m.size else None
    return arr


class L1:
    """Independent restatement (oracle_l1.cpp)."""

    def __init__(self):
        self.lib = C.CDLL(build_l1())
        L = self.lib
        L.l1_sqdist.restype = C.c_int
        L.l1_sqdist.argtypes = [_u8p, _u8p, C.c_int]
        L.l1_knn2.restype = C.c_int
        L.l1_knn2.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_int, _i32p, _i32p]
        L.l1_ratio_pass.restype = C.c_int
        L.l1_ratio_pass.argtypes = [C.c_int, C.c_int, C.c_float]
        L.l1_pair_matches.restype = C.c_int
        L.l1_pair_matches.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_float, _i32p]
        L.l1_dedup_indexed_sorted.restype = C.c_int
        L.l1_dedup_indexed_sorted.argtypes = [_i32p, C.c_int]
        L.l1_dedup_xy.restype = C.c_int
        L.l1_dedup_xy.argtypes = [_i32p, C.c_int, _f32p, _f32p]
        L.l1_match_collection.restype = C.c_longlong
        L.l1_match_collection.argtypes = [C.POINTER(_u8p), C.POINTER(_f32p), _i32p, _i32p, C.c_int, C.c_float, _i32p, _i32p]
        L.l1_export_text.restype = C.c_int
        L.l1_export_text.argtypes = [_i32p, _i32p, _i32p, C.c_int, C.c_char_p]
        L.l1_bench_bf.restype = C.c_longlong
        L.l1_bench_bf.argtypes = [C.POINTER(_u8p), _i32p, C.POINTER(_u8p), _i32p, C.c_int, C.c_float]
        L.l1_num_threads.restype = C.c_int

    def sqdist(self, a, b) -> int:
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        return self.lib.l1_sqdist(a.ctypes.data_as(_u8p), b.ctypes.data_as(_u8p), a.size)

    def knn2(self, db, q, tie_mode: int = 1) -> Optional[Tuple[np.ndarray, np.ndarray]]:
        db, q = _u8(db), _u8(q)
        idx = np.zeros((max(len(q), 1), 2), np.int32)
        dist = np.zeros((max(len(q), 1), 2), np.int32)
        ok = self.lib.l1_knn2(db.ctypes.data_as(_u8p), len(db), q.ctypes.data_as(_u8p), len(q), tie_mode,
                              idx.ctypes.data_as(_i32p), dist.ctypes.data_as(_i32p))
        return (idx[:len(q)], dist[:len(q)]) if ok else None

    def ratio_pass(self, d1: int, d2: int, ratio_sq: float) -> bool:
        return bool(self.lib.l1_ratio_pass(int(d1), int(d2), C.c_float(ratio_sq)))

    def pair_matches(self, db, q, ratio_sq: float) -> np.ndarray:
        db, q = _u8(db), _u8(q)
        out = np.zeros((max(len(q), 1), 2), np.int32)
        n = self.lib.l1_pair_matches(db.ctypes.data_as(_u8p), len(db), q.ctypes.data_as(_u8p), len(q), C.c_float(ratio_sq),
                                     out.ctypes.data_as(_i32p))
        return out[:n].copy()

    def dedup_indexed_sorted(self, m) -> np.ndarray:
        m = np.ascontiguousarray(m, np.int32).reshape(-1, 2).copy()
        n = self.lib.l1_dedup_indexed_sorted(m.ctypes.data_as(_i32p), len(m))
        return m[:n]

    def dedup_xy(self, m, fI, fJ) -> np.ndarray:
        m = np.ascontiguousarray(m, np.int32).reshape(-1, 2).copy()
        fI = np.ascontiguousarray(fI, np.float32).reshape(-1, 2)
        fJ = np.ascontiguousarray(fJ, np.float32).reshape(-1, 2)
        n = self.lib.l1_dedup_xy(m.ctypes.data_as(_i32p), len(m), fI.ctypes.data_as(_f32p), fJ.ctypes.data_as(_f32p))
        return m[:n]

    def match_collection(self, descs, feats, pairs, ratio_sq: float) -> Dict[Tuple[int, int], np.ndarray]:
        descs = [_u8(d) if len(d) else np.zeros((0, DIM), np.uint8) for d in descs]
        rows = np.array([len(d) for d in descs], np.int32)
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        dptr = _ptr_array(descs, C.c_uint8)
        if feats is not None:
            feats = [np.ascontiguousarray(f, np.float32).reshape(-1, 2) for f in feats]
            fptr = _ptr_array(feats, C.c_float)
        else:
            fptr = None
        cap = int(sum(rows[j] for _, j in pairs)) + 1
        counts = np.zeros(max(len(pairs), 1), np.int32)
        out = np.zeros((cap, 2), np.int32)
        self.lib.l1_match_collection(dptr, fptr, rows.ctypes.data_as(_i32p), pairs.ctypes.data_as(_i32p), len(pairs),
                                     C.c_float(ratio_sq), counts.ctypes.data_as(_i32p), out.ctypes.data_as(_i32p))
        res, off = {}, 0
        for p, (i, j) in enumerate(pairs):
            res.setdefault((int(i), int(j)), out[off:off + counts[p]].copy())
            off += counts[p]
        return res

    def export_text(self, pairwise: Dict[Tuple[int, int], np.ndarray], path: str) -> None:
        keys = sorted(pairwise)
        pairs = np.array(keys, np.int32).reshape(-1, 2)
        counts = np.array([len(pairwise[k]) for k in keys], np.int32)
        allm = [np.asarray(pairwise[k], np.int32).reshape(-1, 2) for k in keys]
        m = np.concatenate(allm) if allm else np.zeros((0, 2), np.int32)
        m = np.ascontiguousarray(m if len(m) else np.zeros((1, 2), np.int32))
        ok = self.lib.l1_export_text(pairs.ctypes.data_as(_i32p), counts.ctypes.data_as(_i32p), m.ctypes.data_as(_i32p), len(keys),
                                     path.encode())
        if not ok:
            raise IOError(path)

    def bench_bf(self, dbs, qs, ratio_sq: float) -> int:
        dbs = [_u8(d) for d in dbs]
        qs = [_u8(d) for d in qs]
        dr = np.array([len(d) for d in dbs], np.int32)
        qr = np.array([len(d) for d in qs], np.int32)
        return self.lib.l1_bench_bf(_ptr_array(dbs, C.c_uint8), dr.ctypes.data_as(_i32p), _ptr_array(qs, C.c_uint8),
                                    qr.ctypes.data_as(_i32p), len(dbs), C.c_float(ratio_sq))

    def num_threads(self) -> int:
        return self.lib.l1_num_threads()


class L0:
    """The reference's own code (ref_driver.cpp over the reference headers)."""

    def __init__(self, openmp: bool = False):
        path = L0_OMP_PATH if openmp else L0_PATH
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: build with oracle/build_ref.sh where the reference tree is mounted")
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_metric.restype = C.c_float
        L.ref_metric.argtypes = [_u8p, _u8p, C.c_int]
        L.ref_square.restype = C.c_float
        L.ref_square.argtypes = [C.c_float]
        L.ref_knn.restype = C.c_int
        L.ref_knn.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_int, _i32p, _f32p]
        L.ref_ratio_filter.restype = C.c_int
        L.ref_ratio_filter.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, _i32p]
        L.ref_pair_matches.restype = C.c_int
        L.ref_pair_matches.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_float, _i32p]
        L.ref_dedup_indexed.restype = C.c_int
        L.ref_dedup_indexed.argtypes = [_i32p, C.c_int]
        L.ref_dedup_xy.restype = C.c_int
        L.ref_dedup_xy.argtypes = [_i32p, C.c_int, _f32p, C.c_int, _f32p, C.c_int]
        L.ref_match_dir.restype = C.c_int
        L.ref_match_dir.argtypes = [C.c_char_p, C.c_char_p, C.c_float, C.c_char_p]
        L.ref_roundtrip_matches.restype = C.c_int
        L.ref_roundtrip_matches.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_bench_bf.restype = C.c_longlong
        L.ref_bench_bf.argtypes = [C.POINTER(_u8p), _i32p, C.POINTER(_u8p), _i32p, C.c_int, C.c_float]
        L.ref_has_flann.restype = C.c_int
        L.ref_has_openmp.restype = C.c_int
        if L.ref_has_flann():
            L.ref_bench_flann.restype = C.c_longlong
            L.ref_bench_flann.argtypes = L.ref_bench_bf.argtypes

    def metric(self, a, b) -> float:
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        return float(self.lib.ref_metric(a.ctypes.data_as(_u8p), b.ctypes.data_as(_u8p), a.size))

    def square(self, r: float) -> np.float32:
        return np.float32(self.lib.ref_square(C.c_float(r)))

    def knn(self, db, q, k: int = 2) -> Optional[Tuple[np.ndarray, np.ndarray]]:
        db, q = _u8(db), _u8(q)
        idx = np.zeros((max(len(q), 1), k), np.int32)
        dist = np.zeros((max(len(q), 1), k), np.float32)
        ok = self.lib.ref_knn(db.ctypes.data_as(_u8p), len(db), q.ctypes.data_as(_u8p), len(q), k, idx.ctypes.data_as(_i32p),
                              dist.ctypes.data_as(_f32p))
        return (idx[:len(q)], dist[:len(q)]) if ok else None

    def ratio_filter(self, dist, ratio_sq: float, nn: int = 2) -> np.ndarray:
        d = np.ascontiguousarray(dist, np.float32).reshape(-1)
        out = np.zeros(max(d.size // nn, 1), np.int32)
        n = self.lib.ref_ratio_filter(d.ctypes.data_as(_f32p), d.size, nn, C.c_float(ratio_sq), out.ctypes.data_as(_i32p))
        return out[:n].copy()

    def pair_matches(self, db, q, ratio_sq: float) -> np.ndarray:
        db, q = _u8(db), _u8(q)
        out = np.zeros((max(len(q), 1), 2), np.int32)
        n = self.lib.ref_pair_matches(db.ctypes.data_as(_u8p), len(db), q.ctypes.data_as(_u8p), len(q), C.c_float(ratio_sq),
                                      out.ctypes.data_as(_i32p))
        return out[:n].copy()

    def dedup_indexed(self, m) -> np.ndarray:
        m = np.ascontiguousarray(m, np.int32).reshape(-1, 2).copy()
        n = self.lib.ref_dedup_indexed(m.ctypes.data_as(_i32p), len(m)) if len(m) else 0
        return m[:n]

    def dedup_xy(self, m, fI, fJ) -> np.ndarray:
        m = np.ascontiguousarray(m, np.int32).reshape(-1, 2).copy()
        fI = np.ascontiguousarray(fI, np.float32).reshape(-1, 2)
        fJ = np.ascontiguousarray(fJ, np.float32).reshape(-1, 2)
        n = self.lib.ref_dedup_xy(m.ctypes.data_as(_i32p), len(m), fI.ctypes.data_as(_f32p), len(fI), fJ.ctypes.data_as(_f32p), len(fJ))
        return m[:n]

    def match_dir(self, match_dir: str, names: Sequence[str], dist_ratio: float, out_path: str) -> None:
        ok = self.lib.ref_match_dir(match_dir.encode(), "\n".join(names).encode(), C.c_float(dist_ratio), out_path.encode())
        if not ok:
            raise RuntimeError("reference MatcherAllInMemory failed")

    def match_collection_text(self, descs, feats4, dist_ratio: float) -> bytes:
        """Write .desc (8-byte count) / .feat files to a temp dir, run the reference collection matcher, return
        the bytes of its matches.putative.txt."""
        import importlib
        io = importlib.import_module("3dreconstruction_b200.io")
        with tempfile.TemporaryDirectory() as td:
            names = []
            for k, (d, f) in enumerate(zip(descs, feats4)):
                name = f"img{k:05d}"
                names.append(name + ".jpg")
                io.save_descs_bin(os.path.join(td, name + ".desc"), d, 8)
                io.save_feats(os.path.join(td, name + ".feat"), f)
            out = os.path.join(td, "matches.putative.txt")
            self.match_dir(td, names, dist_ratio, out)
            with open(out, "rb") as fh:
                return fh.read()

    def roundtrip_matches(self, in_path: str, out_path: str) -> None:
        if not self.lib.ref_roundtrip_matches(in_path.encode(), out_path.encode()):
            raise RuntimeError("reference pairedIndexedMatchImport failed")

    def bench(self, dbs, qs, ratio_sq: float, flann: bool = False) -> int:
        dbs = [_u8(d) for d in dbs]
        qs = [_u8(d) for d in qs]
        dr = np.array([len(d) for d in dbs], np.int32)
        qr = np.array([len(d) for d in qs], np.int32)
        fn = self.lib.ref_bench_flann if flann else self.lib.ref_bench_bf
        return fn(_ptr_array(dbs, C.c_uint8), dr.ctypes.data_as(_i32p), _ptr_array(qs, C.c_uint8), qr.ctypes.data_as(_i32p),
                  len(dbs), C.c_float(ratio_sq))


def have_l0() -> bool:
    return os.path.exists(L0_PATH)
