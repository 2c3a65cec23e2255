"""Golden files for the GPU geometric filter beyond data/et: a seeded synthetic collection with PLANTED epipolar geometry
(3-D points seen by cameras of different image sizes, pixel noise, outliers), pure-noise pairs (no model: all 4096
iterations, the reserve logic), long inlier lists (> 256: the global-scratch sort), duplicated correspondences (ties broken
by index), tiny pairs (<= 7 matches: nothing drawn from rand(); 8..17: below the 2.5 x 7 floor) and an empty pair --
filtered by the reference's OWN ImageCollectionGeometricFilter + GeometricFilter_FMatrix_AC
(oracle/_ref/libmvgref_geom.so, built by oracle/build_ref.sh from /root/reference), glibc rand() stream = srand(1).
A second collection whose points lie on one slanted plane goes through GeometricFilter_HMatrix_AC the same way.

    python tests/golden/make_golden_geometric_synth.py   ->  tests/golden/geo_synth.npz, geo_synth_putative.txt,
                                                             geo_synth_matches_f.txt, geo_synth_golden.json,
                                                             geo_synth_h.npz, geo_synth_h_putative.txt,
                                                             geo_synth_matches_h.txt, geo_synth_h_golden.json
"""
import ctypes as C
import hashlib
import json
import os
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GOLD = os.path.join(ROOT, "tests", "golden")
SIZES = [(4000, 3000), (1416, 1064), (640, 480), (3001, 2003), (4000, 3000), (1920, 1080), (800, 600)]
N_IMG = len(SIZES)
N_PTS = 900


def g6(a):
    return np.array([float("%g" % v) for v in np.asarray(a, np.float32).ravel()], np.float32).reshape(np.shape(a))


def collection(seed=2024, planar=False):
    rng = np.random.default_rng(seed)
    X = np.stack([rng.uniform(-2.2, 2.2, N_PTS), rng.uniform(-1.6, 1.6, N_PTS), rng.uniform(6, 14, N_PTS)], 1)
    if planar:  # one slanted plane: every pair of views is related by a homography
        X[:, 2] = 9.0 + 0.8 * X[:, 0] - 0.5 * X[:, 1]
    feats = []
    for k, (w, h) in enumerate(SIZES):
        th = 0.08 * (k - 3)
        R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
        t = np.array([0.35 * (k - 3), 0.08 * k, 0.1 * k])
        Xc = X @ R.T + t
        f = 0.9 * w
        xy = np.stack([f * Xc[:, 0] / Xc[:, 2] + w / 2, f * Xc[:, 1] / Xc[:, 2] + h / 2], 1) + rng.normal(0, 0.6, (N_PTS, 2))
        xy = np.clip(xy, 0, [w - 1, h - 1])
        extra = np.stack([rng.uniform(0, w, 300), rng.uniform(0, h, 300)], 1)    # features that match nothing real
        feats.append(g6(np.concatenate([xy, extra])))
    return feats, rng


def putatives(rng):
    out = {}
    kinds = {}
    pairs = [(i, j) for i in range(N_IMG) for j in range(i + 1, N_IMG)]
    for p, (i, j) in enumerate(pairs):
        kind = ["geo", "geo_long", "noise", "geo_weak", "tiny", "small", "dups", "empty", "dups_few", "noise_small"][p % 10]
        n_tot = N_PTS + 300
        if kind == "empty":
            m = np.zeros((0, 2), np.int64)
        elif kind == "tiny":
            m = np.stack([rng.permutation(N_PTS)[:5 + p % 3]] * 2, 1)
        elif kind == "small":
            idx = rng.permutation(N_PTS)[:8 + p % 9]
            m = np.stack([idx, idx], 1)
        else:
            n_in = {"geo": 120, "geo_long": 600, "noise": 0, "geo_weak": 40, "dups": 90, "dups_few": 110, "noise_small": 0}[kind]
            n_out = {"geo": 80, "geo_long": 150, "noise": 260, "geo_weak": 160, "dups": 60, "dups_few": 70, "noise_small": 30}[kind]
            idx = rng.permutation(N_PTS)[:n_in]
            a = np.concatenate([idx, rng.integers(0, n_tot, n_out)])
            b = np.concatenate([idx, rng.integers(0, n_tot, n_out)])
            if kind == "dups":  # the same correspondence several times (identical residuals: ties go to the lower index)
                a = np.concatenate([a, a[:25]]); b = np.concatenate([b, b[:25]])
            if kind == "dups_few":
                a = np.concatenate([a, a[:2]]); b = np.concatenate([b, b[:2]])
            m = np.stack([a, b], 1)
            m = m[np.argsort(m[:, 1], kind="stable")]   # putative lists are in ascending _j like the matcher's
        out[(i, j)] = m.astype(np.int64)
        kinds[f"{i},{j}"] = kind
    return out, kinds


def generate(model, stem, feats, rng):
    import importlib, sys
    sys.path.insert(0, ROOT)
    io = importlib.import_module("3dreconstruction_b200.io")
    put, kinds = putatives(rng)
    text = io.matches_to_text(put)
    putative = os.path.join(GOLD, stem + "_putative.txt")
    open(putative, "w").write(text)
    np.savez_compressed(os.path.join(GOLD, stem + ".npz"), sizes=np.array(SIZES, np.int32), **{f"feat_{k}": f for k, f in enumerate(feats)})
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmvgref_geom.so"))
    lib.ref_geometric_filter.restype = C.c_int
    lib.ref_geometric_filter.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.c_char_p, C.c_char, C.c_double, C.c_uint, C.c_char_p]
    with tempfile.TemporaryDirectory() as td:
        names = [f"im{k}.jpg" for k in range(N_IMG)]
        for k, f in enumerate(feats):
            io.save_feats(os.path.join(td, f"im{k}.feat"), np.concatenate([f, np.ones((len(f), 2), np.float32)], 1))
        sizes = (C.c_int * (2 * N_IMG))(*[v for wh in SIZES for v in wh])
        out = os.path.join(GOLD, ("geo_synth_matches_%s.txt" % model))
        n = lib.ref_geometric_filter(td.encode(), "\n".join(names).encode(), sizes, putative.encode(), model.encode(), 4.0, 1, out.encode())
    data = open(out, "rb").read()
    got = io.matches_from_text(data.decode())
    meta = {"seed": 1, "max_residual": 4.0, "iterations": 4096, "pairs_kept": n, "matches": int(sum(len(v) for v in got.values())),
            "sha256": hashlib.sha256(data).hexdigest(), "kinds": kinds,
            "kept": {f"{i},{j}": len(v) for (i, j), v in sorted(got.items())}}
    json.dump(meta, open(os.path.join(GOLD, stem + "_golden.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1))


def main():
    feats, rng = collection()
    generate("f", "geo_synth", feats, rng)                 # 3-D scene, GeometricFilter_FMatrix_AC
    feats, rng = collection(seed=2025, planar=True)
    generate("h", "geo_synth_h", feats, rng)               # planar scene, GeometricFilter_HMatrix_AC


if __name__ == "__main__":
    main()
