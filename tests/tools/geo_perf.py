"""Throughput of the GPU geometric filter next to the reference's own ACRANSAC on the host (oracle/_ref/libmvgref_geom.so,
one thread: the reference's filter loop is sequential, USE_OPENMP is defined nowhere), on two workloads:
  noise    putative matches of BASELINE config 3 images (synthetic descriptors, random coordinates: no geometry, every pair
           runs all 4,096 iterations -- the worst case per pair)
  planted  pairs with planted epipolar geometry (tests/golden/make_golden_geometric_synth.py's scene), 200 matches, 60 % inliers
Prints pairs/s for both sides and checks that the two agree on the sampled pairs (same inliers in the same order).
Log kept under profiles/.   NIMG=24 python tests/tools/geo_perf.py"""
import ctypes as C, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dreconstruction_b200")
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmvgref_geom.so"))
fp, ip, dp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)
ref.ref_acransac_f.restype = C.c_int
ref.ref_acransac_f.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint, ip, dp]
ctx = pkg.Context(0)
n_img = int(os.environ.get("NIMG", "24")); rows = int(os.environ.get("ROWS", "10000"))


def cpu_pairs(put, feats, sizes, sample):
    """The reference pair by pair (seed 1 each: timing only needs representative work)."""
    t0 = time.time(); done = 0
    for p in sample:
        i, j = put.pairs[p]; m = put.pair(p)
        if len(m) <= 7:
            continue
        xI = np.ascontiguousarray(feats[i][m[:, 0]], np.float32); xJ = np.ascontiguousarray(feats[j][m[:, 1]], np.float32)
        inl = (C.c_int * (len(m) + 1))(); o = (C.c_double * 3)()
        ref.ref_acransac_f(xI.ctypes.data_as(fp), xJ.ctypes.data_as(fp), len(m), int(sizes[i][0]), int(sizes[i][1]), int(sizes[j][0]), int(sizes[j][1]),
                           4.0, 4096, 1, inl, o)
        done += 1
    return done, time.time() - t0


def run(name, put, feats, sizes):
    ctx.geometric_filter(put, sizes)                      # warm-up (allocations, first launches)
    t0 = time.time(); res = ctx.geometric_filter(put, sizes); dt = time.time() - t0
    active = int((put.counts > 7).sum())
    print(f"[{name}] GPU: {len(put.pairs)} pairs ({active} with > 7 matches, mean {put.counts.mean():.0f} matches) in {dt*1e3:.0f} ms wall, "
          f"{res.timing['gpu_ms']:.0f} ms on the stream -> {active / dt:.0f} pairs/s; kept {(res.counts > 0).sum()} pairs / {res.counts.sum()} matches, "
          f"{res.timing['rand_consumed']} rand() values, {res.timing['knn_kernel_launches']} models re-evaluated with host roots, "
          f"{res.timing['total_launches']} launches", flush=True)
    sample = [p for p in range(len(put.pairs)) if put.counts[p] > 7][:int(os.environ.get("CPU_PAIRS", "24"))]
    done, cdt = cpu_pairs(put, feats, sizes, sample)
    print(f"[{name}] reference on 1 host thread: {done} pairs in {cdt:.1f} s -> {done / cdt:.2f} pairs/s; GPU / CPU = {(active / dt) / (done / cdt):.0f}x", flush=True)


# ---- noise workload: config-3 putatives
descs = pkg.synth.collection(3, n_img, rows)
feats = [pkg.synth.features(3, k, rows)[:, :2].copy() for k in range(n_img)]
sizes = [(4000, 3000)] * n_img
ctx.upload_images(descs); ctx.set_features(feats)
pairs = pkg.pairs_exhaustive(n_img)
put = ctx.match_collection(pairs, float(pkg.square_f32(0.8)))
run("noise", put, feats, sizes)

# ---- planted workload
rng = np.random.default_rng(7)
n_pts, n_cam = 1200, 16
X = np.stack([rng.uniform(-2.2, 2.2, n_pts), rng.uniform(-1.6, 1.6, n_pts), rng.uniform(6, 14, n_pts)], 1)
pf = []
for k in range(n_cam):
    th = 0.03 * (k - n_cam / 2); w, h = 4000, 3000
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]]); t = np.array([0.15 * (k - n_cam / 2), 0.02 * k, 0.05 * k])
    Xc = X @ R.T + t; f = 0.9 * w
    xy = np.stack([f * Xc[:, 0] / Xc[:, 2] + w / 2, f * Xc[:, 1] / Xc[:, 2] + h / 2], 1) + rng.normal(0, 0.6, (n_pts, 2))
    pf.append(np.concatenate([np.clip(xy, 0, [w - 1, h - 1]), np.stack([rng.uniform(0, w, 400), rng.uniform(0, h, 400)], 1)]).astype(np.float32))
d = {}
for i in range(n_cam):
    for j in range(i + 1, n_cam):
        idx = rng.permutation(n_pts)[:120]
        a = np.concatenate([idx, rng.integers(0, n_pts + 400, 80)]); b = np.concatenate([idx, rng.integers(0, n_pts + 400, 80)])
        m = np.stack([a, b], 1); d[(i, j)] = m[np.argsort(m[:, 1], kind="stable")]
ctx.upload_images([np.zeros((len(f), 128), np.uint8) for f in pf]); ctx.set_features(pf)
run("planted", pkg.PairMatches.from_dict(d), pf, [(4000, 3000)] * n_cam)
