"""Developer check of the GPU geometric filter: (1) device bits of the solver core against the reference's own code
(oracle/_ref/libmvgref_geom.so), (2) data/et per-pair comparison with the golden matches.f.txt."""
import ctypes as C, importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dreconstruction_b200")
io = pkg.io
GOLD = os.path.join(ROOT, "tests", "golden")
ctx = pkg.Context(0)
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmvgref_geom.so"))
dp = C.POINTER(C.c_double)
ref.ref_seven_point.restype = C.c_int
ref.ref_seven_point.argtypes = [dp, dp, dp]
ref.ref_epipolar_error.restype = C.c_double
ref.ref_epipolar_error.argtypes = [dp] + [C.c_double] * 4
rng = np.random.default_rng(3)
n = 20000
x1 = rng.uniform(-0.6, 0.6, (n, 14)); x2 = rng.uniform(-0.6, 0.6, (n, 14)); probe = rng.uniform(-0.6, 0.6, (n, 4))
x2[1::4] = x1[1::4] + 0.01 * rng.uniform(-1, 1, (n // 4, 14))
F, nm, err, nfa = ctx.geo_selftest(x1, x2, probe)
cnt_diff = coef_diff = coef_tot = err_diff = 0
worst = 0.0
for t in range(n):
    Fw = np.zeros(27)
    k = ref.ref_seven_point(x1[t].ctypes.data_as(dp), x2[t].ctypes.data_as(dp), Fw.ctypes.data_as(dp))
    if k != nm[t]:
        cnt_diff += 1
        continue
    a, b = Fw[:9 * k], F[t, :9 * k]
    coef_tot += 9 * k
    d = int((a.view(np.int64) != b.view(np.int64)).sum())
    coef_diff += d
    if d:
        worst = max(worst, float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-300))))
    if k:
        e = ref.ref_epipolar_error(F[t, :9].ctypes.data_as(dp), *[float(v) for v in probe[t]])
        err_diff += np.float64(e).view(np.int64) != np.float64(err[t]).view(np.int64)
print(f"solver core on the device vs the reference: {n} samples, model-count differences {cnt_diff}, coefficients {coef_tot}, "
      f"differing in bits {coef_diff} (worst relative {worst:.2e}), residual bit differences {err_diff}")
# one-root vs three-root split of the differences
one = nm == 1
print(f"  1-root samples {int(one.sum())}, 3-root samples {int((nm == 3).sum())}")

z = np.load(os.path.join(GOLD, "et_collection.npz"))
descs = [z[f"desc_{k}"] for k in range(9)]; feats = [z[f"feat_{k}"] for k in range(9)]
ctx.upload_images(descs); ctx.set_features([f[:, :2] for f in feats])
put = pkg.PairMatches.from_dict(io.matches_from_text(open(os.path.join(GOLD, "et_putative_r0.6.txt")).read()))
res = ctx.geometric_filter(put, [(640, 480)] * 9)
want = io.matches_from_text(open(os.path.join(GOLD, "et_matches_f.txt")).read())
print("rand consumed", res.timing["rand_consumed"], "gpu_ms", res.timing["gpu_ms"], "exact re-evaluations", res.timing["knn_kernel_launches"], "launches", res.timing["total_launches"])
for p, (i, j) in enumerate(res.pairs):
    g = res.pair(p); w = want.get((int(i), int(j)), np.zeros((0, 2), np.int64))
    same = len(g) == len(w) and np.array_equal(g, w)
    if not same:
        sg = set(map(tuple, g.tolist())); sw = set(map(tuple, w.tolist()))
        print(f"pair ({i},{j}) n_put {put.counts[p]}: got {len(g)} want {len(w)}; only-got {sorted(sg - sw)[:5]} only-want {sorted(sw - sg)[:5]}; same set {sg == sw}")
print("done")
