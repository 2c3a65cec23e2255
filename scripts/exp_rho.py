"""Kernel time of a pair-level pass by admission factor rho (mvgcuda_set_tuning): smaller rho = tighter bound for failing
queries = fewer slices leave the fast path, but more ambiguous records for the exact second pass (developer probe)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dreconstruction_b200")
rows = int(os.environ.get("ROWS", "10000")); n_img = int(os.environ.get("NIMG", "40"))
ctx = pkg.Context(0)
ctx.upload_images(pkg.synth.collection(3, n_img, rows))
pairs = pkg.pairs_exhaustive(n_img)
rs = float(pkg.square_f32(0.8))
for rho in (0.64, 0.68, 0.72, 0.76, 0.8, 0.85, 0.9, 1.0):
    ctx.set_tuning(rho, 0)
    best, resc = 1e30, 0
    for rep in range(3):
        pm = ctx.match_pairs(pairs, rs, collect=False)
        best = min(best, pm.knn_kernel_ms); resc = pm.rescanned_queries
    ops = 2.0 * rows * rows * 128 * len(pairs)
    print(f"rho {rho:.2f}: knn {best:.2f} ms (incl. second pass) -> {ops / (best * 1e-3) / 1e12:.0f} TOP/s, {resc} of {len(pairs) * rows} queries matched twice", flush=True)

# real SIFT (tests/golden): how many queries the second pass takes by rho, ratio 0.8 and 0.6
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for name, keys in (("data/et", [f"desc_{k}" for k in range(9)]), ("imageData", None)):
    z = np.load(os.path.join(ROOT, "tests", "golden", "et_collection.npz" if name == "data/et" else "imagedata_collection.npz"))
    ks = keys or ["sceaux_desc_0", "sceaux_desc_1"]
    descs = [z[k] for k in ks]
    ctx.upload_images(descs)
    prs = pkg.pairs_exhaustive(len(descs))
    nq = sum(len(descs[j]) for _, j in prs)
    for ratio in (0.8, 0.6):
        line = []
        for rho in (0.36, 0.5, 0.64, 0.72, 0.8, 0.9):
            ctx.set_tuning(rho, 0)
            pm = ctx.match_pairs(prs, float(pkg.square_f32(ratio)), collect=False)
            line.append(f"rho {rho:.2f}: {100.0 * pm.rescanned_queries / nq:.1f}%")
        print(f"{name} ({len(descs)} images, {nq} queries) ratio {ratio}: second pass " + ", ".join(line), flush=True)
