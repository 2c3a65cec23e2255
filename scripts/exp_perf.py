"""Throughput only (no exactness check) -- used with MVGCUDA_LIB=<probe build> while tuning."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dreconstruction_b200")
synth = pkg.synth
rows = int(os.environ.get("ROWS", "10000"))
n_img = int(os.environ.get("NIMG", "40"))
reps = int(os.environ.get("REPS", "3"))
ctx = pkg.Context(0)
descs = synth.collection(3, n_img, rows)
ctx.upload_images(descs)
pairs = pkg.pairs_exhaustive(n_img)
rs = float(pkg.square_f32(0.8))
best = 1e30
for rep in range(reps):
    pm = ctx.match_pairs(pairs, rs, collect=False)
    best = min(best, pm.knn_kernel_ms)
ops = 2.0 * rows * rows * 128 * len(pairs)
print(f"[perf {os.environ.get('MVGCUDA_LIB', 'product')}] {len(pairs)} pairs x {rows}: knn {best:.2f} ms -> {len(pairs) / (best * 1e-3):.0f} pairs/s, "
      f"{ops / (best * 1e-3) / 1e12:.1f} Top/s", flush=True)
