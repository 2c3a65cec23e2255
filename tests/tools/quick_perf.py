"""Quick GPU check used while tuning the kernel: exactness on a few shapes vs the L1 oracle + kernel throughput."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("3dreconstruction_b200")
from oracle import oracle
synth = pkg.synth
l1 = oracle.L1()
ctx = pkg.Context(0)
ok = True
cases = [("uniform", synth.uniform_set(5, 1000), synth.uniform_set(6, 777)),
         ("tie3", synth.tie_set(7, 1500, 3), synth.tie_set(8, 900, 3)),
         ("sift10k",) + tuple(synth.collection(2, 2, 10000)),
         ("uniform10k", synth.uniform_set(41, 10000), synth.uniform_set(42, 4000))]
for name, db, q in cases:
    for tie in (0, 1):
        want = l1.knn2(db, q, tie)
        idx, dist = ctx.knn2_arrays(db, q, tie)
        good = np.array_equal(idx, want[0]) and np.array_equal(dist.astype(np.int32), want[1])
        print(f"[{name} tie={tie}] exact={good}", flush=True)
        if not good:
            bad = np.argwhere((idx != want[0]) | (dist.astype(np.int32) != want[1]))
            print("   bad rows", len(np.unique(bad[:, 0])), "first", bad[0], idx[bad[0, 0]], want[0][bad[0, 0]], dist[bad[0, 0]], want[1][bad[0, 0]])
        ok &= good
rows = int(os.environ.get("ROWS", "10000"))
n_img = int(os.environ.get("NIMG", "40"))
descs = synth.collection(3, n_img, rows)
ctx.upload_images(descs)
pairs = pkg.pairs_exhaustive(n_img)
rs = float(pkg.square_f32(0.8))
for rep in range(3):
    pm = ctx.match_pairs(pairs, rs, collect=False)
ops = 2.0 * rows * rows * 128 * len(pairs)
print(f"[perf] {len(pairs)} pairs x {rows}: knn {pm.knn_kernel_ms:.2f} ms -> {len(pairs) / (pm.knn_kernel_ms * 1e-3):.0f} pairs/s, "
      f"{ops / (pm.knn_kernel_ms * 1e-3) / 1e12:.1f} Top/s; gpu total {pm.gpu_ms:.2f} ms", flush=True)
print("ALL OK" if ok else "FAILURES")
