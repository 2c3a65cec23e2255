"""Where the end-to-end time goes (developer probe): load, match_pairs, match_collection wall / device times."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dreconstruction_b200")
synth = pkg.synth
rows = int(os.environ.get("ROWS", "10000")); n_img = int(os.environ.get("NIMG", "100"))
descs = synth.collection(3, n_img, rows)
feats = [synth.features(3, k, rows)[:, :2].copy() for k in range(n_img)]
pairs = pkg.pairs_exhaustive(n_img)
rs = float(pkg.square_f32(0.8))
ctx = pkg.Context(0)
m = pkg.MatcherCudaAllInMemory(0.8, ctx)
for rep in range(3):
    t0 = time.time(); m.LoadArrays(descs, feats); t1 = time.time()
    a = ctx.match_pairs(pairs, rs, collect=False); t2 = time.time()
    a_ms = (a.gpu_ms, a.knn_kernel_ms, a.knn_kernel_launches)
    b = ctx.match_collection(pairs, rs, collect=False); t3 = time.time()
    print(f"[e2e rep {rep}] load {1e3*(t1-t0):.1f} ms | match_pairs wall {1e3*(t2-t1):.1f} ms (gpu {a_ms[0]:.1f}, knn {a_ms[1]:.1f}, {a_ms[2]} batches) | "
          f"match_collection wall {1e3*(t3-t2):.1f} ms (gpu {b.gpu_ms:.1f}, knn {b.knn_kernel_ms:.1f})", flush=True)
