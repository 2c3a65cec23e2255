"""Probe: how long does the scalar (one thread per case) 7-point solve take on the device?  Wall clock of
mvgcuda_geo_selftest (64 threads per block) for growing n -- the slope is the throughput, the intercept the latency."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dreconstruction_b200")
ctx = pkg.Context(0)
rng = np.random.default_rng(1)
for n in [64, 3687, 3687, 4 * 3687, 16 * 3687, 64 * 3687]:
    x1 = rng.uniform(-0.6, 0.6, (n, 14)); x2 = rng.uniform(-0.6, 0.6, (n, 14)); pr = rng.uniform(-0.6, 0.6, (n, 4))
    ctx.geo_selftest(x1[:64], x2[:64], pr[:64])
    t0 = time.time(); ctx.geo_selftest(x1, x2, pr); dt = time.time() - t0
    print(f"n {n}: {dt*1e3:.2f} ms", flush=True)
for n in [3000, 12000, 48000]:
    x1 = rng.uniform(-0.6, 0.6, (n, 8)); x2 = rng.uniform(-0.6, 0.6, (n, 8)); pr = rng.uniform(-0.6, 0.6, (n, 4))
    ctx.geo_selftest_h(x1[:64], x2[:64], pr[:64])
    t0 = time.time(); ctx.geo_selftest_h(x1, x2, pr); dt = time.time() - t0
    print(f"H warp-per-case n {n}: {dt*1e3:.2f} ms", flush=True)
