"""Randomised parity soak (GPU): random sizes / alphabets / duplicates / ratios, array level (both tie modes) and pair
level, all bit-exact against the L1 oracle.  Log kept under profiles/."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("3dreconstruction_b200")
from oracle import oracle
l1 = oracle.L1()
ctx = pkg.Context(0)
rng = np.random.default_rng(int(os.environ.get("SEED", "12345")))
n_cases = int(os.environ.get("CASES", "150"))
bad = 0
t0 = time.time()
for case in range(n_cases):
    kind = rng.integers(0, 5)
    n_db = int(rng.integers(2, 3000)); n_q = int(rng.integers(1, 2000))
    if case % 25 == 24:
        n_db = int(rng.integers(3000, 12000))   # many db tiles: long scans, all four column/tile parts busy
    if kind == 0:
        alpha = int(rng.choice([2, 3, 4, 8, 16]))
        db = rng.integers(0, alpha, (n_db, 128), dtype=np.uint8); q = rng.integers(0, alpha, (n_q, 128), dtype=np.uint8)
    elif kind == 1:
        db = rng.integers(0, 256, (n_db, 128), dtype=np.uint8); q = rng.integers(0, 256, (n_q, 128), dtype=np.uint8)
    elif kind == 2:   # sparse, many zero rows / duplicates
        db = (rng.random((n_db, 128)) < 0.05).astype(np.uint8) * rng.integers(1, 256, (n_db, 128), dtype=np.uint8)
        q = (rng.random((n_q, 128)) < 0.05).astype(np.uint8) * rng.integers(1, 256, (n_q, 128), dtype=np.uint8)
        db[rng.integers(0, n_db, n_db // 4)] = db[rng.integers(0, n_db, n_db // 4)]
    elif kind == 4:   # clusters: every query has several near copies in the db at different noise levels, scattered over
        # the scan order, so d1/d2 sits on both sides of the ratio threshold (exercises the ratio-aware pruning rule)
        base = pkg.synth.image(700 + case, 0, max(2, n_db // 4), np.zeros((0, 128), np.uint8), shared=0.0)
        src = rng.integers(0, len(base), n_db)
        amp = rng.integers(0, 40, (n_db, 1))
        db = np.clip(base[src].astype(np.int16) + rng.integers(-1, 2, (n_db, 128)) * amp, 0, 255).astype(np.uint8)
        qs = rng.integers(0, len(base), n_q)
        q = np.clip(base[qs].astype(np.int16) + rng.integers(-1, 2, (n_q, 128)) * rng.integers(0, 12, (n_q, 1)), 0, 255).astype(np.uint8)
    else:             # SIFT-like with queries that are noisy copies of db rows
        db = pkg.synth.image(900 + case, 0, n_db, np.zeros((0, 128), np.uint8), shared=0.0)
        src = rng.integers(0, n_db, n_q)
        q = np.clip(db[src].astype(np.int16) + rng.integers(-4, 5, (n_q, 128)), 0, 255).astype(np.uint8)
    ok = True
    for tie in (0, 1):
        want = l1.knn2(db, q, tie)
        idx, dist = ctx.knn2_arrays(db, q, tie)
        ok &= np.array_equal(idx, want[0]) and np.array_equal(dist.astype(np.int32), want[1])
    ctx.upload_images([db, q])
    for r in (0.6, 0.8, float(rng.uniform(0.3, 1.3))):
        rs = float(pkg.square_f32(r))
        res = ctx.match_pairs(np.array([[0, 1], [1, 0]], np.int32), rs)
        ok &= np.array_equal(res.pair(0), l1.pair_matches(db, q, rs)) and np.array_equal(res.pair(1), l1.pair_matches(q, db, rs) if n_q >= 2 else np.zeros((0, 2), np.int32))
    if not ok:
        bad += 1
        print(f"case {case} kind {kind} db {n_db} q {n_q}: MISMATCH", flush=True)
print(f"{n_cases} randomised cases, {bad} mismatches, {time.time() - t0:.1f} s")
print("SOAK OK" if bad == 0 else "SOAK FAILED")
