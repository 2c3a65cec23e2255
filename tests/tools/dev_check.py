"""Developer bring-up check (run on a GPU box via gpurun): numpy brute force vs libmvgcuda.
Not part of the test-suite; the real parity tests are tests/test_gpu_*.py against oracle/."""
import importlib
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("3dreconstruction_b200")
synth = pkg.synth


def brute(db, q):
    d = db.astype(np.float64)
    x = q.astype(np.float64)
    dist = (x * x).sum(1)[:, None] + (d * d).sum(1)[None, :] - 2.0 * (x @ d.T)
    return np.rint(dist).astype(np.int64)


def top2_lowest(dist):
    n = dist.shape[1]
    key = dist * n + np.arange(n)[None, :]
    part = np.sort(key, axis=1)[:, :2]
    return part % n, part // n


def machine(drow):
    T, S = (0, 1) if drow[1] < drow[0] else (1, 0)
    for v in range(2, len(drow)):
        if drow[v] < drow[T]:
            if drow[S] < drow[v]:
                T = v
            else:
                T = S
                S = v
    return S, T


def check_knn(ctx, name, db, q, tie_ref=False):
    t0 = time.time()
    idx, dist = ctx.knn2_arrays(db, q, pkg.TIE_REFERENCE if tie_ref else pkg.TIE_LOWEST_INDEX)
    t1 = time.time()
    D = brute(db, q)
    ridx, rd = top2_lowest(D)
    ok_d = np.array_equal(dist.astype(np.int64), rd)
    if tie_ref:
        ridx = np.array([machine(D[k]) for k in range(len(q))])
    ok_i = np.array_equal(idx.astype(np.int64), ridx)
    print(f"[{name}] db={db.shape[0]} q={q.shape[0]} dist_ok={ok_d} idx_ok={ok_i} ({(t1 - t0) * 1e3:.1f} ms)", flush=True)
    if not ok_d:
        bad = np.argwhere(dist.astype(np.int64) != rd)
        print("   first bad dist rows:", bad[:5].tolist(), "got", dist[bad[0, 0]], "want", rd[bad[0, 0]], "idx got",
              idx[bad[0, 0]], "want", ridx[bad[0, 0]])
        print("   n bad", len(bad), "of", dist.size)
    elif not ok_i:
        bad = np.argwhere(idx.astype(np.int64) != ridx)
        print("   first bad idx rows:", bad[:5].tolist(), idx[bad[0, 0]], ridx[bad[0, 0]], dist[bad[0, 0]])
    return ok_d and ok_i


def ref_pairs(descs, pairs, ratio_sq):
    out = []
    for i, j in pairs:
        db, q = descs[i], descs[j]
        if len(db) < 2 or len(q) < 1:
            out.append(np.zeros((0, 2), np.int64))
            continue
        D = brute(db, q)
        idx, d = top2_lowest(D)
        pas = np.nonzero(d[:, 0].astype(np.float32) < np.float32(ratio_sq) * d[:, 1].astype(np.float32))[0]
        pas = pas[:-1] if len(pas) else pas
        m = [(int(idx[k, 0]), int(k)) for k in pas]
        ded = [m[k] for k in range(len(m)) if k == 0 or m[k][0] != m[k - 1][0]]
        out.append(np.array(ded, np.int64).reshape(-1, 2))
    return out


def main():
    ok = True
    ctx = pkg.Context(0)
    print(ctx.device_info(), flush=True)
    try:
        for iters in (2000, 20000):
            ops, ms = ctx.probe_i8_peak(iters)
            print(f"i8 probe iters={iters}: {ops / 1e12:.1f} Top/s in {ms:.3f} ms", flush=True)
    except Exception:
        traceback.print_exc()
        ok = False
    rng = np.random.default_rng(1)
    cases = [
        ("tiny", synth.uniform_set(1, 2), synth.uniform_set(2, 1)),
        ("small", synth.uniform_set(3, 300), synth.uniform_set(4, 200)),
        ("ragged", synth.uniform_set(5, 1000), synth.uniform_set(6, 777)),
        ("sift", synth.collection(7, 2, 2500)[0], synth.collection(7, 2, 2500)[1]),
        ("maxval", np.full((257, 128), 255, np.uint8), np.zeros((129, 128), np.uint8)),
    ]
    for name, db, q in cases:
        try:
            ok &= check_knn(ctx, name, db, q)
        except Exception:
            traceback.print_exc()
            ok = False
    try:
        ok &= check_knn(ctx, "ties-lowest", synth.tie_set(11, 700), synth.tie_set(12, 300))
        ok &= check_knn(ctx, "ties-ref", synth.tie_set(11, 700), synth.tie_set(12, 300), tie_ref=True)
        ok &= check_knn(ctx, "ties-ref4", synth.tie_set(13, 513, 4), synth.tie_set(14, 130, 4), tie_ref=True)
    except Exception:
        traceback.print_exc()
        ok = False
    try:
        descs = synth.collection(21, 6, 1500) + [synth.uniform_set(9, 1)[:1], np.zeros((0, 128), np.uint8), synth.uniform_set(10, 333)]
        ctx.upload_images(descs)
        pairs = pkg.pairs_exhaustive(len(descs))
        for r in (0.6, 0.8):
            rs = float(pkg.square_f32(r))
            res = ctx.match_pairs(pairs, rs)
            want = ref_pairs(descs, pairs, rs)
            good = all(np.array_equal(res.pair(p).astype(np.int64), want[p]) for p in range(len(pairs)))
            print(f"[match_pairs r={r}] pairs={len(pairs)} matches={int(res.offsets[-1])} ok={good} timing={res.timing}", flush=True)
            if not good:
                for p in range(len(pairs)):
                    if not np.array_equal(res.pair(p).astype(np.int64), want[p]):
                        print("   first bad pair", pairs[p], len(res.pair(p)), len(want[p]))
                        break
            ok &= good
    except Exception:
        traceback.print_exc()
        ok = False
    try:
        n_img, rows = 24, 10000
        descs = synth.collection(3, n_img, rows)
        ctx.upload_images(descs)
        pairs = pkg.pairs_exhaustive(n_img)
        rs = float(pkg.square_f32(0.8))
        for rep in range(3):
            t0 = time.time()
            pm = ctx.match_pairs(pairs, rs, collect=False)
            t1 = time.time()
            ops = 2.0 * rows * rows * 128 * len(pairs)
            print(f"[perf] {len(pairs)} pairs x {rows}: gpu {pm.gpu_ms:.2f} ms (knn {pm.knn_kernel_ms:.2f} ms) wall {(t1 - t0) * 1e3:.1f} ms -> "
                  f"{len(pairs) / (pm.knn_kernel_ms * 1e-3):.0f} pairs/s kernel, {ops / (pm.knn_kernel_ms * 1e-3) / 1e12:.1f} Top/s", flush=True)
    except Exception:
        traceback.print_exc()
        ok = False
    print("ALL OK" if ok else "FAILURES", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
