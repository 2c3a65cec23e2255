"""Golden hashes of BASELINE.json's full-size (or largest affordable) configurations, computed by the CPU oracle (L1, checked
against the reference's own code in tests/test_oracle.py).  TEST INFRASTRUCTURE: run once where CPU time is cheap (about an hour
on 8 cores); the GPU tests (tests/test_gpu_full_size.py) recompute the same hashes from the CUDA path's output.

    python tests/golden/make_golden_full.py [config3] [config4] [config5]

Hash of a configuration = sha256 over, for every pair in list order, int32 (i, j, count) followed by the count x 2 int32
(_i, _j) of the pair; `pairs` = rows 7-12 of the path (match_pairs), `collection` = rows 7-13 (with the coordinate de-dup).
"""
import hashlib
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

pkg = importlib.import_module("3dreconstruction_b200")
synth = pkg.synth
OUT = os.path.join(ROOT, "tests", "golden", "full_size_golden.json")

# name -> (synth config id, images, rows per image)
CONFIGS = {
    "config3": (3, 100, 10000),   # BASELINE configs[2] at FULL size: 4,950 pairs
    "config4": (4, 100, 8000),    # BASELINE configs[3] shape, 100-image slice of the 1,000: 4,950 pairs
    "config5": (5, 16, 40000),    # BASELINE configs[4] shape, 16-image slice of the 200: 120 pairs
}
RATIO = 0.8


def hash_update(h, i, j, m):
    m = np.ascontiguousarray(m, np.int32).reshape(-1, 2)
    h.update(np.array([i, j, len(m)], np.int32).tobytes())
    h.update(m.tobytes())


def run(name):
    cfg, n_img, rows = CONFIGS[name]
    l1 = oracle.L1()
    descs = synth.collection(cfg, n_img, rows)
    feats = [synth.features(cfg, k, rows)[:, :2].copy() for k in range(n_img)]
    pairs = pkg.pairs_exhaustive(n_img)
    rs = float(pkg.square_f32(RATIO))
    hp, hc = hashlib.sha256(), hashlib.sha256()
    tot_p = tot_c = 0
    t0 = time.time()
    for p, (i, j) in enumerate(pairs):
        m = l1.pair_matches(descs[i], descs[j], rs)
        hash_update(hp, i, j, m)
        tot_p += len(m)
        c = l1.dedup_xy(m, feats[i], feats[j])
        hash_update(hc, i, j, c)
        tot_c += len(c)
        if p % 200 == 0:
            print(f"[{name}] pair {p}/{len(pairs)}  {time.time() - t0:.0f} s", flush=True)
    return {"synth_config": cfg, "n_images": n_img, "rows": rows, "n_pairs": int(len(pairs)), "ratio": RATIO,
            "pairs_sha256": hp.hexdigest(), "pairs_matches": int(tot_p),
            "collection_sha256": hc.hexdigest(), "collection_matches": int(tot_c),
            "oracle": "L1 (oracle/oracle_l1.cpp)", "seconds": round(time.time() - t0, 1)}


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for n in names:
        res[n] = run(n)
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)
        print(n, res[n], flush=True)
