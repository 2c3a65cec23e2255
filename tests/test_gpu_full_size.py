"""BASELINE.json configurations at full size (config 3: all 4,950 pairs of 100 x 10,000 rows) or the largest slice the CPU
oracle could afford (config 4: 100 of the 1,000 images; config 5: 16 of the 200), EVERY pair checked: the CUDA path's
output is hashed exactly like tests/golden/make_golden_full.py hashed the oracle's (sha256 over (i, j, count, matches) of
every pair in list order), for rows 7-12 (match_pairs) and rows 7-13 (match_collection, coordinate de-dup on the GPU)."""
import hashlib
import importlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
synth = importlib.import_module("3dreconstruction_b200.synth")
GOLD = os.path.join(GOLDEN, "full_size_golden.json")


def _hash(res):
    h = hashlib.sha256()
    for p, (i, j) in enumerate(res.pairs):
        m = np.ascontiguousarray(res.pair(p), np.int32)
        h.update(np.array([i, j, len(m)], np.int32).tobytes())
        h.update(m.tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("name", ["config3", "config4", "config5"])
def test_full_size_hashes(ctx, pkg, name):
    gold = json.load(open(GOLD)) if os.path.exists(GOLD) else {}
    if name not in gold:
        pytest.skip(f"{name}: no golden hash committed (tests/golden/make_golden_full.py)")
    g = gold[name]
    descs = synth.collection(g["synth_config"], g["n_images"], g["rows"])
    feats = [synth.features(g["synth_config"], k, g["rows"])[:, :2].copy() for k in range(g["n_images"])]
    pairs = pkg.pairs_exhaustive(g["n_images"])
    assert len(pairs) == g["n_pairs"]
    rs = float(pkg.square_f32(g["ratio"]))
    ctx.upload_images(descs)
    res = ctx.match_pairs(pairs, rs)
    assert int(res.counts.sum()) == g["pairs_matches"]
    assert _hash(res) == g["pairs_sha256"], "rows 7-12 differ from the oracle somewhere in the configuration"
    ctx.set_features(feats)
    col = ctx.match_collection(pairs, rs)
    assert int(col.counts.sum()) == g["collection_matches"]
    assert _hash(col) == g["collection_sha256"], "rows 7-13 differ from the oracle somewhere in the configuration"
    # the sharded run (what --gpus N does): two halves of the pair list, hashed in order, give the same digest
    cut = len(pairs) // 2 + 7
    h = hashlib.sha256()
    for part in (pairs[:cut], pairs[cut:]):
        r = ctx.match_collection(part, rs)
        for p, (i, j) in enumerate(r.pairs):
            m = np.ascontiguousarray(r.pair(p), np.int32)
            h.update(np.array([i, j, len(m)], np.int32).tobytes())
            h.update(m.tobytes())
    assert h.hexdigest() == g["collection_sha256"]
