"""Host-side logic that needs no GPU: file formats, pair enumeration, the multi-GPU pair scheduler (incl. a
world_size-2 gloo run), synthetic generator determinism."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

sharding = importlib.import_module("3dreconstruction_b200.sharding")
pkg_io = importlib.import_module("3dreconstruction_b200.io")
synth = importlib.import_module("3dreconstruction_b200.synth")


def test_desc_roundtrip_both_headers(tmp_path):
    d = synth.uniform_set(1, 37)
    for hdr in (8, 4):
        p = str(tmp_path / f"a{hdr}.desc")
        pkg_io.save_descs_bin(p, d, hdr)
        assert os.path.getsize(p) == hdr + 37 * 128
        assert np.array_equal(pkg_io.load_descs_bin(p), d)
    assert pkg_io.load_descs_bin(str(tmp_path / "missing.desc")).shape == (0, 128)


def test_feat_text_roundtrip(tmp_path):
    f = synth.features(3, 0, 50)
    p = str(tmp_path / "a.feat")
    pkg_io.save_feats(p, f)
    assert np.array_equal(pkg_io.load_feats(p), f)      # 6 significant digits survive the text round trip
    assert open(p).readline().count(" ") == 3


def test_matches_text_roundtrip(et):
    from conftest import GOLDEN
    txt = open(os.path.join(GOLDEN, "et_putative_r0.8.txt")).read()
    pw = pkg_io.matches_from_text(txt)
    assert len(pw) == 36 and pkg_io.matches_to_text(pw) == txt
    assert "0 1\n" in txt and list(pw)[0] == (0, 1)


def test_pairs_exhaustive_order(pkg):
    p = pkg.pairs_exhaustive(4)
    assert p.tolist() == [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]]
    assert len(pkg.pairs_exhaustive(100)) == 4950


def test_shard_bounds_cover_and_balance(pkg):
    rows = [10000] * 100
    pairs = pkg.pairs_exhaustive(100)
    for world in (1, 2, 4, 8):
        got = [sharding.shard_pairs(pairs, rows, r, world) for r in range(world)]
        assert np.array_equal(np.concatenate([g[0] for g in got]), pairs)
        sizes = [len(g[0]) for g in got]
        assert max(sizes) - min(sizes) <= 1
    # ragged images: cost-balanced, contiguous, complete
    rng = np.random.default_rng(0)
    rows = rng.integers(0, 40000, 50).tolist()
    pairs = pkg.pairs_exhaustive(50)
    costs = sharding.pair_costs(pairs, rows)
    b = sharding.shard_bounds(costs, 8)
    assert b[0] == 0 and b[-1] == len(pairs) and all(x <= y for x, y in zip(b, b[1:]))
    per = [costs[b[k]:b[k + 1]].sum() for k in range(8)]
    assert max(per) <= costs.sum() / 8 + costs.max()


def test_weak_scaling_collection_sizes():
    assert [sharding.images_for_pairs_per_gpu(w) for w in (1, 2, 4, 8)] == [100, 142, 200, 282]


def test_synth_deterministic_and_sift_like():
    a = synth.collection(3, 3, 2000)
    b = synth.collection(3, 3, 2000)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    norms = np.linalg.norm(a[0].astype(np.float64), axis=1)
    assert 480 < norms.mean() < 530 and a[0].max() <= 255
    assert 0.04 < (a[0] == 0).mean() < 0.2


_GLOO_WORKER = r'''
import importlib, os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MVG_ROOT"])
pkg = importlib.import_module("3dreconstruction_b200")
sharding = importlib.import_module("3dreconstruction_b200.sharding")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rows = [1000 + 10 * k for k in range(30)]
pairs = pkg.pairs_exhaustive(30)
mine, (b, e) = sharding.shard_pairs(pairs, rows, rank, world)
# every rank reports (count, checksum); the union must be the whole list with no overlap
t = torch.tensor([len(mine), int(mine.astype(np.int64).sum()), b, e], dtype=torch.int64)
out = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(out, t)
ms = torch.tensor([float(rank + 1)])
dist.all_reduce(ms, op=dist.ReduceOp.MAX)   # the "max over ranks" timing reduction bench.py uses
if rank == 0:
    print(json.dumps({"parts": [o.tolist() for o in out], "total": len(pairs), "sum": int(pairs.astype(np.int64).sum()), "max": ms.item()}))
dist.destroy_process_group()
'''


def test_shard_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MVG_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert sum(p[0] for p in r["parts"]) == r["total"] and sum(p[1] for p in r["parts"]) == r["sum"]
    assert r["parts"][0][3] == r["parts"][1][2] and r["parts"][0][2] == 0 and r["parts"][1][3] == r["total"]
    assert r["max"] == 2.0


def test_rbtree_dedup_restatement_matches_std_set(tmp_path):
    """3dreconstruction_b200/csrc/rbtree_dedup.cuh (the device code of row 13, compiled for the host: restatement of libstdc++'s std::set range construction for the
    reference's non-strict-weak coordinate comparator) against the real std::set, on the CPU."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = tmp_path / "test_rbtree"
    src = os.path.join(ROOT, "tests", "native", "test_rbtree_dedup.cpp")
    subprocess.check_call([gxx, "-std=c++17", "-O2", "-o", str(exe), src])
    r = subprocess.run([str(exe), "8000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rbtree_dedup == std::set" in r.stdout


def test_upload_friendly_order_is_a_permutation_with_a_streamable_head():
    """The pair order of the streamed path: same pairs; the head (pairs among the first third of the images) by larger image
    id, so that batch k only touches images 0..k; the tail in the reference's (i, j) order (db image constant over runs)."""
    import importlib
    pkg = importlib.import_module("3dreconstruction_b200")
    for n in (2, 9, 30, 100):
        pairs = pkg.pairs_exhaustive(n)
        got = pkg.upload_friendly_order(pairs)
        assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, pairs.tolist()))
        m = max(8, (n + 2) // 3)
        head = [p for p in got.tolist() if max(p) < m]
        tail = [p for p in got.tolist() if max(p) >= m]
        assert got.tolist() == head + tail
        assert [max(p) for p in head] == sorted(max(p) for p in head)
        assert tail == sorted(tail)
    # a shard of the list (one rank's contiguous chunk) and an explicit head size
    shard = pkg.pairs_exhaustive(50)[300:700]
    got = pkg.upload_friendly_order(shard, head_images=20)
    assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, shard.tolist()))
    assert len(pkg.upload_friendly_order(np.zeros((0, 2), np.int32))) == 0


def test_pair_chain_model(tmp_path):
    """csrc/pair_chain.h (when the geometric filter starts its pairs and at which rand() offset: speculative starts, guesses,
    refutations, held pairs) under a randomised model of the pipeline -- 4,000 collections, verdicts in random order, wrong
    starts reporting arbitrary counts: every pair is counted once and its last start is at the reference's offset."""
    exe = str(tmp_path / "test_pair_chain")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "native", "test_pair_chain.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PAIR CHAIN OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
