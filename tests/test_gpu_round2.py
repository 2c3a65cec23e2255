"""Round-2 additions, all through the C ABI on a GPU: resident databases (ArrayMatcher::Build puts the rows in HBM once),
k = 1 under ties, the streaming upload, clones, argument hardening, the device-planned second pass and its bounded
fallback, the coordinate de-duplication on the GPU against the reference's std::set, and the multi-GPU driver."""
import importlib
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
synth = importlib.import_module("3dreconstruction_b200.synth")
EXE = os.path.join(ROOT, "build", "compute_matches")


def test_k1_keeps_first_minimum_under_ties(ctx, pkg, l0):
    """SearchNeighbours(k=1) == partial_sort(first, first+1, last): the FIRST minimum; checked against
    the reference's own ArrayMatcherBruteForce."""
    db = synth.tie_set(71, 500, 2)
    db[3::4] = db[0]                                   # many exact duplicates of row 0
    q = np.concatenate([db[:1], synth.tie_set(72, 199, 2)])
    m = pkg.ArrayMatcherCuda(ctx)
    assert m.Build(db, len(db))
    for k in (1, 2):
        vi, vd = [], []
        assert m.SearchNeighbours(q, len(q), vi, vd, k)
        want = l0.knn(db, q, k)
        assert vi == want[0].reshape(-1).tolist(), f"k={k}"
        assert vd == want[1].reshape(-1).astype(float).tolist()
    assert vi[0] in (0, 3) and vd[0] == 0.0


def test_resident_db_reused_across_searches(ctx, pkg, l1):
    db = synth.uniform_set(81, 3000)
    rdb = ctx.db_create(db)
    try:
        for seed, nq in ((82, 1), (83, 300), (84, 1025)):
            q = synth.uniform_set(seed, nq)
            for tie in (pkg.TIE_REFERENCE, pkg.TIE_LOWEST_INDEX):
                idx, dist = rdb.knn2(q, tie)
                want = l1.knn2(db, q, tie)
                assert np.array_equal(idx, want[0]) and np.array_equal(dist.astype(np.int32), want[1])
        # the uploaded collection of the context is untouched by array-level calls
        descs = synth.collection(8, 3, 400)
        ctx.upload_images(descs)
        rdb.knn2(synth.uniform_set(85, 10))
        rs = float(pkg.square_f32(0.8))
        res = ctx.match_pairs(pkg.pairs_exhaustive(3), rs)
        for p, (i, j) in enumerate(res.pairs):
            assert np.array_equal(res.pair(p), l1.pair_matches(descs[i], descs[j], rs))
    finally:
        rdb.close()


def test_streamed_upload_equals_bulk_upload(ctx, pkg):
    descs = synth.collection(13, 7, 900) + [np.zeros((0, 128), np.uint8), synth.uniform_set(4, 3)]
    feats = [synth.features(13, k, len(d), dup_frac=0.1).reshape(-1, 4)[:, :2].copy() for k, d in enumerate(descs)]
    pairs = pkg.pairs_exhaustive(len(descs))
    rs = float(pkg.square_f32(0.8))
    ctx.upload_images(descs)
    ctx.set_features(feats)
    want = ctx.match_collection(pairs, rs)
    ctx.stream_images(descs, feats, order=[5, 0, 8, 3, 1, 7, 2, 6, 4])       # any order, each image once
    got = ctx.match_collection(pairs, rs)
    assert np.array_equal(got.counts, want.counts) and np.array_equal(got.matches, want.matches)
    assert want.counts.sum() > 100


def test_clone_then_array_level_uses_library_row_counts(ctx, pkg, l1):
    """knn2 on a cloned context sized its outputs from a stale Python cache."""
    descs = synth.collection(14, 3, 700)
    ctx.upload_images(descs)
    with pkg.Context(0) as rep:
        rep.clone_images_from(ctx)
        idx, dist = rep.knn2(0, 2, pkg.TIE_REFERENCE)
        want = l1.knn2(descs[0], descs[2], 1)
        assert idx.shape == (700, 2) and np.array_equal(idx, want[0]) and np.array_equal(dist.astype(np.int32), want[1])


def test_abi_rejects_bad_arguments_without_crashing(ctx, pkg):
    import ctypes as C
    lib = pkg.load_library()
    descs = synth.collection(15, 2, 300)
    ctx.upload_images(descs)
    pm = pkg.mvgcuda._PairMatches()
    pairs = np.array([[0, 1]], np.int32)
    p32 = pairs.ctypes.data_as(C.POINTER(C.c_int32))
    for n in (-1, -(1 << 40), 1 << 62):                # absurd pair counts: an error code, not std::length_error / bad_alloc
        assert lib.mvgcuda_match_pairs(ctx._h, n, p32, C.c_float(0.64), C.byref(pm)) != 0
        assert lib.mvgcuda_match_collection(ctx._h, n, p32, C.c_float(0.64), 0, C.byref(pm)) != 0
    assert lib.mvgcuda_match_pairs(ctx._h, 1, None, C.c_float(0.64), C.byref(pm)) != 0
    assert lib.mvgcuda_set_features(ctx._h, 2, None, None) != 0
    assert lib.mvgcuda_match_collection(ctx._h, 1, p32, C.c_float(0.64), 0, C.byref(pm)) != 0   # no features yet
    assert b"set_features" in lib.mvgcuda_last_error(ctx._h)
    assert lib.mvgcuda_upload_images(ctx._h, -3, None, None, 0) != 0
    assert lib.mvgcuda_stream_image(ctx._h, 99, None, None) != 0
    assert lib.mvgcuda_device_ordinal(0) >= 0 and lib.mvgcuda_device_ordinal(10 ** 6) == -1
    # the context still works afterwards
    assert len(ctx.match_pairs(pairs, 0.64)) == 1


@pytest.mark.parametrize("rescan_rows", [0, 64])
def test_second_pass_planned_on_device_and_bounded_fallback(ctx, pkg, l1, et, rescan_rows):
    """Default: the second pass is planned on the device (no host round trip).  With a 64-row gather buffer every batch
    overflows it and is repaired through the bounded multi-round path.  Same matches either way; several batches."""
    descs, feats = et
    big = [np.concatenate([descs[k % 9], descs[(k + 3) % 9]]) for k in range(14)]      # ~800 real SIFT rows each
    ctx.upload_images(big)
    pairs = pkg.pairs_exhaustive(len(big))
    ctx.set_tuning(0.8, rescan_rows)
    try:
        rs = float(pkg.square_f32(0.8))
        res = ctx.match_pairs(pairs, rs)
        assert res.timing["rescanned_queries"] > 0
        for p in range(0, len(pairs), 5):
            i, j = pairs[p]
            assert np.array_equal(res.pair(p), l1.pair_matches(big[i], big[j], rs)), (i, j)
    finally:
        ctx.set_tuning()


@pytest.mark.parametrize("rows,ratio", [(1500, 0.9), (7000, 1.0)])
def test_gpu_coordinate_dedup_equals_reference_std_set(ctx, pkg, l0, rows, ratio):
    """Row 13 on the GPU: heavy duplicate x / duplicate (x, y) features, against IndexedMatchDecorator<float>::getDeduplicated
    of the reference itself (its std::set with the non-strict-weak comparator).  The small case runs in the shared-memory
    kernel (16-bit nodes), the large one has pairs with more than 2,047 matches, which take the global-memory kernel."""
    descs = synth.collection(16, 4, rows)
    feats = []
    rng = np.random.default_rng(5)
    for k, d in enumerate(descs):
        f = synth.features(16, k, len(d), dup_frac=0.5)[:, :2].copy()
        f[rng.integers(0, len(f), 200), 1] = f[rng.integers(0, len(f), 200), 1]       # equal y with different x as well
        f[:, 0] = np.round(f[:, 0] / 50.0) * 50.0                                       # coarse x grid: long runs of equal x
        feats.append(f.astype(np.float32))
    ctx.upload_images(descs)
    ctx.set_features(feats)
    pairs = pkg.pairs_exhaustive(4)
    rs = float(pkg.square_f32(ratio))
    raw = ctx.match_pairs(pairs, rs)
    col = ctx.match_collection(pairs, rs)
    dropped = 0
    for p, (i, j) in enumerate(pairs):
        want = l0.dedup_xy(raw.pair(p), feats[i], feats[j])
        assert np.array_equal(col.pair(p), want), (i, j)
        dropped += len(raw.pair(p)) - len(want)
    assert dropped > 20, "the inputs must actually exercise the de-duplication"
    assert (raw.counts.max() > 2047) == (rows > 2047), raw.counts


def test_driver_sharded_over_two_contexts_is_byte_identical(pkg, et, tmp_path):
    """compute_matches --gpus 2 (two shards; on a one-GPU box both contexts live on GPU 0) == --gpus 1 == golden."""
    descs, feats = et
    outs = []
    for gpus in ("1", "2", "3"):
        d = tmp_path / f"g{gpus}"
        d.mkdir()
        names = []
        for k, (dd, ff) in enumerate(zip(descs, feats)):
            names.append(f"et{k:03d}.jpg")
            pkg.io.save_descs_bin(str(d / f"et{k:03d}.desc"), dd, 8)
            pkg.io.save_feats(str(d / f"et{k:03d}.feat"), ff)
        (d / "lists.txt").write_text("".join(f"{n};640;480\n" for n in names))
        out = subprocess.run([EXE, "-i", str(d), "-o", str(d), "-r", "0.8", "--gpus", gpus], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr + out.stdout
        outs.append((d / "matches.putative.txt").read_bytes())
    assert outs[0] == outs[1] == outs[2] == open(os.path.join(GOLDEN, "et_putative_r0.8.txt"), "rb").read()
    # resume: the exported file is imported again (pairedIndexedMatchImport) and matching is skipped
    out = subprocess.run([EXE, "-i", str(tmp_path / "g2"), "-o", str(tmp_path / "g2"), "-r", "0.8"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "PREVIOUS RESULTS LOADED" in out.stdout and "36 pairs, 1647 putative matches imported" in out.stdout


def test_driver_refuses_feat_desc_row_mismatch(pkg, et, tmp_path):
    """Documented deviation: the reference takes the row count from the .feat file and over-reads
    the descriptors; the driver refuses such a collection."""
    descs, feats = et
    for k in range(3):
        pkg.io.save_descs_bin(str(tmp_path / f"a{k}.desc"), descs[k], 8)
        pkg.io.save_feats(str(tmp_path / f"a{k}.feat"), feats[k] if k != 1 else feats[k][:-5])
    (tmp_path / "lists.txt").write_text("".join(f"a{k}.jpg;640;480\n" for k in range(3)))
    out = subprocess.run([EXE, "-i", str(tmp_path), "-o", str(tmp_path), "--gpus", "1"], capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 and "row counts differ" in out.stderr
    assert not (tmp_path / "matches.putative.txt").exists()
