"""GPU parity tests proper: the CUDA path, called THROUGH THE C ABI (ctypes -> libmvgcuda.so), against the oracle
(L1 always; L0 = the reference's own code when oracle/_ref travelled) on the same seeded inputs, against the committed
golden fixtures, and -- at BASELINE.json's full sizes -- through size-independent properties.

Bar: bit-exact.  All arithmetic on the path is integer except the fp32 ratio test, which is reproduced operation by
operation (so its tolerance is 0 as well)."""
import hashlib
import importlib
import importlib.util
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

synth = importlib.import_module("3dreconstruction_b200.synth")
pkg_io = importlib.import_module("3dreconstruction_b200.io")


def _mg():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _assert_knn_equal(got, want_idx, want_dist, what=""):
    idx, dist = got
    assert dist.dtype == np.float32
    assert np.array_equal(dist.astype(np.int64), np.asarray(want_dist, np.int64)), f"{what}: distances differ"   # P1
    assert np.array_equal(idx, want_idx), f"{what}: indices differ"                                               # P2


# ---------------------------------------------------------------- array level (ArrayMatcher::SearchNeighbours)

@pytest.mark.parametrize("case", ["tie2", "tie4", "uniform", "sift", "two_rows", "three_rows"])
def test_knn2_golden(ctx, pkg, synth_golden, case):
    mg = _mg()
    db, q = mg.knn_case_inputs([c for c in mg.KNN_CASES if c[0] == case][0])
    _assert_knn_equal(ctx.knn2_arrays(db, q, pkg.TIE_REFERENCE), synth_golden[f"knn_{case}_idx"], synth_golden[f"knn_{case}_dist"], case)


@pytest.mark.parametrize("rows_db,rows_q", [(2, 1), (3, 2), (127, 129), (128, 128), (255, 257), (256, 256), (257, 255),
                                            (511, 1), (513, 383), (1000, 777), (4097, 300)])
@pytest.mark.parametrize("kind", ["uniform", "tie"])
def test_knn2_shapes_vs_oracle(ctx, pkg, l1, rows_db, rows_q, kind):
    if kind == "uniform":
        db, q = synth.uniform_set(rows_db * 7 + 1, rows_db), synth.uniform_set(rows_q * 11 + 2, rows_q)
    else:
        db, q = synth.tie_set(rows_db * 7 + 1, rows_db, 3), synth.tie_set(rows_q * 11 + 2, rows_q, 3)
    for tie in (pkg.TIE_REFERENCE, pkg.TIE_LOWEST_INDEX):
        want = l1.knn2(db, q, tie)
        _assert_knn_equal(ctx.knn2_arrays(db, q, tie), want[0], want[1], f"{kind} {rows_db}x{rows_q} tie={tie}")


def test_knn2_extreme_values(ctx, pkg, l1):
    db = np.zeros((300, 128), np.uint8)
    db[::2] = 255
    q = np.full((130, 128), 255, np.uint8)
    q[::3] = 0
    want = l1.knn2(db, q, 1)
    got = ctx.knn2_arrays(db, q, pkg.TIE_REFERENCE)
    _assert_knn_equal(got, want[0], want[1], "extreme")
    # farthest possible pair: all-255 query against an all-zero database
    idx, dist = ctx.knn2_arrays(np.zeros((2, 128), np.uint8), np.full((1, 128), 255, np.uint8), pkg.TIE_LOWEST_INDEX)
    assert dist[0].tolist() == [8323200.0, 8323200.0] and idx[0].tolist() == [0, 1]


def test_knn2_errors_like_reference(ctx, pkg):
    some = synth.uniform_set(1, 10)
    with pytest.raises(pkg.MvgCudaError, match="Too much asked nearest neighbors"):    # matcher_brute_force.h:107-110
        ctx.knn2_arrays(some[:1], some)
    with pytest.raises(pkg.MvgCudaError, match="Too much asked nearest neighbors"):
        ctx.knn2_arrays(some, some[:0])


def test_array_matcher_mirror_semantics(ctx, pkg, l1, capsys):
    db, q = synth.uniform_set(31, 400), synth.uniform_set(32, 50)
    m = pkg.ArrayMatcherCuda(ctx)
    assert m.Build(db, 0) is False
    vi, vd = [7], [1.5]
    assert m.SearchNeighbours(q, 50, vi, vd, 2) is False and vi == [7]        # no db -> false, outputs untouched
    assert "Too much asked nearest neighbors" in capsys.readouterr().err
    assert m.Build(db, len(db), 128)
    assert m.SearchNeighbours(q, 50, vi, vd, 2)
    want = l1.knn2(db, q, 1)
    assert vi[0] == 7 and vd[0] == 1.5                                          # APPENDS (push_back, :128-131)
    assert vi[1:] == want[0].reshape(-1).tolist() and vd[1:] == want[1].reshape(-1).astype(float).tolist()
    ok, i, d = m.SearchNeighbour(q[3])
    assert ok and i == l1.knn2(db, q[3:4], 0)[0][0, 0] and d == want[1][3, 0]


# ---------------------------------------------------------------- pair level (rows 7-12 on the GPU)

def _check_pairs_vs_l1(ctx, pkg, l1, descs, ratios=(0.6, 0.8)):
    ctx.upload_images(descs)
    pairs = pkg.pairs_exhaustive(len(descs))
    for r in ratios:
        rs = float(pkg.square_f32(r))
        res = ctx.match_pairs(pairs, rs)
        assert len(res) == len(pairs) and res.offsets[-1] == len(res.matches)
        for p, (i, j) in enumerate(pairs):
            want = l1.pair_matches(descs[i], descs[j], rs)
            assert np.array_equal(res.pair(p), want), f"pair ({i},{j}) ratio {r}"
    return res


def test_match_pairs_small_collection(ctx, pkg, l1):
    descs, _ = _mg().collection_inputs()        # ragged + 1-row + 2-row + tie-heavy members
    res = _check_pairs_vs_l1(ctx, pkg, l1, descs)
    assert res.timing["knn_kernel_launches"] >= 1


def test_match_pairs_empty_and_degenerate(ctx, pkg, l1):
    descs = [synth.uniform_set(1, 300), np.zeros((0, 128), np.uint8), synth.uniform_set(2, 1), synth.uniform_set(3, 2)]
    _check_pairs_vs_l1(ctx, pkg, l1, descs, ratios=(0.8, 1.0))
    res = ctx.match_pairs(np.zeros((0, 2), np.int32), 0.64)
    assert len(res) == 0 and len(res.matches) == 0
    with pytest.raises(pkg.MvgCudaError):
        ctx.match_pairs(np.array([[0, 9]], np.int32), 0.64)


def test_match_pairs_reverse_and_self_pairs(ctx, pkg, l1):
    descs = synth.collection(5, 3, 700)
    ctx.upload_images(descs)
    pairs = np.array([[1, 0], [2, 2], [0, 1], [0, 1]], np.int32)   # not i<j, self pair, duplicates
    rs = float(pkg.square_f32(0.8))
    res = ctx.match_pairs(pairs, rs)
    for p, (i, j) in enumerate(pairs):
        assert np.array_equal(res.pair(p), l1.pair_matches(descs[i], descs[j], rs))


def test_ratio_bit_pattern_matters(ctx, pkg, l1):
    # same inputs, the two fp32 roundings of 0.64: result must follow the value PASSED (the caller computes Square(0.8f))
    descs = synth.collection(9, 2, 3000)
    ctx.upload_images(descs)
    pairs = np.array([[0, 1]], np.int32)
    for rs in (float(pkg.square_f32(0.8)), float(np.float32(0.64))):
        assert np.array_equal(ctx.match_pairs(pairs, rs).pair(0), l1.pair_matches(descs[0], descs[1], rs))


def _row_at_distance(qrow: np.ndarray, d: int, rng) -> np.ndarray:
    """A u8 row at squared-L2 distance exactly d from qrow (qrow must leave +-11 of head-room in every bin)."""
    delta = np.zeros(128, np.int16)
    pos = rng.permutation(128)
    k = 0
    while d > 0:
        v = min(11, int(np.sqrt(d)))
        delta[pos[k]] = v if rng.random() < 0.5 else -v
        d -= v * v
        k += 1
    return (qrow.astype(np.int16) + delta).astype(np.uint8)


@pytest.mark.parametrize("rho,rescan_rows", [(0.8, 0), (1.0, 0), (0.65, 256)])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_ratio_aware_pruning_adversarial_orders(ctx, pkg, l1, seed, rho, rescan_rows):
    """Scan orders built to break an unsound pruning rule: per query a handful of near rows
    whose distances straddle the ratio threshold (e.g. 150, 200 early, then 130, then 90: the true pair (90, 130) fails
    at 0.8^2 although (90, 150) would pass), scattered over tiles and column halves among far filler rows."""
    rng = np.random.default_rng(4242 + seed)
    n_q, n_db = 96, 2300
    levels = rng.permutation(np.arange(20, 20 + 2 * n_q, 2))[:n_q]          # every query sits at its own grey level
    q = np.repeat(levels[:, None], 128, axis=1).astype(np.uint8)
    menu = [90, 100, 110, 128, 130, 140, 150, 160, 200, 250, 90, 141, 140, 157, 156]
    dbs = []
    for _ in range(6):
        db = rng.integers(0, 256, (n_db, 128), dtype=np.uint8)                # fillers: distance ~ 1e6
        slots = rng.permutation(n_db)
        k = 0
        for j in range(n_q):
            for d in rng.choice(menu, int(rng.integers(2, 7)), replace=True):
                db[slots[k]] = _row_at_distance(q[j], int(d), rng)
                k += 1
        dbs.append(db)
    ctx.upload_images(dbs + [q])
    pairs = np.array([[i, len(dbs)] for i in range(len(dbs))], np.int32)
    ctx.set_tuning(rho, rescan_rows)
    try:
        rescanned = 0
        for r in (0.8, 0.6, 0.95, 1.0):
            rs = float(pkg.square_f32(r))
            res = ctx.match_pairs(pairs, rs)
            rescanned += res.timing["rescanned_queries"]
            n_pass = 0
            for p, (i, j) in enumerate(pairs):
                want = l1.pair_matches(dbs[i], q, rs)
                assert np.array_equal(res.pair(p), want), f"ratio {r} db {i}"
                n_pass += len(want)
            if r in (0.8, 0.95):
                assert 0 < n_pass < len(pairs) * (n_q - 1)                    # both outcomes occur
        assert (rescanned == 0) if rho == 1.0 else (rescanned > 0)            # the exact second pass really ran
    finally:
        ctx.set_tuning()


@pytest.mark.parametrize("rho,rescan_rows", [(0.8, 0), (0.65, 128), (0.9, 1000)])
def test_rescan_of_ambiguous_queries_real_sift(ctx, pkg, l1, et, rho, rescan_rows):
    """Real SIFT (data/et, 9 images, both directions of every pair): many queries sit in the ambiguous band, so the exact
    second pass (gather -> same kernel -> scatter), its multi-round buffer and the splitting of a pair over rounds all
    run; the matches must not depend on the tuning."""
    descs, _ = et
    ctx.upload_images(descs)
    pairs = pkg.pairs_exhaustive(len(descs))
    pairs = np.concatenate([pairs, pairs[:, ::-1]]).astype(np.int32)
    ctx.set_tuning(rho, rescan_rows)
    try:
        for r in (0.6, 0.8):
            rs = float(pkg.square_f32(r))
            res = ctx.match_pairs(pairs, rs)
            assert res.timing["rescanned_queries"] > 0
            for p, (i, j) in enumerate(pairs):
                assert np.array_equal(res.pair(p), l1.pair_matches(descs[i], descs[j], rs)), f"pair {i},{j} ratio {r}"
    finally:
        ctx.set_tuning()


def test_clone_images_replica(ctx, pkg, l1):
    """mvgcuda_clone_images: a second context takes its arena (and features) from the first device to device -- the
    path the multi-GPU driver uses instead of one PCIe upload per GPU -- and matches identically."""
    descs = synth.collection(12, 4, 900) + [synth.uniform_set(3, 0)]          # incl. an empty image
    feats = [synth.features(12, k, len(d)).reshape(-1, 4)[:, :2].copy() for k, d in enumerate(descs)]
    ctx.upload_images(descs)
    ctx.set_features(feats)
    pairs = pkg.pairs_exhaustive(len(descs))
    rs = float(pkg.square_f32(0.8))
    want = ctx.match_collection(pairs, rs)
    with pkg.Context(0) as replica:
        replica.clone_images_from(ctx)
        ctx.upload_images([synth.uniform_set(1, 300)])                          # the source may move on afterwards
        got = replica.match_collection(pairs, rs)
        assert np.array_equal(got.counts, want.counts) and np.array_equal(got.matches, want.matches)
        for p in (0, 3, 5):
            i, j = pairs[p]
            assert np.array_equal(got.pair(p), l1.dedup_xy(l1.pair_matches(descs[i], descs[j], rs), feats[i], feats[j]))
        with pytest.raises(pkg.MvgCudaError):
            replica.clone_images_from(replica)


# ---------------------------------------------------------------- collection level + export (rows 7-14)

@pytest.mark.parametrize("r", [0.6, 0.8])
def test_collection_golden_et(ctx, pkg, et, r, tmp_path):
    """The reference's real SIFT set: matches.putative.txt byte-identical to the reference BF run (P5)."""
    descs, feats = et
    meta = json.load(open(os.path.join(GOLDEN, "et_golden.json")))
    m = pkg.MatcherCudaAllInMemory(r, ctx)
    assert m.LoadArrays(descs, [f[:, :2] for f in feats])
    pw = m.Match()
    assert len(pw) == 36 and sum(len(v) for v in pw.values()) == meta[f"r{r}"]["matches"]
    out = str(tmp_path / "matches.putative.txt")
    m.Export(out)
    data = open(out, "rb").read()
    assert hashlib.sha256(data).hexdigest() == meta[f"r{r}"]["sha256"]
    assert data == open(os.path.join(GOLDEN, f"et_putative_r{r}.txt"), "rb").read()


def test_collection_from_files_like_compute_matches(ctx, pkg, et, tmp_path):
    """LoadData from .feat/.desc files (4-byte count headers as shipped in data/et, and 8-byte as written on Linux)."""
    descs, feats = et
    names = []
    for k, (d, f) in enumerate(zip(descs, feats)):
        names.append(f"et{k:03d}.jpg")
        pkg_io.save_descs_bin(str(tmp_path / f"et{k:03d}.desc"), d, 4 if k % 2 else 8)
        pkg_io.save_feats(str(tmp_path / f"et{k:03d}.feat"), f)
    m = pkg.MatcherCudaAllInMemory(0.6, ctx)
    assert m.LoadData(names, str(tmp_path))
    m.Match(names)
    out = str(tmp_path / "matches.putative.txt")
    m.Export(out)
    assert open(out, "rb").read() == open(os.path.join(GOLDEN, "et_putative_r0.6.txt"), "rb").read()


@pytest.mark.parametrize("r", [0.6, 0.8])
def test_collection_golden_synthetic(ctx, pkg, synth_golden, r, tmp_path):
    descs, feats = _mg().collection_inputs()
    m = pkg.MatcherCudaAllInMemory(r, ctx, host_threads=3)
    m.LoadArrays(descs, [f[:, :2] for f in feats])
    m.Match()
    out = str(tmp_path / "m.txt")
    m.Export(out)
    assert open(out, "rb").read() == synth_golden[f"collection_text_r{r}"].tobytes()


def test_collection_vs_reference_code_directly(ctx, pkg, l0, tmp_path):
    """When oracle/_ref (the reference's own MatcherAllInMemory) is available: byte-compare on a fresh random set."""
    descs = synth.collection(123, 5, 800)
    feats = [synth.features(123, k, len(d), dup_frac=0.1) for k, d in enumerate(descs)]
    want = l0.match_collection_text(descs, feats, 0.8)
    m = pkg.MatcherCudaAllInMemory(0.8, ctx)
    m.LoadArrays(descs, [f[:, :2] for f in feats])
    m.Match()
    out = str(tmp_path / "m.txt")
    m.Export(out)
    assert open(out, "rb").read() == want
    # file-boundary acceptance: our export round-trips through the reference's own
    # pairedIndexedMatchImport / PairedIndexedMatchToStream byte-identically
    back = str(tmp_path / "back.txt")
    l0.roundtrip_matches(out, back)
    assert open(back, "rb").read() == want


# ---------------------------------------------------------------- BASELINE.json sizes

def test_full_size_pair_10k_vs_oracle(ctx, pkg, l1):
    """configs[1]: one 10k x 10k pair, every query, bit-exact against the oracle (0.3 s of CPU)."""
    a, b = synth.collection(2, 2, 10000)
    want = l1.knn2(a, b, 1)
    _assert_knn_equal(ctx.knn2_arrays(a, b, pkg.TIE_REFERENCE), want[0], want[1], "10k sift")
    u, v = synth.uniform_set(41, 10000), synth.uniform_set(42, 10000)
    want = l1.knn2(u, v, 1)
    _assert_knn_equal(ctx.knn2_arrays(u, v, pkg.TIE_REFERENCE), want[0], want[1], "10k uniform")


def test_full_size_40k_tiled_epilogue(ctx, pkg, l1):
    """configs[4] shape: 40k rows (157 db tiles per query block, indices need 16 bits)."""
    a = synth.image(5, 0, 40000, synth.scene_pool(5, 40000))
    b = synth.image(5, 1, 40000, synth.scene_pool(5, 40000))
    want = l1.knn2(a, b, 1)
    got = ctx.knn2_arrays(a, b, pkg.TIE_REFERENCE)
    _assert_knn_equal(got, want[0], want[1], "40k")
    assert got[0].max() > 65535 // 2
    # pair level at the same size: the batch is long enough for the filter-first epilogue schedule, with pruning and the
    # exact second pass on top
    ctx.upload_images([a, b])
    rs = float(pkg.square_f32(0.8))
    res = ctx.match_pairs(np.array([[0, 1]], np.int32), rs)
    assert np.array_equal(res.pair(0), l1.pair_matches(a, b, rs))


def test_config3_slice_vs_oracle_and_properties(ctx, pkg, l1):
    """configs[2] inputs (config 3 generator, 10k rows): a 12-image slice (66 pairs) bit-exact vs the oracle, then
    size-independent properties over it."""
    descs = synth.collection(3, 12, 10000)
    ctx.upload_images(descs)
    pairs = pkg.pairs_exhaustive(12)
    rs = float(pkg.square_f32(0.8))
    res = ctx.match_pairs(pairs, rs)
    for p, (i, j) in enumerate(pairs):
        assert np.array_equal(res.pair(p), l1.pair_matches(descs[i], descs[j], rs)), (i, j)
    assert res.counts.min() > 50, "planted correspondences must survive the ratio test"
    # idempotence / determinism
    res2 = ctx.match_pairs(pairs, rs)
    assert np.array_equal(res.matches, res2.matches) and np.array_equal(res.counts, res2.counts)
    # pair-order independence (the scheduler may shard the list anywhere)
    perm = np.random.default_rng(0).permutation(len(pairs))
    res3 = ctx.match_pairs(pairs[perm], rs)
    for k, p in enumerate(perm):
        assert np.array_equal(res3.pair(k), res.pair(p))
    # within a pair: ascending _j, no repeated consecutive _i, indices in range
    for p in range(len(pairs)):
        m = res.pair(p)
        assert (np.diff(m[:, 1]) > 0).all() and (np.diff(m[:, 0]) != 0).all()
        assert m[:, 0].max() < 10000 and m[:, 1].max() < 10000


def test_self_match_and_permutation_properties(ctx, pkg):
    a = synth.uniform_set(77, 10000)            # rows are distinct with overwhelming probability
    idx, dist = ctx.knn2_arrays(a, a, pkg.TIE_LOWEST_INDEX)
    assert (dist[:, 0] == 0).all() and np.array_equal(idx[:, 0], np.arange(10000))
    q = synth.uniform_set(78, 3000)
    perm = np.random.default_rng(1).permutation(10000)
    i1, d1 = ctx.knn2_arrays(a, q, pkg.TIE_LOWEST_INDEX)
    i2, d2 = ctx.knn2_arrays(a[perm], q, pkg.TIE_LOWEST_INDEX)
    assert np.array_equal(d1, d2)
    uniq = d1[:, 0] < d1[:, 1]
    assert np.array_equal(perm[i2[uniq, 0]], i1[uniq, 0])


@pytest.mark.parametrize("name", ["sceaux", "ace"])
@pytest.mark.parametrize("r", [0.6, 0.8])
def test_collection_golden_imagedata(ctx, pkg, name, r, tmp_path):
    """BASELINE configs[0]: bundled data/imageData image pairs (real RootSIFT u8 regions from the reference's own VLFeat
    wrapper), putative file byte-identical to the reference's brute-force run."""
    z = np.load(os.path.join(GOLDEN, "imagedata_collection.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "imagedata_golden.json")))[name]
    descs = [z[f"{name}_desc_{k}"] for k in range(2)]
    feats = [z[f"{name}_feat_{k}"][:, :2] for k in range(2)]
    m = pkg.MatcherCudaAllInMemory(r, ctx)
    m.LoadArrays(descs, feats)
    pw = m.Match()
    assert sum(len(v) for v in pw.values()) == meta[f"r{r}"]["matches"]
    out = str(tmp_path / "m.txt")
    m.Export(out)
    data = open(out, "rb").read()
    assert data == z[f"{name}_text_r{r}"].tobytes() and hashlib.sha256(data).hexdigest() == meta[f"r{r}"]["sha256"]


def test_config4_shape_8k_rows(ctx, pkg, l1):
    """configs[3] shape (8,000 rows per image): a 6-image slice, every pair bit-exact vs the oracle."""
    descs = synth.collection(4, 6, 8000)
    ctx.upload_images(descs)
    pairs = pkg.pairs_exhaustive(6)
    rs = float(pkg.square_f32(0.8))
    res = ctx.match_pairs(pairs, rs)
    for p, (i, j) in enumerate(pairs):
        assert np.array_equal(res.pair(p), l1.pair_matches(descs[i], descs[j], rs)), (i, j)


def test_ragged_collection_many_sizes(ctx, pkg, l1):
    """Ragged images (sizes straddling every tile / block / chunk boundary) in one batch."""
    sizes = [2, 3, 15, 16, 17, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 513, 1025]
    descs = [synth.tie_set(900 + k, n, 4) if k % 2 else synth.uniform_set(900 + k, n) for k, n in enumerate(sizes)]
    ctx.upload_images(descs)
    pairs = pkg.pairs_exhaustive(len(sizes))
    rs = float(pkg.square_f32(0.8))
    res = ctx.match_pairs(pairs, rs)
    for p, (i, j) in enumerate(pairs):
        assert np.array_equal(res.pair(p), l1.pair_matches(descs[i], descs[j], rs)), (sizes[i], sizes[j])


def test_ratio_above_one_uses_reference_tie_rule(ctx, pkg, l1):
    """With ratio > 1 a tie d1 == d2 passes the ratio test, so the reported _i depends on the tie rule: the pair path
    must then reproduce the reference's std::partial_sort choice."""
    descs = [synth.tie_set(301, 900, 2), synth.tie_set(302, 700, 2), synth.tie_set(303, 257, 3)]
    ctx.upload_images(descs)
    pairs = pkg.pairs_exhaustive(3)
    for r in (1.0, 1.2):
        rs = float(pkg.square_f32(r))
        res = ctx.match_pairs(pairs, rs)
        for p, (i, j) in enumerate(pairs):
            assert np.array_equal(res.pair(p), l1.pair_matches(descs[i], descs[j], rs)), (i, j, r)
        if r > 1:
            assert res.counts.sum() > 100
