"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE (oracle L0 = the reference's own
headers, oracle/build_ref.sh).  Run in the build container, where /root/reference is mounted:

    python tests/golden/make_golden.py

Outputs (committed):
  et_collection.npz   inputs: the reference's shipped real-SIFT set data/et/et00{0..8}.{desc,feat} (3,508 descriptors;
                      .desc carry a 4-byte count), as arrays desc_k / feat_k
  et_golden.json      outputs: sha256 + match counts of the reference BF collection matcher
                      (MatcherAllInMemory<..,ArrayMatcherBruteForce<uchar,SquaredEuclideanDistanceVectorized>>,
                      matcher_all_in_memory.h:62-141 -> PairedIndexedMatchToStream) at distRatio 0.6 and 0.8, and the
                      per-pair counts of the shipped (Windows, FLANN) data/et/matches.putative.txt for the soft check
  et_putative_r0.6.txt / et_putative_r0.8.txt   the reference BF output itself (byte-exact golden)
  synth_golden.npz    reference outputs on seeded synthetic inputs (inputs are regenerated from the seeds by
                      3dreconstruction_b200.synth, so only outputs are stored): raw SearchNeighbours (idx, dist)
                      on tie-heavy / uniform / SIFT-like sets, DistanceRatioFilter lists, per-pair match lists,
                      and a small full collection's putative text.
"""
import hashlib
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

pkg_io = importlib.import_module("3dreconstruction_b200.io")
synth = importlib.import_module("3dreconstruction_b200.synth")
REF = os.environ.get("MVG_REF", "/root/reference")

# (name, kind, seed_db, seed_q, rows_db, rows_q, arg)
KNN_CASES = [
    ("tie2", "tie", 101, 102, 700, 300, 2),
    ("tie4", "tie", 103, 104, 513, 257, 4),
    ("uniform", "uniform", 105, 106, 1000, 777, 0),
    ("sift", "sift", 107, 0, 2500, 2500, 0),
    ("two_rows", "uniform", 108, 109, 2, 5, 0),
    ("three_rows", "tie", 110, 111, 3, 64, 2),
]
COLLECTION = dict(config=77, n_images=6, rows=900)


def knn_case_inputs(case):
    name, kind, s_db, s_q, r_db, r_q, arg = case
    if kind == "tie":
        return synth.tie_set(s_db, r_db, arg), synth.tie_set(s_q, r_q, arg)
    if kind == "uniform":
        return synth.uniform_set(s_db, r_db), synth.uniform_set(s_q, r_q)
    imgs = synth.collection(s_db, 2, r_db)
    return imgs[0], imgs[1][:r_q]


def collection_inputs():
    c = COLLECTION
    descs = synth.collection(c["config"], c["n_images"], c["rows"])
    # ragged + degenerate members: a 1-row image, a 2-row image, a tie-heavy image
    descs[3] = descs[3][:431]
    descs.append(synth.uniform_set(901, 1))
    descs.append(synth.uniform_set(902, 2))
    descs.append(synth.tie_set(903, 300, 3))
    feats = [synth.features(c["config"], k, len(d), dup_frac=0.05) for k, d in enumerate(descs)]
    return descs, feats


def main():
    oracle.build_l0()
    l0 = oracle.L0()
    # ---- data/et
    descs, feats = [], []
    for k in range(9):
        descs.append(pkg_io.load_descs_bin(os.path.join(REF, "data", "et", f"et{k:03d}.desc")))
        feats.append(pkg_io.load_feats(os.path.join(REF, "data", "et", f"et{k:03d}.feat")))
        assert len(descs[-1]) == len(feats[-1]), k
    np.savez_compressed(os.path.join(HERE, "et_collection.npz"),
                        **{f"desc_{k}": d for k, d in enumerate(descs)}, **{f"feat_{k}": f for k, f in enumerate(feats)})
    meta = {"rows": [int(len(d)) for d in descs]}
    for r in (0.6, 0.8):
        txt = l0.match_collection_text(descs, feats, r)
        with open(os.path.join(HERE, f"et_putative_r{r}.txt"), "wb") as f:
            f.write(txt)
        pw = pkg_io.matches_from_text(txt.decode())
        meta[f"r{r}"] = {"sha256": hashlib.sha256(txt).hexdigest(), "pairs": len(pw), "matches": int(sum(len(v) for v in pw.values()))}
    with open(os.path.join(REF, "data", "et", "matches.putative.txt")) as f:
        shipped = pkg_io.matches_from_text(f.read())
    meta["shipped_flann_r0.6"] = {"pairs": len(shipped), "matches": int(sum(len(v) for v in shipped.values())),
                                  "per_pair": {f"{i} {j}": int(len(v)) for (i, j), v in sorted(shipped.items())}}
    with open(os.path.join(HERE, "et_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("data/et:", {k: v for k, v in meta.items() if k.startswith("r")}, "shipped", meta["shipped_flann_r0.6"]["matches"])

    # ---- synthetic
    out = {}
    for case in KNN_CASES:
        db, q = knn_case_inputs(case)
        idx, dist = l0.knn(db, q, 2)
        out[f"knn_{case[0]}_idx"] = idx
        out[f"knn_{case[0]}_dist"] = dist.astype(np.int32)
        assert np.array_equal(dist, dist.astype(np.int32).astype(np.float32))
        for r in (0.6, 0.8):
            rs = float(l0.square(r))
            out[f"knn_{case[0]}_pass_r{r}"] = l0.ratio_filter(dist, rs)
            out[f"knn_{case[0]}_matches_r{r}"] = l0.pair_matches(db, q, rs)
    cdescs, cfeats = collection_inputs()
    for r in (0.6, 0.8):
        txt = l0.match_collection_text(cdescs, cfeats, r)
        out[f"collection_text_r{r}"] = np.frombuffer(txt, dtype=np.uint8)
        print(f"synthetic collection r={r}: sha256 {hashlib.sha256(txt).hexdigest()[:16]} bytes {len(txt)}")
    out["ratio_sq_bits"] = np.array([np.float32(l0.square(0.6)).view(np.uint32), np.float32(l0.square(0.8)).view(np.uint32)], np.uint32)
    np.savez_compressed(os.path.join(HERE, "synth_golden.npz"), **out)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
