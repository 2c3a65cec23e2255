"""Debug aid: GPU vs reference geometric filter over subsets of the synthetic planar collection (model h)."""
import ctypes as C, importlib, json, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dreconstruction_b200")
io = importlib.import_module("3dreconstruction_b200.io")
GOLD = os.path.join(ROOT, "tests", "golden")
model = sys.argv[1] if len(sys.argv) > 1 else "h"
stem = "geo_synth_h" if model == "h" else "geo_synth"
z = np.load(os.path.join(GOLD, stem + ".npz"))
sizes = z["sizes"]; feats = [z[f"feat_{k}"] for k in range(len(sizes))]
put = io.matches_from_text(open(os.path.join(GOLD, stem + "_putative.txt")).read())
keys = sorted(put)
lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmvgref_geom.so"))
lib.ref_geometric_filter.restype = C.c_int
lib.ref_geometric_filter.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.c_char_p, C.c_char, C.c_double, C.c_uint, C.c_char_p]
ctx = pkg.Context(0)
ctx.upload_images([np.zeros((len(f), 128), np.uint8) for f in feats]); ctx.set_features(feats)
td = tempfile.mkdtemp()
names = [f"im{k}.jpg" for k in range(len(sizes))]
for k, f in enumerate(feats):
    io.save_feats(os.path.join(td, f"im{k}.feat"), np.concatenate([f, np.ones((len(f), 2), np.float32)], 1))
csz = (C.c_int * (2 * len(sizes)))(*[int(v) for wh in sizes for v in wh])

def run(sub):
    d = {k: put[k] for k in sub}
    p = os.path.join(td, "put.txt"); open(p, "w").write(io.matches_to_text(d))
    o = os.path.join(td, "out.txt")
    lib.ref_geometric_filter(td.encode(), "\n".join(names).encode(), csz, p.encode(), model.encode(), 4.0, 1, o.encode())
    want = io.matches_from_text(open(o).read())
    lib.ref_rand_position.restype = C.c_long
    lib.ref_rand_position.argtypes = [C.c_uint, C.c_long]
    ref_pos = lib.ref_rand_position(1, 4096 * 7 * len(sub) + 16)
    res = ctx.geometric_filter(pkg.PairMatches.from_dict(d), sizes, model=model)
    got = {tuple(k): res.pair(i) for i, k in enumerate(res.pairs.tolist()) if res.counts[i] > 0}
    line = []
    for k in sub:
        w = np.asarray(want.get(k, np.zeros((0, 2))), np.int64).reshape(-1, 2); g = np.asarray(got.get(k, np.zeros((0, 2))), np.int64).reshape(-1, 2)
        line.append(f"{k}:{len(put[k])}:{len(w)}/{len(g)}{'' if np.array_equal(w, g) else ' DIFF'}")
    print(len(sub), "pairs, rand", res.timing["rand_consumed"], "ref", ref_pos, "waves/launches", res.timing.get("total_launches"), " ".join(line), flush=True)

for K in range(1, 13):
    run(keys[:K])
run([(0, 5)]); run([(0, 6)]); run([(0, 5), (1, 2)]); run([(0, 6), (1, 2)]); run([(0, 1), (1, 2)]); run([(0, 5), (0, 6)])
