"""The geometric filter on the synthetic planted-geometry golden set, meant to be run under compute-sanitizer
(memcheck / racecheck / synccheck); prints whether the output still equals the golden file."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dreconstruction_b200")
GOLD = os.path.join(ROOT, "tests", "golden")
MODEL = os.environ.get("MODEL", "f")
STEM = "geo_synth" if MODEL == "f" else "geo_synth_h"
z = np.load(os.path.join(GOLD, STEM + ".npz"))
sizes = z["sizes"]
feats = [z[f"feat_{k}"] for k in range(len(sizes))]
ctx = pkg.Context(0)
ctx.upload_images([np.zeros((len(f), 128), np.uint8) for f in feats])
ctx.set_features(feats)
put = pkg.PairMatches.from_dict(pkg.io.matches_from_text(open(os.path.join(GOLD, STEM + "_putative.txt")).read()))
keep = int(os.environ.get("PAIRS", "8"))
sub = pkg.PairMatches.from_dict({tuple(map(int, put.pairs[p])): put.pair(p) for p in range(keep)})
res = ctx.geometric_filter(sub, sizes, model=MODEL, iterations=int(os.environ.get("ITERS", "512")))
print("pairs", len(sub.pairs), "kept", int((res.counts > 0).sum()), "matches", int(res.counts.sum()), "rand", res.timing["rand_consumed"])
