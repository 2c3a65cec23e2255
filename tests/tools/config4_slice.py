"""BASELINE configs[3] at reduced image count (8,000 SIFT per image, exhaustive pairs): collection matcher over several
batches + text export, then (i) a random sample of pairs bit-exact against the oracle, (ii) the exported file round-trips
through the reference's own pairedIndexedMatchImport / PairedIndexedMatchToStream byte-identically (when oracle/_ref is
present) -- the file-boundary acceptance check of SURVEY.md 8(d) config 4.  Log kept under profiles/."""
import hashlib, importlib, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("3dreconstruction_b200")
from oracle import oracle

n_img = int(os.environ.get("NIMG", "300")); rows = 8000
t0 = time.time()
descs = pkg.synth.collection(4, n_img, rows)
feats = [pkg.synth.features(4, k, rows)[:, :2].copy() for k in range(n_img)]
print(f"generated {n_img} x {rows} in {time.time() - t0:.1f} s", flush=True)
ctx = pkg.Context(0)
m = pkg.MatcherCudaAllInMemory(0.8, ctx)
t0 = time.time(); m.LoadArrays(descs, feats); t_up = time.time() - t0
pairs = pkg.pairs_exhaustive(n_img)
rs = float(pkg.square_f32(0.8))
t0 = time.time(); res = ctx.match_collection(pairs, rs, 0); t_match = time.time() - t0
print(f"upload {t_up*1e3:.0f} ms; {len(pairs)} pairs matched in {t_match*1e3:.0f} ms wall ({len(pairs)/t_match:.0f} pairs/s end to end), "
      f"gpu {res.timing['gpu_ms']:.0f} ms in {res.timing['knn_kernel_launches']} knn launches, {int(res.offsets[-1])} putative matches", flush=True)
with tempfile.TemporaryDirectory() as td:
    out = os.path.join(td, "matches.putative.txt")
    t0 = time.time(); ctx.export_matches(pairs, out); t_exp = time.time() - t0
    size = os.path.getsize(out)
    sha = hashlib.sha256(open(out, "rb").read()).hexdigest()
    print(f"export {size/1e6:.1f} MB in {t_exp*1e3:.0f} ms, sha256 {sha[:16]}", flush=True)
    l1 = oracle.L1()
    rng = np.random.default_rng(0)
    sample = rng.choice(len(pairs), 24, replace=False)
    ok = True
    for p in sample:
        i, j = pairs[p]
        want = l1.dedup_xy(l1.pair_matches(descs[i], descs[j], rs), feats[i], feats[j])
        ok &= np.array_equal(res.pair(p), want)
    print(f"24 sampled pairs vs oracle (rows 7-13): {'bit-exact' if ok else 'MISMATCH'}", flush=True)
    if oracle.have_l0():
        back = os.path.join(td, "back.txt")
        t0 = time.time(); oracle.L0().roundtrip_matches(out, back)
        same = open(back, "rb").read() == open(out, "rb").read()
        print(f"reference import/export round trip ({time.time()-t0:.1f} s): {'byte-identical' if same else 'DIFFERENT'}", flush=True)
        ok &= same
print("CONFIG4 SLICE OK" if ok else "CONFIG4 SLICE FAILED")
