"""BASELINE configs[3] FILE -> FILE through the C++ driver: write NIMG x ROWS synthetic .desc/.feat files (config-4 generator,
all host cores), run build/compute_matches --gpus G on the directory, print the driver's own phase times (start-up / load /
match + format / export) and the sha256 + size of matches.putative.txt.  With CHECK_GPUS=1 the same directory is matched
again on one GPU and the two files must be byte-identical.  Log kept under profiles/.

    NIMG=1000 ROWS=8000 GPUS=8 python tests/tools/config4_files.py
    NIMG=350 ROWS=8000 GPUS=1 GEO=f python tests/tools/config4_files.py     (putative stage + AC-RANSAC filter stage)
"""
import hashlib, importlib, multiprocessing as mp, os, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
EXE = os.path.join(ROOT, "build", "compute_matches")
N = int(os.environ.get("NIMG", "1000")); ROWS = int(os.environ.get("ROWS", "8000")); GPUS = os.environ.get("GPUS", "8")
CFG = 4
GEO = os.environ.get("GEO", "")   # "f" / "h": run the geometric filter stage too (compute_matches -g f)
_pool = None


def _write(args):
    k, td = args
    global _pool
    pkg = importlib.import_module("3dreconstruction_b200")
    if _pool is None:
        _pool = pkg.synth.scene_pool(CFG, ROWS)
    pkg.io.save_descs_bin(os.path.join(td, f"im{k:04d}.desc"), pkg.synth.image(CFG, k, ROWS, _pool))
    f = pkg.synth.features(CFG, k, ROWS)
    with open(os.path.join(td, f"im{k:04d}.feat"), "w") as fh:
        fh.write("\n".join(" ".join("%g" % v for v in r) for r in f.tolist()) + "\n")
    return k


def run(td, gpus):
    out = os.path.join(td, "matches.putative.txt")
    if os.path.exists(out):
        os.remove(out)
    t0 = time.time()
    r = subprocess.run([EXE, "-i", td, "-o", td, "-r", "0.8", "--gpus", str(gpus)] + (["-g", GEO] if GEO else []), capture_output=True, text=True, timeout=2400)
    wall = time.time() - t0
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:], flush=True)
        return None
    line = [l for l in r.stdout.splitlines() if l.startswith("start-up")][-1]
    h = hashlib.sha256()
    with open(out, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    print(f"[--gpus {gpus}] wall {wall:.2f} s (process start to exit); {line}; file {os.path.getsize(out) / 1e6:.1f} MB sha256 {h.hexdigest()[:16]}", flush=True)
    for l in r.stdout.splitlines():
        if l.startswith("geometric filter ("):
            print("   " + l, flush=True)
    return h.hexdigest()


if __name__ == "__main__":
    td = tempfile.mkdtemp(prefix="cfg4_", dir=os.environ.get("TMPDIR_BIG", "/dev/shm" if os.path.isdir("/dev/shm") else None))
    try:
        t0 = time.time()
        with mp.Pool(min(os.cpu_count() or 1, 64)) as pool:
            for _ in pool.imap_unordered(_write, [(k, td) for k in range(N)], chunksize=4):
                pass
        open(os.path.join(td, "lists.txt"), "w").write("".join(f"im{k:04d}.jpg;4000;3000\n" for k in range(N)))
        print(f"wrote {N} x {ROWS} .desc/.feat in {time.time() - t0:.1f} s on {os.cpu_count()} cores -> {td}", flush=True)
        ok = True
        shas = [run(td, g) for g in GPUS.split(",")]
        ok &= all(s is not None for s in shas)
        if os.environ.get("CHECK_GPUS"):
            s1 = run(td, int(os.environ["CHECK_GPUS"]))
            ok &= s1 is not None and all(s == s1 for s in shas)
            print("byte-identical across GPU counts" if ok else "OUTPUT DIFFERS ACROSS GPU COUNTS", flush=True)
        print("CONFIG4 FILES OK" if ok else "CONFIG4 FILES FAILED", flush=True)
    finally:
        shutil.rmtree(td, ignore_errors=True)
    sys.exit(0 if ok else 1)
