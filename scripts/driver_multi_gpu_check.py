"""Multi-GPU check of the C++ driver (needs >= 2 GPUs): build/compute_matches with --gpus 1 and --gpus N on the same
directory of .feat/.desc files must write byte-identical matches.putative.txt (GPUs 1..N-1 work on replicas cloned from
GPU 0 over NVLink), equal to the Python API's export."""
import hashlib, importlib, os, shutil, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dreconstruction_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "compute_matches")
n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_img, rows = 14, 3000
ok = True
with tempfile.TemporaryDirectory() as td:
    names = []
    for k in range(n_img):
        r = rows - 37 * k
        pkg.io.save_descs_bin(os.path.join(td, f"im{k:03d}.desc"), pkg.synth.image(6, k, r, pkg.synth.scene_pool(6, rows)))
        pkg.io.save_feats(os.path.join(td, f"im{k:03d}.feat"), pkg.synth.features(6, k, r))
        names.append(f"im{k:03d}.jpg;4000;3000")
    open(os.path.join(td, "lists.txt"), "w").write("\n".join(names) + "\n")
    shas = {}
    for g in (1, n_gpus):
        out = os.path.join(td, "matches.putative.txt")
        if os.path.exists(out):
            os.remove(out)
        r = subprocess.run([EXE, "-i", td, "-o", td, "-r", "0.8", "--gpus", str(g)], capture_output=True, text=True, timeout=300)
        print(r.stdout.strip().splitlines()[-2] if r.returncode == 0 else r.stderr, flush=True)
        ok &= r.returncode == 0
        if r.returncode == 0:
            shas[g] = hashlib.sha256(open(out, "rb").read()).hexdigest()
    m = pkg.MatcherCudaAllInMemory(0.8, pkg.Context(0))
    ok &= m.LoadData([f"im{k:03d}.jpg" for k in range(n_img)], td)
    m.Match()
    ref = os.path.join(td, "ref.txt")
    m.Export(ref)
    shas["python"] = hashlib.sha256(open(ref, "rb").read()).hexdigest()
    print(shas)
    ok &= len(set(shas.values())) == 1
print("DRIVER MULTI-GPU OK" if ok else "DRIVER MULTI-GPU FAILED")
sys.exit(0 if ok else 1)
