"""Start-up cost of the C++ driver on this box: raw CUDA initialisation (build/startup_probe) next to build/compute_matches on
a tiny collection, repeated, with and without CUDA_VISIBLE_DEVICES restricted to one GPU.  Log kept under profiles/."""
import importlib, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dreconstruction_b200")
PROBE = os.path.join(ROOT, "build", "startup_probe")
if not os.path.exists(PROBE):  # nvcc is in the image (here and on the GPU box)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-o", PROBE, os.path.join(ROOT, "tests", "tools", "startup_probe.cu"), "-lcuda"])
for rep in range(3):
    print(subprocess.run([os.path.join(ROOT, "build", "startup_probe")], capture_output=True, text=True).stdout.strip(), flush=True)
print("CUDA_VISIBLE_DEVICES=0:", subprocess.run([os.path.join(ROOT, "build", "startup_probe")], capture_output=True, text=True,
                                              env={**os.environ, "CUDA_VISIBLE_DEVICES": "0"}).stdout.strip(), flush=True)
with tempfile.TemporaryDirectory() as td:
    names = []
    for k in range(4):
        pkg.io.save_descs_bin(f"{td}/im{k}.desc", pkg.synth.uniform_set(k, 300))
        pkg.io.save_feats(f"{td}/im{k}.feat", pkg.synth.features(1, k, 300))
        names.append(f"im{k}.jpg;640;480")
    open(f"{td}/lists.txt", "w").write("\n".join(names) + "\n")
    for env in ({}, {"CUDA_VISIBLE_DEVICES": "0"}):
        for rep in range(2):
            if os.path.exists(f"{td}/matches.putative.txt"):
                os.remove(f"{td}/matches.putative.txt")
            t0 = time.time()
            r = subprocess.run([os.path.join(ROOT, "build", "compute_matches"), "-i", td, "-o", td, "-r", "0.8", "--gpus", "1"], capture_output=True, text=True, env={**os.environ, **env})
            dt = time.time() - t0
            line = [l for l in r.stdout.splitlines() if l.startswith("start-up")]
            print(env, "run", rep, f"wall {dt:.2f} s;", (line[-1][:110] if line else r.stderr[-200:]), flush=True)
