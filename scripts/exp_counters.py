import ctypes, importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dreconstruction_b200")
lib = pkg.load_library()
fn = lib.mvgcuda_debug_counters
fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
rows = int(os.environ.get("ROWS", "10000")); n_img = int(os.environ.get("NIMG", "8"))
ctx = pkg.Context(0)
descs = pkg.synth.collection(3, n_img, rows)
ctx.upload_images(descs)
pairs = pkg.pairs_exhaustive(n_img)
buf = (ctypes.c_ulonglong * 8)()
fn(buf, 1)
ctx.match_pairs(pairs, float(pkg.square_f32(0.8)), collect=False)
fn(buf, 1)
c = list(buf)
print(f"chunks {c[0]}  slow chunks {c[1]} ({100*c[1]/c[0]:.1f}%)  group hits {c[2]} ({c[2]/max(c[1],1):.2f} per slow chunk)  lane hits {c[3]} ({c[3]/max(c[1],1):.2f} per slow chunk, {c[3]/(c[0]*32)*100:.2f}% of lane-chunks)")
print(f"queries {c[4]}  tracked pass {c[5]} ({100*c[5]/max(c[4],1):.2f}%)  ambiguous (needs rescan) {c[6]} ({100*c[6]/max(c[4],1):.2f}% of queries)")
