"""Golden files of the step AFTER the hot path: the reference's own AC-RANSAC geometric filter
(oracle/_ref/libmvgref_geom.so, built by oracle/build_ref.sh from /root/reference) on the reference's real SIFT set data/et,
from the brute-force putatives this repo reproduces byte for byte (tests/golden/et_putative_r0.6.txt).  glibc rand() stream
pinned to srand(1) == the reference's never-seeded default.

    python tests/golden/make_golden_geometric.py      ->  tests/golden/et_matches_{f,h}.txt + et_geometric_golden.json
"""
import ctypes as C
import hashlib
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("MVG_REF", "/root/reference")
lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmvgref_geom.so"))
lib.ref_geometric_filter.restype = C.c_int
lib.ref_geometric_filter.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.c_char_p, C.c_char, C.c_double, C.c_uint, C.c_char_p]

names, sizes = [], []
for line in open(os.path.join(REF, "data", "et", "lists.txt")):
    f = line.strip().split(";")
    if len(f) >= 3:
        names.append(f[0])
        sizes += [int(f[1]), int(f[2])]
arr = (C.c_int * len(sizes))(*sizes)
meta = {"source": "data/et (.feat files of the reference) + tests/golden/et_putative_r0.6.txt", "max_residual": 4.0, "seed": 1,
        "iterations": 4096}
for model in ("f", "h"):
    out = os.path.join(GOLD, f"et_matches_{model}.txt")
    n = lib.ref_geometric_filter(os.path.join(REF, "data", "et").encode(), "\n".join(names).encode(), arr,
                                 os.path.join(GOLD, "et_putative_r0.6.txt").encode(), model.encode(), 4.0, 1, out.encode())
    data = open(out, "rb").read()
    tok = data.split()
    k, pairs, matches = 0, 0, 0
    while k + 3 <= len(tok):
        c = int(tok[k + 2]); pairs += 1; matches += c; k += 3 + 2 * c
    meta[model] = {"pairs": pairs, "matches": matches, "sha256": hashlib.sha256(data).hexdigest(), "returned": n}
    print(model, meta[model])
json.dump(meta, open(os.path.join(GOLD, "et_geometric_golden.json"), "w"), indent=1, sort_keys=True)
