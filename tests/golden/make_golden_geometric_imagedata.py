"""Golden files for the GPU geometric filter on REAL images: the two bundled pairs of BASELINE configs[0]
(data/imageData: SceauxCastle, StanfordMobileVisualSearch/Ace) with the reference's own SIFT regions and brute-force
putatives (tests/golden/imagedata_collection.npz, made by make_golden_imagedata.py), filtered by the reference's OWN
ImageCollectionGeometricFilter + GeometricFilter_FMatrix_AC / GeometricFilter_HMatrix_AC (oracle/_ref/libmvgref_geom.so),
max residual 4 px, 4096 iterations, rand() == srand(1), for the putatives at ratio 0.8 and 0.6.  Real geometry, > 1,000
putatives in one pair (long inlier lists), multi-orientation keypoints (coincident coordinates).

    python tests/golden/make_golden_geometric_imagedata.py  ->  tests/golden/imagedata_geometric.npz + imagedata_geometric_golden.json
"""
import ctypes as C
import hashlib
import importlib
import json
import os
import sys
import tempfile

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
io = importlib.import_module("3dreconstruction_b200.io")
REF = os.environ.get("MVG_REF", "/root/reference")


def main():
    z = np.load(os.path.join(HERE, "imagedata_collection.npz"))
    meta_in = json.load(open(os.path.join(HERE, "imagedata_golden.json")))
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmvgref_geom.so"))
    lib.ref_geometric_filter.restype = C.c_int
    lib.ref_geometric_filter.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.c_char_p, C.c_char, C.c_double, C.c_uint, C.c_char_p]
    arrays, meta = {}, {}
    for name, info in meta_in.items():
        sizes = [Image.open(os.path.join(REF, f)).size for f in info["files"]]   # (width, height)
        meta[name] = {"sizes": [list(map(int, s)) for s in sizes], "cases": {}}
        with tempfile.TemporaryDirectory() as td:
            names = [f"im{k}.jpg" for k in range(2)]
            for k in range(2):
                io.save_feats(os.path.join(td, f"im{k}.feat"), z[f"{name}_feat_{k}"])
            csz = (C.c_int * 4)(*[int(v) for s in sizes for v in s])
            for r in ("0.8", "0.6"):
                put = os.path.join(td, f"put_{r}.txt")
                open(put, "wb").write(z[f"{name}_text_r{r}"].tobytes())
                for model in ("f", "h"):
                    out = os.path.join(td, f"out_{r}_{model}.txt")
                    kept = lib.ref_geometric_filter(td.encode(), "\n".join(names).encode(), csz, put.encode(), model.encode(), 4.0, 1, out.encode())
                    data = open(out, "rb").read()
                    got = io.matches_from_text(data.decode())
                    arrays[f"{name}_r{r}_{model}"] = np.frombuffer(data, np.uint8)
                    meta[name]["cases"][f"r{r}_{model}"] = {"pairs_kept": int(kept), "matches": int(sum(len(v) for v in got.values())),
                                                           "sha256": hashlib.sha256(data).hexdigest()}
        print(name, json.dumps(meta[name]))
    np.savez_compressed(os.path.join(HERE, "imagedata_geometric.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "imagedata_geometric_golden.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
