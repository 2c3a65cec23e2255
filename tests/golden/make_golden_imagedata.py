"""BASELINE.json configs[0]: "apps/compute_matches on bundled data/imageData images: SIFT regions, brute-force L2 + ratio
0.8, CPU reference".  Generates tests/golden/imagedata_collection.npz + imagedata_golden.json by running the
reference's own SIFT wrapper (oracle/_ref/libmvgsift.so, oracle/build_sift_ref.sh) and the reference's own
brute-force collection matcher (oracle L0) here, where /root/reference is mounted.

Image decode: PIL instead of the reference's vendored libjpeg/libpng, then the reference's grey formula
uchar(0.3 R + 0.59 G + 0.11 B) (libs/image/include/mvg/image/image_converter.h:27-31).  Detector call as in
apps/compute_matches/compute_matches.cpp:209-211: SIFTDetector(img, feats, descs, is_zoom=false, root_sift=true, 0.04f).
Feature coordinates are stored after the .feat text round trip (6 significant digits), as compute_matches reads them.
"""
import ctypes as C
import hashlib
import importlib
import json
import os
import subprocess
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

pkg_io = importlib.import_module("3dreconstruction_b200.io")
REF = os.environ.get("MVG_REF", "/root/reference")
PAIRS = {
    "sceaux": ["data/imageData/SceauxCastle/100_7101.jpg", "data/imageData/SceauxCastle/100_7102.jpg"],
    "ace": ["data/imageData/StanfordMobileVisualSearch/Ace_0.png", "data/imageData/StanfordMobileVisualSearch/Ace_1.png"],
}


def grey(path):
    rgb = np.asarray(Image.open(path).convert("RGB"), dtype=np.float64)
    return (0.3 * rgb[..., 0] + 0.59 * rgb[..., 1] + 0.11 * rgb[..., 2]).astype(np.uint8)


def main():
    lib_path = os.path.join(ROOT, "oracle", "_ref", "libmvgsift.so")
    if not os.path.exists(lib_path):
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_sift_ref.sh")])
    lib = C.CDLL(lib_path)
    lib.ref_sift_u8.restype = C.c_int
    lib.ref_sift_u8.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_float),
                                C.POINTER(C.c_uint8), C.c_int]
    oracle.build_l0()
    l0 = oracle.L0()
    arrays, meta = {}, {}
    for name, files in PAIRS.items():
        descs, feats = [], []
        for k, f in enumerate(files):
            g = np.ascontiguousarray(grey(os.path.join(REF, f)))
            cap = 20000
            fo = np.zeros((cap, 4), np.float32)
            do = np.zeros((cap, 128), np.uint8)
            n = lib.ref_sift_u8(g.ctypes.data_as(C.POINTER(C.c_uint8)), g.shape[1], g.shape[0], 0, 1, C.c_float(0.04),
                                fo.ctypes.data_as(C.POINTER(C.c_float)), do.ctypes.data_as(C.POINTER(C.c_uint8)), cap)
            assert 0 < n <= cap
            ftxt = np.array([[float("%g" % v) for v in r] for r in fo[:n]], np.float32)   # .feat text round trip
            descs.append(do[:n].copy())
            feats.append(ftxt)
            arrays[f"{name}_desc_{k}"] = descs[-1]
            arrays[f"{name}_feat_{k}"] = feats[-1]
        meta[name] = {"files": files, "rows": [int(len(d)) for d in descs]}
        for r in (0.6, 0.8):
            txt = l0.match_collection_text(descs, feats, r)
            pw = pkg_io.matches_from_text(txt.decode())
            arrays[f"{name}_text_r{r}"] = np.frombuffer(txt, np.uint8)
            meta[name][f"r{r}"] = {"sha256": hashlib.sha256(txt).hexdigest(), "matches": int(sum(len(v) for v in pw.values()))}
        print(name, meta[name])
    np.savez_compressed(os.path.join(HERE, "imagedata_collection.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "imagedata_golden.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
