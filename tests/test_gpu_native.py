"""The C++ host side on a GPU: the compute_matches driver (flags/files as the reference's CLI) and the header-only
adaptors compiled against the REFERENCE'S OWN headers (binary built where /root/reference is mounted)."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "build", "compute_matches")
ADAPT = os.path.join(ROOT, "build", "native", "test_adaptors")


def _write_et(pkg, et, d, hdr=8):
    descs, feats = et
    names = []
    for k, (dd, ff) in enumerate(zip(descs, feats)):
        names.append(f"et{k:03d}.jpg")
        pkg.io.save_descs_bin(str(d / f"et{k:03d}.desc"), dd, hdr)
        pkg.io.save_feats(str(d / f"et{k:03d}.feat"), ff)
    (d / "lists.txt").write_text("".join(f"{n};640;480;649.156;Canon;Canon PowerShot A10\n" for n in names))
    return names


@pytest.mark.parametrize("r,hdr", [("0.6", 4), ("0.8", 8)])
def test_compute_matches_driver_golden(pkg, et, tmp_path, r, hdr):
    _write_et(pkg, et, tmp_path, hdr)
    out = subprocess.run([EXE, "-i", str(tmp_path), "-o", str(tmp_path), "-r", r, "--gpus", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    got = (tmp_path / "matches.putative.txt").read_bytes()
    assert got == open(os.path.join(GOLDEN, f"et_putative_r{r}.txt"), "rb").read()
    # second invocation takes the resume short-cut
    out = subprocess.run([EXE, "-i", str(tmp_path), "-o", str(tmp_path), "-r", r], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "PREVIOUS RESULTS LOADED" in out.stdout


def test_adaptors_against_reference_classes(pkg, et, tmp_path):
    if not os.path.exists(ADAPT):
        pytest.skip("build/native/test_adaptors not built (needs the reference headers at build time)")
    names = _write_et(pkg, et, tmp_path, 8)
    out = subprocess.run([ADAPT, str(tmp_path), *names], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ADAPTORS OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_compute_matches_geometric_stage_golden(pkg, et, tmp_path):
    """`compute_matches -r 0.6 -g f` == the reference's two stages on data/et: matches.putative.txt and matches.f.txt
    (GeometricFilter_FMatrix_AC(4.0), never-seeded rand()) byte for byte; a second run imports the putatives
    (pairedIndexedMatchImport) and filters again -- same file."""
    _write_et(pkg, et, tmp_path, 4)
    want_f = open(os.path.join(GOLDEN, "et_matches_f.txt"), "rb").read()
    for second in (False, True):
        out = subprocess.run([EXE, "-i", str(tmp_path), "-o", str(tmp_path), "-r", "0.6", "-g", "f", "--gpus", "1"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr + out.stdout
        assert ("PREVIOUS RESULTS LOADED" in out.stdout) == second
        assert (tmp_path / "matches.putative.txt").read_bytes() == open(os.path.join(GOLDEN, "et_putative_r0.6.txt"), "rb").read()
        assert (tmp_path / "matches.f.txt").read_bytes() == want_f
        os.remove(tmp_path / "matches.f.txt")
    # -g h on the imported putatives: GeometricFilter_HMatrix_AC
    out = subprocess.run([EXE, "-i", str(tmp_path), "-o", str(tmp_path), "-r", "0.6", "-g", "h"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    assert (tmp_path / "matches.h.txt").read_bytes() == open(os.path.join(GOLDEN, "et_matches_h.txt"), "rb").read()
    # the essential model (the reference's default, needs K.txt) stops after the putative stage
    out = subprocess.run([EXE, "-i", str(tmp_path), "-o", str(tmp_path), "-r", "0.6", "-g", "e"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "only -g f and -g h" in out.stdout and not (tmp_path / "matches.e.txt").exists()


@pytest.mark.parametrize("name", ["sceaux", "ace"])
def test_compute_matches_on_the_bundled_image_pairs(pkg, tmp_path, name):
    """BASELINE configs[0] end to end through the driver: the reference's SIFT regions of a bundled data/imageData pair as
    .feat/.desc files + lists.txt with the real image sizes -> `compute_matches -r 0.8 -g f` and `-g h`: matches.putative.txt,
    matches.f.txt and matches.h.txt byte-identical to the reference's two stages (goldens: make_golden_imagedata.py,
    make_golden_geometric_imagedata.py)."""
    import json
    import numpy as np
    z = np.load(os.path.join(GOLDEN, "imagedata_collection.npz"))
    g = np.load(os.path.join(GOLDEN, "imagedata_geometric.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "imagedata_geometric_golden.json")))[name]
    lines = []
    for k in range(2):
        pkg.io.save_descs_bin(str(tmp_path / f"im{k}.desc"), z[f"{name}_desc_{k}"], 8)
        pkg.io.save_feats(str(tmp_path / f"im{k}.feat"), z[f"{name}_feat_{k}"])
        lines.append(f"im{k}.jpg;{meta['sizes'][k][0]};{meta['sizes'][k][1]}\n")
    (tmp_path / "lists.txt").write_text("".join(lines))
    for model in ("f", "h"):
        out = subprocess.run([EXE, "-i", str(tmp_path), "-o", str(tmp_path), "-r", "0.8", "-g", model, "--gpus", "1"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr + out.stdout
        assert (tmp_path / "matches.putative.txt").read_bytes() == z[f"{name}_text_r0.8"].tobytes()
        assert (tmp_path / f"matches.{model}.txt").read_bytes() == g[f"{name}_r0.8_{model}"].tobytes()
