"""compute_matches driver (apps/compute_matches): flag spellings, file rules and the resume short-cut need no GPU."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "build", "compute_matches")


def _run(*args):
    return subprocess.run([EXE, *args], capture_output=True, text=True, timeout=120)


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(EXE):
        import __graft_entry__ as g
        g.build()
    assert os.path.exists(EXE)


def test_usage_errors():
    r = _run()
    assert r.returncode != 0 and "invalid output directory" in r.stderr          # compute_matches.cpp:77-80 wording
    r = _run("-x", "1")
    assert r.returncode != 0 and "Unrecognized option" in r.stderr               # cmd_line.h:182-188
    r = _run("-o", "/tmp", "-r", "0.8x")                                          # value must consume the whole token
    assert r.returncode != 0


def test_geometric_model_flag(tmp_path):
    """-g takes the first character, either case; an unknown model ends the run (compute_matches.cpp:99-116).  The essential
    model (the reference's default) is accepted but not built: the run stops after the putative stage with a message."""
    (tmp_path / "lists.txt").write_text("a.jpg;640;480\nb.jpg;640;480\n")
    (tmp_path / "matches.putative.txt").write_text("0 1\n0\n")
    r = _run("-i", str(tmp_path), "-o", str(tmp_path), "-g", "x")
    assert r.returncode != 0 and "Unknown geometric model" in r.stderr
    for spelling in ("e", "E", "essential"):
        r = _run("-i", str(tmp_path), "-o", str(tmp_path), "-g", spelling)
        assert r.returncode == 0 and "only -g f and -g h" in r.stdout
    r = _run("-i", str(tmp_path), "-o", str(tmp_path))           # default: e
    assert r.returncode == 0 and "--geometricModel e" in r.stdout and "only -g f and -g h" in r.stdout


def test_resume_rule_skips_matching(tmp_path):
    (tmp_path / "lists.txt").write_text("a.jpg;640;480\nb.jpg;640;480\n")
    (tmp_path / "matches.putative.txt").write_text("0 1\n0\n")
    for spelling in (["-r", "0.8"], ["-r0.8"], ["--distratio", "0.8"], ["--distratio=0.8"]):   # cmd_line.h:80-97
        r = _run("-i", str(tmp_path), "-o", str(tmp_path), *spelling)
        assert r.returncode == 0, r.stderr
        assert "PREVIOUS RESULTS LOADED" in r.stdout and "--distratio 0.8" in r.stdout
    assert (tmp_path / "matches.putative.txt").read_text() == "0 1\n0\n"


def test_missing_features_is_an_error(tmp_path):
    (tmp_path / "lists.txt").write_text("a.jpg;640;480\nb.jpg;640;480\n")
    r = _run("-i", str(tmp_path), "-o", str(tmp_path))
    assert r.returncode != 0 and "SIFT extraction" in r.stderr


def test_no_gpu_no_fallback(tmp_path, pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    io = pkg.io
    (tmp_path / "lists.txt").write_text("a.jpg;640;480\nb.jpg;640;480\n")
    for n in "ab":
        io.save_descs_bin(str(tmp_path / f"{n}.desc"), pkg.synth.uniform_set(1, 20))
        io.save_feats(str(tmp_path / f"{n}.feat"), pkg.synth.features(1, 0, 20))
    r = _run("-i", str(tmp_path), "-o", str(tmp_path), "-r", "0.8")
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    assert not (tmp_path / "matches.putative.txt").exists()
