"""The C-ABI library loads and exports every symbol include/mvgcuda.h declares; without a GPU the product
fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mvgcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mvgcuda_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_bound(pkg):
    syms = _declared_symbols()
    assert len(syms) >= 15
    assert set(syms) == set(pkg.ABI.keys()), "binding table and header disagree"


def test_library_exports_every_symbol(pkg):
    lib = ctypes.CDLL(pkg.mvgcuda.LIB_PATH)
    for s in _declared_symbols():
        assert hasattr(lib, s), f"libmvgcuda.so does not export {s}"
    assert pkg.load_library().mvgcuda_version() >= 100


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert pkg.load_library().mvgcuda_device_count() == 0
    with pytest.raises(pkg.MvgCudaError, match="no CPU fallback"):
        pkg.Context(0)


def test_product_does_not_import_oracle():
    """Nothing under the package (or include/) may reference oracle/."""
    bad = []
    for base in ("3dreconstruction_b200", "include", "apps"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".txt", ".cmake")):
                    t = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|liboracle|libmvgref", t) and "TEST INFRASTRUCTURE" not in t:
                        if fn not in ("mvgcuda.py", "__init__.py") or re.search(r"(from|import)\s+oracle|liboracle|libmvgref", t):
                            bad.append(os.path.join(dp, fn))
    assert not bad, bad
