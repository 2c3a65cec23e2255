import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def pkg():
    mod = importlib.import_module("3dreconstruction_b200")
    if not os.path.exists(mod.mvgcuda.LIB_PATH):  # fresh checkout: compile libmvgcuda.so (nvcc cross-compiles without a GPU)
        import __graft_entry__ as g
        g.build()
    return mod


@pytest.fixture(scope="session")
def l1():
    from oracle import oracle
    return oracle.L1()


@pytest.fixture(scope="session")
def l0():
    """The reference's own code, when oracle/_ref was built (here) or travelled (GPU box)."""
    from oracle import oracle
    if os.path.isdir(os.path.join(oracle.REF_ROOT, "libs", "feature", "include")):
        oracle.build_l0()
    if not oracle.have_l0():
        pytest.skip("oracle/_ref not built (reference tree not visible)")
    return oracle.L0()


@pytest.fixture(scope="session")
def et():
    z = np.load(os.path.join(GOLDEN, "et_collection.npz"))
    descs = [z[f"desc_{k}"] for k in range(9)]
    feats = [z[f"feat_{k}"] for k in range(9)]
    return descs, feats


@pytest.fixture(scope="session")
def synth_golden():
    return np.load(os.path.join(GOLDEN, "synth_golden.npz"))


@pytest.fixture(scope="session")
def ctx(pkg):
    """One libmvgcuda context on cuda:0 for the whole GPU session.  Fails loudly without a GPU."""
    c = pkg.Context(0)
    yield c
    c.close()
