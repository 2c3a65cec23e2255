"""3dreconstruction_b200 -- B200-native exhaustive putative matching (u8 SIFT-128, BF-L2 + ratio test).

The package name starts with a digit, so import it with
``importlib.import_module("3dreconstruction_b200")`` (see ``__graft_entry__.py``).
Only the hot path lives here: ``csrc/`` (CUDA kernels + C ABI -> ``libmvgcuda.so``), ``mvgcuda``
(ctypes binding + mirrors of the reference's matcher interfaces), ``io`` (.desc/.feat/matches
file formats) and ``synth`` (seeded synthetic SIFT-like collections for tests and bench).
"""
from . import io, synth  # noqa: F401
from .mvgcuda import (  # noqa: F401
    ABI, DIM, TIE_LOWEST_INDEX, TIE_REFERENCE, ArrayMatcherCuda, Context, MatcherCudaAllInMemory,
    MvgCudaError, PairMatches, ResidentDb, load_library, pairs_exhaustive, square_f32, upload_friendly_order, write_matches,
)
