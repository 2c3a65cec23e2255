"""File formats either side of the path (reference file:line relative to the reference tree).

  .desc  binary  [size_t n][n x 128 u8]     descriptor.h:160-203 (8-byte count on Linux x86-64; the
                                             shipped data/et files carry a 4-byte count, written by a
                                             32-bit Windows build -- both are accepted on load)
  .feat  text    "x y scale orientation\n"  feature.h:79-148 (ostream default: 6 significant digits)
  matches.*.txt  "i j\ncount\n_i _j\n..."   indexed_match_utils.h:22-73
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import numpy as np

DIM = 128


def load_descs_bin(path: str) -> np.ndarray:
    """LoadDescsFromBinFile (descriptor.h:160-181).  A missing file yields an empty set, as the
    reference does (its `!file.bad()` is true when open failed)."""
    if not os.path.exists(path):
        return np.zeros((0, DIM), np.uint8)
    raw = np.fromfile(path, dtype=np.uint8)
    for hdr in (8, 4):
        if raw.size >= hdr:
            n = int.from_bytes(raw[:hdr].tobytes(), "little")
            if raw.size == hdr + n * DIM:
                return raw[hdr:].reshape(n, DIM).copy()
    raise ValueError(f"{path}: not a [count][n x {DIM}] descriptor file")


def save_descs_bin(path: str, descs: np.ndarray, header_bytes: int = 8) -> None:
    """SaveDescsToBinFile (descriptor.h:185-203)."""
    d = np.ascontiguousarray(descs, dtype=np.uint8).reshape(-1, DIM)
    with open(path, "wb") as f:
        f.write(int(d.shape[0]).to_bytes(header_bytes, "little"))
        f.write(d.tobytes())


def load_feats(path: str) -> np.ndarray:
    """LoadFeatsFromFile<ScalePointFeature> (feature.h:117-133): whitespace-separated floats, 4 per feature."""
    if not os.path.exists(path):
        return np.zeros((0, 4), np.float32)
    with open(path, "r") as f:
        vals = np.array(f.read().split(), dtype=np.float32)
    return vals[: vals.size // 4 * 4].reshape(-1, 4)


def _fmt_g6(v: float) -> str:
    return "%g" % float(v)  # == ostream << float with default precision 6


def save_feats(path: str, feats: np.ndarray) -> None:
    """saveFeatsToFile (feature.h:137-148): "x y scale orientation" per line."""
    f4 = np.asarray(feats, dtype=np.float32).reshape(-1, 4)
    with open(path, "w") as f:
        for r in f4:
            f.write(" ".join(_fmt_g6(v) for v in r) + "\n")


def matches_to_text(pairwise: Dict[Tuple[int, int], np.ndarray]) -> str:
    """PairedIndexedMatchToStream (indexed_match_utils.h:22-38): std::map order = lexicographic (i, j)."""
    out = []
    for (i, j) in sorted(pairwise):
        m = np.asarray(pairwise[(i, j)]).reshape(-1, 2)
        out.append(f"{i} {j}\n{len(m)}\n")
        out.extend(f"{a} {b}\n" for a, b in m.tolist())
    return "".join(out)


def matches_from_text(text: str) -> Dict[Tuple[int, int], np.ndarray]:
    """pairedIndexedMatchImport (indexed_match_utils.h:48-73): blocks "i j count" + count x "_i _j" until the first token
    that does not parse; a key seen twice keeps its LAST block (operator[] assignment); a block cut short by the end of the
    file is kept with the missing entries left at IndexedMatch's default (0, 0), as `std::vector<IndexedMatch>(number)`
    followed by failing extractions leaves them."""
    tok = text.split()
    out: Dict[Tuple[int, int], np.ndarray] = {}
    k = 0
    while k + 3 <= len(tok):
        try:
            i, j, n = int(tok[k]), int(tok[k + 1]), int(tok[k + 2])
        except ValueError:
            break
        k += 3
        m = np.zeros((n, 2), np.int64)
        flat = []
        for t in tok[k:k + 2 * n]:
            try:
                flat.append(int(t))
            except ValueError:
                break
        # an odd number of tokens: the dangling _i was read, its _j was not
        m.reshape(-1)[:len(flat)] = flat
        k += 2 * n
        out[(i, j)] = m
    return out
