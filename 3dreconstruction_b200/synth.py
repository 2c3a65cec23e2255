"""Seeded synthetic collections of SIFT-like uint8 descriptors (SURVEY.md 8(d) generator).

Rows are non-negative, ~6-11 % zeros, gamma-ish bins scaled so the row L2 norm is ~508 and the
maximum <= 255 (the statistics of the reference's RootSIFT output, sift.hpp:40-52, measured on
data/et), with planted correspondences: every image draws ~30 % of its rows as noisy copies of a
shared "scene" pool, so a realistic few hundred queries per pair pass the ratio test.
Also: a worst-case exactness set (uniform 0..255) and a tie set (tiny alphabet + duplicated rows).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

DIM = 128


def _sift_like(rng: np.random.Generator, n: int) -> np.ndarray:
    g = rng.gamma(shape=0.9, scale=1.0, size=(n, DIM)).astype(np.float32)
    g[rng.random((n, DIM)) < 0.08] = 0.0
    g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-6)
    g = np.minimum(g, 0.45)  # clip like SIFT's 0.2 rule (looser), renormalise
    g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-6)
    return np.clip(np.floor(512.0 * g), 0, 255).astype(np.uint8)


def scene_pool(config: int, rows: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(1000 * config + 999))
    return _sift_like(rng, 2 * rows)


def image(config: int, image_id: int, rows: int, pool: np.ndarray, shared: float = 0.3) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(1000 * config + image_id))
    d = _sift_like(rng, rows)
    k = int(rows * shared)
    if k and len(pool):
        src = rng.integers(0, len(pool), size=k)
        noise = rng.integers(-3, 4, size=(k, DIM))
        noisy = np.clip(pool[src].astype(np.int16) + noise, 0, 255).astype(np.uint8)
        where = rng.permutation(rows)[:k]
        d[where] = noisy
    return d


def collection(config: int, n_images: int, rows: int) -> List[np.ndarray]:
    pool = scene_pool(config, rows)
    return [image(config, i, rows, pool) for i in range(n_images)]


def features(config: int, image_id: int, rows: int, dup_frac: float = 0.02) -> np.ndarray:
    """[rows][4] x y scale orientation; x,y in a 4000x3000 frame rounded to 6 significant digits (the .feat
    text round-trip), with some deliberate duplicate x and duplicate (x,y) to exercise de-dup-2."""
    rng = np.random.Generator(np.random.PCG64(1000 * config + image_id + 500000))
    f = np.empty((rows, 4), np.float32)
    f[:, 0] = rng.uniform(0, 4000, rows)
    f[:, 1] = rng.uniform(0, 3000, rows)
    f[:, 2] = rng.uniform(1, 20, rows)
    f[:, 3] = rng.uniform(-3.14, 3.14, rows)
    k = int(rows * dup_frac)
    if k >= 2:
        a = rng.integers(0, rows, k)
        b = rng.integers(0, rows, k)
        f[a[: k // 2], 0] = f[b[: k // 2], 0]          # same x only
        f[a[k // 2:], :2] = f[b[k // 2:], :2]          # same (x, y)
    return np.array([[float("%g" % v) for v in r] for r in f], dtype=np.float32) if rows <= 4096 else \
        np.asarray(np.char.mod("%g", f).astype(np.float32))


def uniform_set(seed: int, rows: int) -> np.ndarray:
    return np.random.Generator(np.random.PCG64(seed)).integers(0, 256, size=(rows, DIM), dtype=np.uint8)


def tie_set(seed: int, rows: int, alphabet: int = 2, dup_frac: float = 0.25) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    d = rng.integers(0, alphabet, size=(rows, DIM), dtype=np.uint8)
    k = int(rows * dup_frac)
    if k:
        d[rng.integers(0, rows, k)] = d[rng.integers(0, rows, k)]
    return d
