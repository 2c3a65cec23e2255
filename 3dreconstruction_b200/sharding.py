"""Pair scheduler: shard the image-pair list across the GPUs of one box.

The reference matches pairs in one process, one after the other (matcher_all_in_memory.h:71-139; its only
parallelism is an optional OpenMP loop over j, :87-90).  Pairs are independent, so the multi-GPU design
has NO data-path collective: every GPU holds a replica of the descriptor arena (<= ~1 GB for the largest
BASELINE config vs 180 GB of HBM) and takes a contiguous, cost-balanced slice of the (i, j)-ordered pair
list -- contiguous so that consecutive work items on one GPU share the same db image i and its tiles stay
hot in the 126 MB L2.  Cost of a pair = rows_i * rows_j (the GEMM volume).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def pair_costs(pairs: np.ndarray, rows: Sequence[int]) -> np.ndarray:
    r = np.asarray(rows, dtype=np.int64)
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    return np.maximum(r[p[:, 0]] * r[p[:, 1]], 1)


def shard_bounds(costs: np.ndarray, world: int) -> List[int]:
    """world+1 indices b with shard k = [b[k], b[k+1]); prefix cost of b[k] is the first >= k/world of the total."""
    n = len(costs)
    if world < 1:
        raise ValueError("world must be >= 1")
    csum = np.concatenate([[0], np.cumsum(costs, dtype=np.int64)])
    total = int(csum[-1])
    b = [0]
    for k in range(1, world):
        target = (total * k + world - 1) // world
        b.append(int(np.searchsorted(csum, target, side="left")))
    b.append(n)
    for k in range(1, world + 1):  # monotone
        b[k] = max(b[k], b[k - 1])
    return b


def shard_pairs(pairs: np.ndarray, rows: Sequence[int], rank: int, world: int) -> Tuple[np.ndarray, Tuple[int, int]]:
    """This rank's slice of the pair list and its [begin, end) position in it."""
    pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    b = shard_bounds(pair_costs(pairs, rows), world)
    return pairs[b[rank]:b[rank + 1]], (b[rank], b[rank + 1])


def images_for_pairs_per_gpu(world: int, pairs_per_gpu: int = 4950) -> int:
    """Smallest collection size n with n(n-1)/2 >= world * pairs_per_gpu (weak scaling of the exhaustive job:
    100 images at 1 GPU, 142 at 2, 200 at 4, 282 at 8)."""
    n = 2
    while n * (n - 1) // 2 < world * pairs_per_gpu:
        n += 1
    return n
