"""CPU model of the kernel's pruning rules under adversarial schedules.

The CUDA epilogue never sees all rows of a query: four warps ("parts") scan disjoint column slices concurrently, share a
bound through one shared-memory word, skip whole chunks / groups whose best possible row cannot beat the bound, and --
for queries that currently fail the ratio test -- admit only rows within rho * d(best).  Records that pass the ratio
test but could have lost a "blocker" are flagged and matched again exactly.  This file restates exactly that decision
logic (same fp32 expressions, same rounding directions) on plain distance arrays and checks, over many random inputs,
part interleavings, stale reads of the shared bound and loose chunk lower bounds, the two properties the proofs claim:

  * the final pass/fail decision of every query equals the brute-force decision, and
  * for every passing query the reported nearest row is the brute-force one (lowest index on ties).

It needs no GPU and no oracle: brute force here is three lines of numpy.  The GPU tests check the CUDA code against the
oracle; this test checks the RULE itself, including interleavings a real run only produces by chance."""
import numpy as np
import pytest

F32 = np.float32
BIG = 1 << 30


def ratio_pass(d1: int, d2: int, r: F32) -> bool:
    """matching_filters.h:44 -- float(d1) < ratio * float(d2), one fp32 multiply, strict."""
    return bool(F32(d1) < F32(r * F32(d2)))


def mul_ru_ceil(rho: F32, d: int) -> int:
    """__float2int_ru(__fmul_ru(rho, float(d))): product rounded UP to fp32, then ceil."""
    exact = float(rho) * float(F32(d))          # 24-bit x 24-bit mantissas: exact in double
    f = F32(exact)
    if float(f) < exact:
        f = np.nextafter(f, F32(np.inf))
    return int(np.ceil(float(f)))


def mul_rd_floor(rho: F32, d: int) -> int:
    """__float2int_rd(__fmul_rd(rho, float(d)))."""
    exact = float(rho) * float(F32(d))
    f = F32(exact)
    if float(f) > exact:
        f = np.nextafter(f, F32(-np.inf))
    return int(np.floor(float(f)))


def brute_force(d: np.ndarray, r: F32):
    order = np.lexsort((np.arange(len(d)), d))          # by (distance, row)
    i1, i2 = int(order[0]), int(order[1])
    return i1, ratio_pass(int(d[i1]), int(d[i2]), r)


class Part:
    """One epilogue warp's view of one query: running top-2 by (d, row) of the rows it has examined exactly."""

    def __init__(self):
        self.top = []          # sorted list of (d, row), at most 2
        self.T = BIG           # local copy of the bound (admit d <= T)

    def insert(self, d, row):
        self.top.append((d, row))
        self.top.sort()
        del self.top[2:]

    def new_bound(self, r: F32, rho: F32) -> int:
        """The bound this part publishes after an exact step (epi_exact16)."""
        if len(self.top) < 2:
            return BIG
        (c1, _), (c2, _) = self.top
        if ratio_pass(c1, c2, r):
            return c2                                   # a passing query keeps its exact 2nd neighbour
        return mul_ru_ceil(rho, c1)                     # failing: admit only d <= rho * d(best), rounded up


def run_query(d: np.ndarray, r: F32, rho: F32, rng, n_parts=4, chunk=16, group=4):
    """Returns (idx1, d1, d2) as the kernel would record them for one query, under a random schedule."""
    n = len(d)
    # column slices of the parts: interleaved blocks, like (column half) x (tile parity)
    block = chunk * int(rng.integers(1, 5))
    owner = (np.arange(n) // block) % n_parts
    parts = [Part() for _ in range(n_parts)]
    work = []
    for p in range(n_parts):
        cols = np.flatnonzero(owner == p)
        work.append([cols[k:k + chunk] for k in range(0, len(cols), chunk)])
    cursor = [0] * n_parts
    shared = BIG
    alive = [p for p in range(n_parts) if work[p]]
    while alive:
        p = alive[int(rng.integers(len(alive)))]        # arbitrary interleaving of the warps
        P = parts[p]
        cols = work[p][cursor[p]]
        cursor[p] += 1
        if cursor[p] == len(work[p]):
            alive.remove(p)
        if rng.random() < 0.6:                          # the shared word is re-read only now and then (stale otherwise)
            P.T = min(P.T, shared)
        # chunk filter on a LOWER bound of the chunk's best row (min norm of the chunk - 2 * max dot product)
        slack = int(rng.integers(0, 6)) if rng.random() < 0.5 else 0
        if int(d[cols].min()) - slack > P.T:
            continue
        touched = False
        for g0 in range(0, len(cols), group):           # groups of 4 rows, same kind of test
            gc = cols[g0:g0 + group]
            if int(d[gc].min()) - slack > P.T:
                continue
            for c in gc:                                # every row of an admitted group is examined exactly
                P.insert(int(d[c]), int(c))
            touched = True
        if touched:
            P.T = min(P.T, P.new_bound(r, rho))
            shared = min(shared, P.T)                   # red.shared.min
    merged = sorted(t for P in parts for t in P.top)[:2]
    (d1, i1), (d2, _) = merged
    return i1, d1, d2


def final_decision(d, r, rho, rec):
    """K3 + the second pass: (idx1, passes) the library would export for this query."""
    i1, d1, d2 = rec
    if not ratio_pass(d1, d2, r):
        return i1, False, False
    if ratio_pass(d1, mul_rd_floor(rho, d2), r):
        return i1, True, False
    bi1, bpass = brute_force(d, r)                      # ambiguous: matched again exactly
    return bi1, bpass, True


def make_distances(rng, n, kind):
    if kind == 0:      # concentrated "typical" distances plus a few near rows on both sides of the ratio threshold
        d = rng.normal(100000, 9000, n).astype(np.int64).clip(1)
        best = int(d.min())
        for _ in range(int(rng.integers(0, 6))):
            d[int(rng.integers(n))] = max(1, int(best * rng.uniform(0.2, 1.05)))
    elif kind == 1:    # tiny range: ties everywhere
        d = rng.integers(0, 12, n).astype(np.int64)
    elif kind == 2:    # geometric ladder: every new best improves by a factor around the ratio / rho
        d = rng.integers(50000, 60000, n).astype(np.int64)
        v = 50000.0
        for pos in np.sort(rng.choice(n, min(n, 12), replace=False)):
            v *= rng.uniform(0.55, 0.98)
            d[pos] = max(1, int(v))
    else:              # uniform
        d = rng.integers(1, 1 << 23, n).astype(np.int64)
    return d


@pytest.mark.parametrize("ratio,rho", [(0.8, 0.8), (0.8, 1.0), (0.6, 0.8), (0.8, 0.65), (0.95, 0.95), (1.0, 1.0)])
def test_pruning_rule_is_exact_under_any_schedule(ratio, rho):
    r = F32(F32(ratio) * F32(ratio))                    # Square(float), numeric.h:108-111
    rho_eff = F32(1.0) if not (F32(rho) < F32(1.0)) else F32(min(1.0, max(float(F32(rho)), float(F32(r * F32(1.002))))))
    rng = np.random.default_rng(int(ratio * 1000) * 7 + int(rho * 1000))
    n_amb = n_pass = 0
    for case in range(1500):
        n = int(rng.integers(2, 400))
        d = make_distances(rng, n, case % 4)
        want_i1, want_pass = brute_force(d, r)
        rec = run_query(d, r, rho_eff, rng)
        got_i1, got_pass, amb = final_decision(d, r, rho_eff, rec)
        n_amb += amb
        n_pass += want_pass
        assert got_pass == want_pass, (case, d.tolist(), rec)
        if want_pass:
            assert got_i1 == want_i1, (case, d.tolist(), rec)
        if float(rho_eff) >= 1.0 or want_pass:
            assert rec[1] == int(d.min()) or amb         # rho = 1: d1 is exact for every query; rho < 1: for every passing one
    assert n_pass > 20                                  # both outcomes were exercised
    if rho >= 1.0:
        assert n_amb == 0                               # rho = 1: nothing is ever matched twice


def test_unsound_rule_is_caught():
    """Sanity of the model: a bound of ratio * d(best) for failing queries WITHOUT the second pass (the tempting
    shortcut) must produce wrong decisions on the same inputs -- otherwise the test above proves nothing."""
    r = F32(F32(0.8) * F32(0.8))
    rng = np.random.default_rng(1)
    wrong = 0
    for case in range(1500):
        n = int(rng.integers(2, 400))
        d = make_distances(rng, n, case % 4)
        want_i1, want_pass = brute_force(d, r)
        i1, d1, d2 = run_query(d, r, F32(0.64), rng)    # rho = ratio, and no ambiguity check:
        wrong += ratio_pass(d1, d2, r) != want_pass
    assert wrong > 0
