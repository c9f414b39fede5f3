"""Executable specification of the exact pruning rules the CUDA kernels rely on, in plain NumPy float64 (the same
IEEE operations the device performs: one rounded subtraction, abs, compare; no FMA).  It re-states, at toy chunk
sizes, (a) the two-level k-NN search with the per-lane window walk of ``knn_scan_lane`` (eb2_kernels.cuh) and
(b) the bracketed marginal search of ``search_kernel`` (eb2_aux_kernels.cuh), and holds both to the brute-force
definitions of SURVEY.md Appendix A on inputs chosen to stress rounding: huge offsets, ties, duplicates, tiny scales.
The GPU parity tests check the kernels; this checks that the RULES are exact, on any machine."""
import numpy as np
import pytest

TC, SEED, GROUP = 16, 2, 4        # toy chunk length, seed half-width, slots per walk group (device: 2048, 8, 4)
SLACK = 2.0 ** -50
INF = float("inf")


def brute_eps(x, y, k):
    out = np.empty(len(x))
    for i in range(len(x)):
        d = np.maximum(np.abs(x[i] - x), np.abs(y[i] - y))
        out[i] = np.sort(d)[k]                        # (k+1)-th smallest, self included
    return out


def layout(x, y):
    """slots ordered by x across chunks of TC and by y inside each chunk (stable, as the device's partition)"""
    by_x = np.argsort(x, kind="stable")
    order = np.concatenate([c[np.argsort(y[c], kind="stable")] for c in (by_x[i:i + TC] for i in range(0, len(x), TC))])
    xs = x[by_x]
    nch = (len(x) + TC - 1) // TC
    cell_lo = np.array([xs[c * TC] for c in range(nch)])
    cell_hi = np.array([xs[min((c + 1) * TC, len(x)) - 1] for c in range(nch)])
    return order, cell_lo, cell_hi


def insert(best, v):
    best[-1] = v
    best.sort()


def walk(cx, cy, q0, q1, best, skip=(-1, -1)):
    """knn_scan_lane: binary search for the first slot with fl(q1 - y_s) < thr, then aligned groups of GROUP slots
    until fl(y_s - q1) >= thr at the first slot of a group; the exact test decides on every slot looked at"""
    n = len(cy)
    lo, hi, t0 = 0, n, best[-1]
    while lo < hi:
        mid = (lo + hi) >> 1
        if (q1 - cy[mid]) < t0:
            hi = mid
        else:
            lo = mid + 1
    s = lo - lo % GROUP
    while s < n:
        if not ((cy[s] - q1) < best[-1]):
            break
        for u in range(s, min(s + GROUP, n)):
            if skip[0] <= u < skip[1]:
                continue
            if abs(q0 - cx[u]) < best[-1] and abs(q1 - cy[u]) < best[-1]:
                insert(best, max(abs(q0 - cx[u]), abs(q1 - cy[u])))
        s += GROUP


def two_level_eps(x, y, k):
    order, cell_lo, cell_hi = layout(x, y)
    px, py = x[order], y[order]
    nch = len(cell_lo)
    eps = np.empty(len(x))
    for slot in range(len(x)):
        q0, q1 = px[slot], py[slot]
        best = np.full(k + 1, INF)
        home = slot // TC
        a, b = home * TC, min((home + 1) * TC, len(x))
        so = slot - a
        sa, sb = max(so - SEED, 0), min(so + SEED + 1, b - a)
        for u in range(sa, sb):                                     # seeds: no window, thr starts at +inf
            if abs(q0 - px[a + u]) < best[-1] and abs(q1 - py[a + u]) < best[-1]:
                insert(best, max(abs(q0 - px[a + u]), abs(q1 - py[a + u])))
        walk(px[a:b], py[a:b], q0, q1, best, (sa, sb))
        for j in range(home + 1, nch):                              # outwards: stop at the first chunk that cannot matter
            if (cell_lo[j] - q0) >= best[-1]:
                break
            walk(px[j * TC:(j + 1) * TC], py[j * TC:(j + 1) * TC], q0, q1, best)
        for j in range(home - 1, -1, -1):
            if (q0 - cell_hi[j]) >= best[-1]:
                break
            walk(px[j * TC:(j + 1) * TC], py[j * TC:(j + 1) * TC], q0, q1, best)
        eps[order[slot]] = best[k]
    return eps


def datasets():
    rng = np.random.default_rng(3)
    n = 150
    yield "gauss", rng.normal(size=n), rng.normal(size=n)
    yield "huge_offset", 1e15 + rng.integers(0, 40, n).astype(float), -3e14 + rng.integers(0, 40, n) * 0.125
    yield "ties_plus_noise", rng.integers(0, 6, n) + rng.normal(0, 1e-10, n), rng.integers(0, 6, n) + rng.normal(0, 1e-10, n)
    yield "duplicates", rng.integers(0, 4, n).astype(float), rng.integers(0, 4, n).astype(float)
    yield "tiny_scale", rng.normal(size=n) * 1e-300, rng.normal(size=n) * 1e-300
    yield "mixed_magnitudes", rng.normal(size=n) * 10.0 ** rng.integers(-8, 8, n), rng.standard_cauchy(n)
    yield "sorted_lattice", np.arange(n) * 0.1, (np.arange(n) % 7) * 0.1


@pytest.mark.parametrize("k", [1, 3, 7])
def test_two_level_lane_walk_is_exact(k):
    for name, x, y in datasets():
        assert np.array_equal(two_level_eps(x, y, k), brute_eps(x, y, k)), name


def test_bracketed_marginal_search_is_exact():
    """search_kernel: counts #{j : |x_i - s_j| <= r_i} from two bisections with the exact rounded predicates, run
    inside a bracket [min x - max r, max x + max r] (widened by 2^-50 relative) shared by a group of queries —
    identical to the all-pairs count, including r < 0 (eps = 0: count 0) and r = inf (count = all)."""
    rng = np.random.default_rng(5)
    for name, x, _ in datasets():
        s = np.sort(x)
        r = np.abs(rng.normal(size=len(x))) * (np.ptp(x) / 10 + 1e-300)
        r[::11] = -1e-12
        r[5::17] = INF
        r[3::13] = 0.0
        want = np.array([np.sum(np.abs(x[i] - x) <= r[i]) for i in range(len(x))])
        got = np.empty(len(x), dtype=np.int64)
        for g0 in range(0, len(x), 8):                               # a "warp" of 8 neighbouring queries
            idx = np.arange(g0, min(g0 + 8, len(x)))
            xmin, xmax, rmax = x[idx].min(), x[idx].max(), max(r[idx].max(), 0.0)
            lo_v = (xmin - rmax) - (abs(xmin) + rmax) * SLACK
            hi_v = (xmax + rmax) + (abs(xmax) + rmax) * SLACK
            wl = int(np.searchsorted(s, lo_v, side="left"))
            wh = int(np.searchsorted(s, hi_v, side="right"))
            for i in idx:
                first, fhi = wl, wh
                while first < fhi:
                    mid = (first + fhi) >> 1
                    if (x[i] - s[mid]) <= r[i]:
                        fhi = mid
                    else:
                        first = mid + 1
                lo, hi = wl, wh
                while lo < hi:
                    mid = (lo + hi) >> 1
                    if (s[mid] - x[i]) > r[i]:
                        hi = mid
                    else:
                        lo = mid + 1
                got[i] = max(0, lo - first)
        assert np.array_equal(got, want), name
