"""Timing of the reference's CPU path for ``bench.py`` (``cpu_baseline`` and ``--impl reference``).
Test/measurement infrastructure only — never imported by the product.

The reference's ``_estimate_single_mi`` (``_entropy_estimators.py:100-110``) builds three cKDTrees on
all rows, then runs one k-NN query and two ball counts over all rows, on ONE thread (a single
estimate cannot be parallelised in the reference, ``benchmarks/bench_large_sample_mi.py:6-7``).
To keep a bench step bounded, the trees are built on all rows but only every ``stride``-th row is
queried; the query time is scaled by ``stride`` (queries are independent per row)."""
import time

import numpy as np


def ksg_mi_seconds(xs: np.ndarray, ys: np.ndarray, k: int = 3, stride: int = 20) -> float:
    """Extrapolated seconds of one full KSG estimate on the reference's SciPy path."""
    from scipy.spatial import cKDTree
    t0 = time.perf_counter()
    xy = np.column_stack((xs, ys))
    grid, gx, gy = cKDTree(xy), cKDTree(xs.reshape(-1, 1)), cKDTree(ys.reshape(-1, 1))
    t1 = time.perf_counter()
    q = slice(0, None, stride)
    eps = grid.query(xy[q], k=[k + 1], p=np.inf)[0].ravel()
    gx.query_ball_point(xs[q].reshape(-1, 1), eps - 1e-12, p=np.inf, return_length=True)
    gy.query_ball_point(ys[q].reshape(-1, 1), eps - 1e-12, p=np.inf, return_length=True)
    t2 = time.perf_counter()
    return (t1 - t0) + (t2 - t1) * stride
