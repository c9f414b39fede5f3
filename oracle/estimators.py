"""Plain CPU restatement of the reference's k-NN estimators (test oracle, not product code).

Reference algorithm: ``/root/reference/ennemi/_entropy_estimators.py`` (ennemi 1.5.0).
The arithmetic itself lives in a third-party dependency that is not vendored in the
reference tree: ``scipy.spatial.cKDTree`` (``scipy~=1.10`` in the reference's ``setup.py:49``;
SciPy 1.18.1 in this image).  Its semantics for the two calls on the path are restated here
(SURVEY.md Appendix A):

* ``query(pts, k=[k+1], p=inf)``  -> the (k+1)-th smallest Chebyshev distance from each
  point to the tree's points, the point itself included at distance 0; ``inf`` if the tree
  holds fewer than k+1 points.  Distance = max over dimensions of ``|a_d - b_d|`` with one
  rounded fp64 subtraction per dimension.
* ``query_ball_point(pts, r, p=inf, return_length=True)`` -> ``#{j : dist(i, j) <= r_i}``,
  inclusive, the point itself included when ``r_i >= 0``.

Three interchangeable backends compute those two primitives:

``"brute"``  numpy all-pairs restatement (this file), O(N^2), for N up to a few 10^4;
``"c"``      the same all-pairs loops in C (``oracle/oracle_c.c``, built by ``oracle/build.py``),
             OpenMP-threaded, for N up to a few 10^5;
``"scipy"``  the very SciPy calls the reference makes (the reference's CPU path), any N.

Every estimator returns a dict with the intermediate arrays (``eps``, neighbour counts)
and the final ``value`` so that parity tests can compare the CUDA path stage by stage.
"""
from __future__ import annotations

import ctypes
import os
import warnings

import numpy as np

BACKENDS = ("brute", "c", "scipy")

_RADIUS_SHRINK = 1e-12  # _entropy_estimators.py:109 — cKDTree is inclusive, the algorithm wants strict


# --------------------------------------------------------------------------------------
# digamma, _entropy_estimators.py:327-350
# --------------------------------------------------------------------------------------
def psi(n):
    """The reference's digamma for non-negative integers (``_psi``, :327-350).

    Any zero makes the whole result a scalar ``+inf`` (:338-339); ``psi(1)`` is the literal
    (:343); everything else is the asymptotic expansion at :348, evaluated with the same
    numpy ufuncs so the bits agree.
    """
    n = np.asarray(n)
    if np.any(n == 0):
        return np.asarray(np.inf)
    out = np.full(n.shape, -0.5772156649015331)
    big = n != 1
    v = np.asarray(n[big], dtype=np.float64)
    out[big] = np.log(v) - np.power(v, -6) * (
        np.power(v, 2) * (np.power(v, 2) * (v / 2 + 1 / 12) - 1 / 120) + 1 / 252)
    return out


# --------------------------------------------------------------------------------------
# the two k-NN primitives
# --------------------------------------------------------------------------------------
def _as_points(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    a = np.ascontiguousarray(a)
    if not np.all(np.isfinite(a)):
        # what cKDTree(...) raises for such input
        raise ValueError("data must be finite, check for nan or inf values")
    return a


def _chunk_rows(n_query: int, n_cand: int) -> int:
    return max(1, min(n_query, (1 << 24) // max(n_cand, 1)))


def _cheb_block(q: np.ndarray, c: np.ndarray) -> np.ndarray:
    """All-pairs max-norm distances between rows of q and rows of c (rounded subtraction)."""
    dist = np.abs(q[:, None, 0] - c[None, :, 0])
    for d in range(1, q.shape[1]):
        np.maximum(dist, np.abs(q[:, None, d] - c[None, :, d]), out=dist)
    return dist


def _kth_brute(cand: np.ndarray, query: np.ndarray, k: int) -> np.ndarray:
    nq, nc = len(query), len(cand)
    out = np.full(nq, np.inf)
    if nc < k + 1:
        return out
    step = _chunk_rows(nq, nc)
    for lo in range(0, nq, step):
        dist = _cheb_block(query[lo:lo + step], cand)
        out[lo:lo + step] = np.partition(dist, k, axis=1)[:, k]
    return out


def _count_brute(cand: np.ndarray, query: np.ndarray, radius: np.ndarray) -> np.ndarray:
    nq, nc = len(query), len(cand)
    out = np.zeros(nq, dtype=np.int64)
    step = _chunk_rows(nq, nc)
    for lo in range(0, nq, step):
        dist = _cheb_block(query[lo:lo + step], cand)
        out[lo:lo + step] = np.count_nonzero(dist <= radius[lo:lo + step, None], axis=1)
    return out


_clib = None


def _load_c():
    global _clib
    if _clib is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle_c.so")
        if not os.path.exists(path):
            from . import build as _b
            _b.build()
        lib = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        lp = ctypes.POINTER(ctypes.c_int64)
        lib.orc_kth_distance.argtypes = [dp, ctypes.c_int64, dp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, dp]
        lib.orc_kth_distance.restype = ctypes.c_int
        lib.orc_ball_count.argtypes = [dp, ctypes.c_int64, dp, ctypes.c_int64, ctypes.c_int, dp, lp]
        lib.orc_ball_count.restype = ctypes.c_int
        _clib = lib
    return _clib


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _kth_c(cand, query, k):
    lib = _load_c()
    out = np.empty(len(query))
    rc = lib.orc_kth_distance(_dptr(cand), len(cand), _dptr(query), len(query), cand.shape[1], k, _dptr(out))
    if rc:
        raise RuntimeError("orc_kth_distance failed")
    return out


def _count_c(cand, query, radius):
    lib = _load_c()
    radius = np.ascontiguousarray(radius, dtype=np.float64)
    out = np.empty(len(query), dtype=np.int64)
    rc = lib.orc_ball_count(_dptr(cand), len(cand), _dptr(query), len(query), cand.shape[1], _dptr(radius),
                            out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    if rc:
        raise RuntimeError("orc_ball_count failed")
    return out


def _kth_scipy(cand, query, k):
    from scipy.spatial import cKDTree
    return cKDTree(cand).query(query, k=[k + 1], p=np.inf)[0].ravel()


def _count_scipy(cand, query, radius):
    from scipy.spatial import cKDTree
    return np.asarray(cKDTree(cand).query_ball_point(query, radius, p=np.inf, return_length=True), dtype=np.int64)


_KTH = {"brute": _kth_brute, "c": _kth_c, "scipy": _kth_scipy}
_COUNT = {"brute": _count_brute, "c": _count_c, "scipy": _count_scipy}


def kth_distance(cand, k: int, query=None, backend: str = "brute") -> np.ndarray:
    """(k+1)-th smallest Chebyshev distance, self included — ``tree.query(q, k=[k+1], p=inf)``
    at ``_entropy_estimators.py:39,108,142,194,240``."""
    cand = _as_points(cand)
    query = cand if query is None else _as_points(query)
    return _KTH[backend](cand, query, k)


def ball_count(cand, radius, query=None, backend: str = "brute") -> np.ndarray:
    """``#{j: dist <= radius_i}`` — ``tree.query_ball_point(q, r, p=inf, return_length=True)``
    at ``_entropy_estimators.py:109-110,152-154,196,243-245``."""
    cand = _as_points(cand)
    query = cand if query is None else _as_points(query)
    return _COUNT[backend](cand, query, np.asarray(radius, dtype=np.float64))


# --------------------------------------------------------------------------------------
# the five estimators
# --------------------------------------------------------------------------------------
def ksg_mi(x, y, k: int = 3, backend: str = "brute") -> dict:
    """KSG algorithm 1, ``_estimate_single_mi`` (``_entropy_estimators.py:69-113``)."""
    xs, ys = _as_points(x), _as_points(y)
    n = len(xs)
    eps = kth_distance(np.column_stack((xs, ys)), k, backend=backend)          # :108
    rad = eps - _RADIUS_SHRINK
    nx = ball_count(xs, rad, backend=backend)                                  # :109
    ny = ball_count(ys, rad, backend=backend)                                  # :110
    value = psi(n) + psi(k) - np.mean(psi(nx) + psi(ny))                       # :113
    return {"eps": eps, "nx": nx, "ny": ny, "value": float(value)}


def conditional_mi(x, y, cond, k: int = 3, backend: str = "brute") -> dict:
    """Frenzel-Pompe, ``_estimate_conditional_mi`` (``_entropy_estimators.py:116-156``)."""
    xs, ys, zs = _as_points(x), _as_points(y), _as_points(cond)
    eps = kth_distance(np.column_stack((xs, ys, zs)), k, backend=backend)      # :142
    rad = eps - _RADIUS_SHRINK
    nxz = ball_count(np.column_stack((xs, zs)), rad, backend=backend)          # :152
    nyz = ball_count(np.column_stack((ys, zs)), rad, backend=backend)          # :153
    nz = ball_count(zs, rad, backend=backend)                                  # :154
    value = psi(k) - np.mean(psi(nxz) + psi(nyz) - psi(nz))                    # :156
    return {"eps": eps, "nxz": nxz, "nyz": nyz, "nz": nz, "value": float(value)}


def semidiscrete_mi(x, y, k: int = 3, backend: str = "brute") -> dict:
    """Ross, ``_estimate_semidiscrete_mi`` (``_entropy_estimators.py:159-200``).
    ``x`` continuous, ``y`` discrete (any dtype)."""
    xs = _as_points(x)
    y = np.asarray(y)
    n = len(xs)
    labels, sizes = np.unique(y, return_counts=True)                           # :177
    if len(labels) > n / 4:                                                    # :179-181
        warnings.warn("The discrete variable has relatively many unique values."
                      " Did you pass y and x in correct order?", UserWarning)
    eps = np.empty(n)
    n_full = np.empty(n)                                                       # fp64, as :191
    for lab in labels:
        sel = y == lab
        sub = xs[sel]
        e = kth_distance(sub, k, backend=backend)                              # :194 within the class
        eps[sel] = e
        n_full[sel] = ball_count(xs, e - _RADIUS_SHRINK, query=sub, backend=backend)   # :196 over all x
    weighted = np.sum(np.dot(psi(sizes), sizes / n))                           # :199
    value = psi(n) + psi(k) - np.mean(psi(n_full)) - weighted                  # :200
    return {"eps": eps, "n_full": n_full.astype(np.int64), "value": float(value)}


def conditional_semidiscrete_mi(x, y, cond, k: int = 3, backend: str = "brute") -> dict:
    """``_estimate_conditional_semidiscrete_mi`` (``_entropy_estimators.py:203-247``)."""
    xs, zs = _as_points(x), _as_points(cond)
    y = np.asarray(y)
    n = len(y)
    labels = np.unique(y)                                                      # :216
    if len(labels) > n / 4:                                                    # :249-254
        warnings.warn("A discrete variable has relatively many unique values."
                      " Have you set marked the discrete variables in correct order?"
                      " If both X and Y are discrete, the conditioning variable cannot be continuous"
                      " (this limitation can be lifted in the future).", UserWarning)
    xz = np.column_stack((xs, zs))
    eps = np.empty(n)
    nxz, nyz, nz = np.empty(n), np.empty(n), np.empty(n)
    for lab in labels:
        sel = y == lab
        e = kth_distance(xz[sel], k, backend=backend)                          # :240 within the class
        rad = e - _RADIUS_SHRINK
        eps[sel] = e
        nxz[sel] = ball_count(xz, rad, query=xz[sel], backend=backend)         # :243 all (x,z)
        nyz[sel] = ball_count(zs[sel], rad, backend=backend)                   # :244 z within the class
        nz[sel] = ball_count(zs, rad, query=zs[sel], backend=backend)          # :245 all z
    value = psi(k) - np.mean(psi(nxz) + psi(nyz) - psi(nz))                    # :247
    return {"eps": eps, "nxz": nxz.astype(np.int64), "nyz": nyz.astype(np.int64),
            "nz": nz.astype(np.int64), "value": float(value)}


def knn_entropy(x, k: int = 3, backend: str = "brute") -> dict:
    """Kozachenko-Leonenko, ``_estimate_single_entropy`` (``_entropy_estimators.py:21-42``)."""
    xs = _as_points(x)
    n, ndim = xs.shape
    dist = kth_distance(xs, k, backend=backend)                                # :39
    with np.errstate(divide="ignore"):
        value = psi(n) - psi(k) + ndim * (np.mean(np.log(dist)) + np.log(2))   # :42
    return {"dist": dist, "value": float(value)}
