"""The estimator seam: same private names and call signatures as the reference's
``ennemi/_entropy_estimators.py`` (the functions imported at ``ennemi/_driver.py:18-21``), so that
code written against the reference's internals keeps working.

The five k-NN estimators hand their data to the CUDA library (``_native``); nothing here computes a
neighbour search on the CPU.  The three purely discrete (plug-in) estimators involve no neighbour
search at all and stay on the host, as in the reference (out of scope for CUDA).
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _native
from . import _devices

_MANY_CLASSES_ROSS = ("The discrete variable has relatively many unique values."
                      " Did you pass y and x in correct order?")
_MANY_CLASSES_COND = ("A discrete variable has relatively many unique values."
                      " Have you set marked the discrete variables in correct order?"
                      " If both X and Y are discrete, the conditioning variable cannot be continuous"
                      " (this limitation can be lifted in the future).")
_NONFINITE = "data must be finite, check for nan or inf values"   # what cKDTree(...) raises in the reference


def _coords(*groups) -> np.ndarray:
    """(d, n) dimension-major fp64 block for the C ABI; rejects non-finite data like cKDTree does."""
    block = _native.pack_coords(groups)
    if not np.isfinite(block).all():
        raise ValueError(_NONFINITE)
    return block


def _classes(y):
    """np.unique labels -> (int32 class id per row, number of classes, class sizes).

    Integer / boolean labels spanning a small range take a counting path (O(n), no sort) that yields
    exactly what ``np.unique(y, return_inverse=True, return_counts=True)`` does: classes numbered in
    ascending label order."""
    y = np.asarray(y)
    if y.ndim == 1 and y.dtype.kind in "iub" and y.size:
        lo, hi = int(y.min()), int(y.max())
        if hi - lo < (1 << 22):
            shifted = (y.astype(np.int64) - lo) if y.dtype.kind != "b" else y.astype(np.int64)
            if y.dtype.kind == "b":
                lo = 0
            hist = np.bincount(shifted, minlength=hi - lo + 1)
            present = hist > 0
            rank = np.cumsum(present) - 1                     # label value -> class id (ascending labels)
            return np.ascontiguousarray(rank[shifted], dtype=np.int32), int(present.sum()), hist[present]
    labels, inverse, sizes = np.unique(y, return_inverse=True, return_counts=True)
    return np.ascontiguousarray(np.ravel(inverse), dtype=np.int32), len(labels), sizes


# ---------------------------------------------------------------------------------------------
# k-NN estimators (CUDA)
# ---------------------------------------------------------------------------------------------
def _estimate_single_entropy(x, k: int = 3) -> float:
    """Kozachenko-Leonenko entropy in nats; replaces ``_entropy_estimators.py:21-42``.
    ``x`` is (n,) or (n, m), one row per observation."""
    coords = _coords(np.asarray(x))
    from . import distributed
    if distributed.row_sharding_enabled():
        return distributed.sharded_entropy(coords, k)
    return _native.entropy(coords, k, dev=_devices.current())


def _estimate_single_mi(x, y, k: int = 3) -> float:
    """KSG mutual information in nats; replaces ``_entropy_estimators.py:69-113``."""
    coords = _coords(np.asarray(x), np.asarray(y))
    from . import distributed
    if distributed.row_sharding_enabled():      # one process per GPU: shard the query rows, one all-reduce
        return distributed.sharded_ksg_mi(coords, k)
    return _native.ksg_mi(coords, k, dev=_devices.current())


def _estimate_conditional_mi(x, y, cond, k: int = 3) -> float:
    """Frenzel-Pompe conditional MI; replaces ``_entropy_estimators.py:116-156``."""
    coords = _coords(np.asarray(x), np.asarray(y), np.asarray(cond))
    from . import distributed
    if distributed.row_sharding_enabled():
        return distributed.sharded_cmi(coords, k)
    return _native.cmi(coords, k, dev=_devices.current())


def _estimate_semidiscrete_mi(x, y, k: int = 3) -> float:
    """Ross MI between continuous ``x`` and discrete ``y``; replaces ``_entropy_estimators.py:159-200``."""
    cls, ncls, _ = _classes(y)
    if ncls > len(cls) / 4:                                       # :179-181
        warnings.warn(_MANY_CLASSES_ROSS, UserWarning)
    return _native.ross_mi(_coords(np.asarray(x)), cls, ncls, k, dev=_devices.current())


def _estimate_conditional_semidiscrete_mi(x, y, cond, k: int = 3) -> float:
    """Conditional Ross MI; replaces ``_entropy_estimators.py:203-247``."""
    cls, ncls, _ = _classes(y)
    _verify_not_continuous(ncls, len(cls))                        # :217
    return _native.ross_cmi(_coords(np.asarray(x), np.asarray(cond)), cls, ncls, k, dev=_devices.current())


def _psi(x):
    """Digamma for non-negative integers with the reference's conventions (``:327-350``):
    a scalar ``+inf`` as soon as any entry is zero, the reference's expansion otherwise.
    Evaluated by the device kernel that the estimators use."""
    arr = np.asarray(x)
    if np.any(arr == 0):
        return np.asarray(np.inf)
    out = _native.psi(np.ascontiguousarray(arr, dtype=np.int64).ravel(), dev=_devices.current())
    return out.reshape(arr.shape)


# ---------------------------------------------------------------------------------------------
# plug-in estimators for all-discrete data (host; no neighbour search involved)
# ---------------------------------------------------------------------------------------------
def _verify_not_continuous(n_unique: int, n: int) -> None:
    if n_unique > n / 4:                                          # :249-254
        warnings.warn(_MANY_CLASSES_COND, UserWarning)


def _assert_not_object(a: np.ndarray) -> None:
    if a.dtype.kind == "O":                                       # :58-66
        raise TypeError("Data type 'object' is not supported."
                        " Please pass only numeric, boolean, or string data."
                        " If your data is in a pandas DataFrame, convert string categories"
                        " to integers (pandas stores strings as objects).")


def _estimate_discrete_entropy(x) -> float:
    """-sum p log p over the distinct rows of ``x`` (``:44-56``)."""
    x = np.asarray(x)
    _assert_not_object(x)
    _, counts = np.unique(x, axis=0, return_counts=True)
    p = counts / x.shape[0]
    return -np.sum(np.dot(p, np.log(p)))


def _estimate_discrete_mi(x, y) -> float:
    """sum_xy p(x,y) log(p(x,y) / (p(x) p(y))) from the contingency counts (``:257-289``)."""
    n = len(x)
    both = np.column_stack((x, y))          # one common dtype, so mixed str/int labels compare equal
    _assert_not_object(both)
    x_vals, x_cnt = np.unique(both[:, 0], return_counts=True)
    y_vals, y_cnt = np.unique(both[:, 1], return_counts=True)
    cells, cell_cnt = np.unique(both, axis=0, return_counts=True)
    _verify_not_continuous(len(x_vals), n)
    _verify_not_continuous(len(y_vals), n)
    wx = x_cnt[np.searchsorted(x_vals, cells[:, 0])]
    wy = y_cnt[np.searchsorted(y_vals, cells[:, 1])]
    terms = cell_cnt * np.log(n * cell_cnt / (wx * wy))
    return np.sum(terms) / n


def _estimate_conditional_discrete_mi(x, y, cond) -> float:
    """sum_z p(z) I(X;Y | Z=z) with a discrete condition (``:291-320``)."""
    n = len(x)
    cond = np.asarray(cond)
    _assert_not_object(cond)
    _, which, sizes = np.unique(cond, axis=0, return_inverse=True, return_counts=True)
    which = np.ravel(which)
    x, y = np.asarray(x), np.asarray(y)
    weighted = np.zeros(len(sizes))
    for g in range(len(sizes)):
        sel = which == g
        weighted[g] = sizes[g] * _estimate_discrete_mi(x[sel], y[sel])
    return np.sum(weighted).item() / n
