"""Host-side preparation of one (variable, lag) task: lag alignment, masking, NaN dropping,
validation and rescaling.  Semantics follow ``ennemi/_driver.py:788-812, 834-902`` exactly — the
buffers this module produces are "the same preprocessed inputs" the parity contract refers to.
"""
from __future__ import annotations

import threading
import warnings
from collections import OrderedDict
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _checks

CONSTANT_DATA_WARNING = (
    "A variable not marked as discrete takes only a single value, "
    "or values in a very small numerical range. "
    "If this is intentional, you can suppress this warning by passing preprocess=False. "
    "Note that this disables rescaling on all other variables as well.")

NOISE_SEED = 2_718281828      # _driver.py:874, fresh generator per task
NOISE_SCALE = 1e-10           # _driver.py:883
CONSTANT_STD = 1e-20          # _driver.py:879


@dataclass
class MiTask:
    """One estimation task; the field order mirrors the reference's 13-tuple (``_driver.py:479-483``)."""
    x: np.ndarray
    y: np.ndarray
    lag: int
    max_lag: int
    min_lag: int
    k: int
    mask: Optional[np.ndarray]
    cond: Optional[np.ndarray]
    cond_lag: np.ndarray
    discrete_x: bool
    discrete_y: bool
    preprocess: bool
    drop_nan: bool


def lagged_windows(t: MiTask) -> Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]:
    """``y(t) ~ x(t - lag) | z(t - cond_lag)``: y keeps the window ``[max_lag, len + min_lag)``,
    x and every cond column are shifted by their own lag (``_driver.py:794-806``)."""
    lo_pad = max(t.max_lag, 0)          # rows lost at the start
    hi_pad = min(t.min_lag, 0)          # rows lost at the end (<= 0)
    n = len(t.x)
    xs = t.x[lo_pad - t.lag: n - t.lag + hi_pad]
    ys = t.y[lo_pad: len(t.y) + hi_pad]
    zs = None
    if t.cond is not None:
        nz = len(t.cond)
        zs = np.column_stack([t.cond[lo_pad - t.cond_lag[j]: nz - t.cond_lag[j] + hi_pad, j]
                              for j in range(len(t.cond_lag))])
        
    return xs, ys, zs


def masked(xs, ys, zs, t: MiTask):
    """Mask is aligned with y; NaN rows are dropped across x, y and every cond column (``:834-854``)."""
    if t.mask is not None:
        keep = t.mask[max(t.max_lag, 0): len(t.mask) + min(t.min_lag, 0)]
        xs, ys = xs[keep], ys[keep]
        if zs is not None:
            zs = zs[keep]
    if t.drop_nan:
        ok = ~(np.isnan(xs) | np.isnan(ys))
        if zs is not None:
            ok &= ~np.max(np.isnan(zs), axis=1)
            zs = zs[ok]
        xs, ys = xs[ok], ys[ok]
    return xs, ys, zs


def validate(xs, ys, zs, t: MiTask) -> None:
    """Enough rows left and no NaNs in anything that is treated as numeric (``:856-869``)."""
    if len(ys) <= t.k:
        raise ValueError(_checks.MSG_K_TOO_LARGE)
    if (not t.discrete_x or xs.dtype.kind in "iufc") and np.isnan(xs).any():
        raise ValueError(_checks.MSG_NANS_LEFT)
    if (not t.discrete_y or ys.dtype.kind in "iufc") and np.isnan(ys).any():
        raise ValueError(_checks.MSG_NANS_LEFT)
    if not (t.discrete_x and t.discrete_y) and zs is not None and np.isnan(zs).any():
        raise ValueError(_checks.MSG_NANS_LEFT)


class _NoiseStream:
    """The reference draws its noise from ``default_rng(2718281828)`` created afresh for every
    task, so the values depend only on the sequence of draw shapes.  Tasks of one call repeat the
    same sequence thousands of times; the draws are memoised per shape sequence (bit-identical,
    small LRU) instead of regenerating ~N Gaussians per variable per task."""

    _cache: "OrderedDict[tuple, np.ndarray]" = OrderedDict()
    _lock = threading.Lock()
    MAX_BYTES = 1 << 28

    def __init__(self):
        self._shapes = ()
        self._rng = None

    def normal(self, shape) -> np.ndarray:
        shape = tuple(shape)
        key = self._shapes + (shape,)
        with self._lock:
            hit = self._cache.get(key)
            if hit is not None:
                self._cache.move_to_end(key)
        if hit is None:
            if self._rng is None:                     # replay the earlier draws to reach the stream position
                self._rng = np.random.default_rng(NOISE_SEED)
                for sh in self._shapes:
                    self._rng.normal(0.0, NOISE_SCALE, sh)
            hit = self._rng.normal(0.0, NOISE_SCALE, shape)
            hit.setflags(write=False)
            with self._lock:
                self._cache[key] = hit
                total = sum(a.nbytes for a in self._cache.values())
                while total > self.MAX_BYTES and len(self._cache) > 1:
                    _, old = self._cache.popitem(last=False)
                    total -= old.nbytes
        elif self._rng is not None:
            self._rng.normal(0.0, NOISE_SCALE, shape)  # keep a live generator in step
        self._shapes = key
        return hit


    def skip(self, shape) -> None:
        """Advance past a draw whose values the caller already has."""
        shape = tuple(shape)
        if self._rng is not None:
            self._rng.normal(0.0, NOISE_SCALE, shape)
        self._shapes = self._shapes + (shape,)


def rescaled(xs, ys, zs, discrete_x: bool, discrete_y: bool):
    """Unit variance plus N(0, 1e-10) noise from a fixed-seed generator; draw order x, y, z;
    discrete variables consume no draws; (near-)constant data is left alone with a warning
    (``_driver.py:871-902``)."""
    rng = _NoiseStream()

    def one(v):
        spread = v.std()
        if np.abs(spread) < CONSTANT_STD:
            warnings.warn(CONSTANT_DATA_WARNING)
            return v
        v = (v - v.mean()) / spread
        v += rng.normal(v.shape)
        return v

    if not discrete_x:
        xs = one(xs)
    if not discrete_y:
        ys = one(ys)
    if zs is not None and not (discrete_x and discrete_y):
        spread = zs.std(axis=0)
        if np.any(np.abs(spread) < CONSTANT_STD):
            warnings.warn(CONSTANT_DATA_WARNING)
        else:
            zs = (zs - zs.mean(axis=0)) / spread
            zs += rng.normal(zs.shape)
    return xs, ys, zs


def prepare(t: MiTask):
    xs, ys, zs = lagged_windows(t)
    xs, ys, zs = masked(xs, ys, zs, t)
    validate(xs, ys, zs, t)
    if t.preprocess:
        xs, ys, zs = rescaled(xs, ys, zs, t.discrete_x, t.discrete_y)
    return xs, ys, zs
