"""Argument validation shared by the public functions.

The messages are part of the drop-in contract: the reference's tests pin them verbatim
(``/root/reference/tests/unit/test_driver.py:26-37``; raised at ``ennemi/_driver.py:501-537``).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

MSG_K_TOO_LARGE = "k must be smaller than number of observations (after lag and mask)"
MSG_NANS_LEFT = "input contains NaNs (after applying the mask), pass drop_nan=True to ignore"
MSG_LAG_TOO_LARGE = "lag is too large, no observations left"


def k_is_valid(k) -> None:
    if not isinstance(k, int):
        raise TypeError("k must be int")
    if k <= 0:
        raise ValueError("k must be greater than zero")


def mask_is_valid(mask: np.ndarray, n_obs: int) -> None:
    if mask.ndim > 1:
        raise ValueError("mask must be one-dimensional")
    if len(mask) != n_obs:
        raise ValueError("mask length does not match input length")
    if mask.dtype != bool:
        raise TypeError("mask must contain only booleans")


def cond_is_valid(cond: np.ndarray, n_obs: int) -> None:
    if cond.ndim < 1 or cond.ndim > 2:
        raise ValueError("cond must be one- or two-dimensional")
    if len(cond) != n_obs:
        raise ValueError("x and cond must have same length")


def x_is_valid(x: np.ndarray) -> None:
    if x.ndim < 1 or x.ndim > 2:
        raise ValueError("x must be one- or two-dimensional")


def mi_arguments(x: np.ndarray, y: Optional[np.ndarray], k, cond: Optional[np.ndarray],
                 mask: Optional[np.ndarray]) -> None:
    """Checks in the reference's order (``_driver.py:501-517``): k, x shape, y shape, lengths, mask, cond."""
    k_is_valid(k)
    x_is_valid(x)
    if y is not None:
        if y.ndim > 1:
            raise ValueError("y must be one-dimensional")
        if x.shape[0] != y.shape[0]:
            raise ValueError("x and y must have same length")
    if mask is not None:
        mask_is_valid(mask, x.shape[0])
    if cond is not None:
        cond_is_valid(cond, x.shape[0])
