"""Public API — a drop-in for ``ennemi``'s (``ennemi/__init__.py:6-10``): same function names,
keyword arguments, output shapes, pandas behaviour, warnings and error messages
(``ennemi/_driver.py``; contract summarised in SURVEY.md Appendix B).

What differs is underneath: each (variable, lag) or (variable pair) task is prepared on the host
exactly as the reference does and then estimated by the CUDA library; tasks fan out over the
GPUs of the box instead of a CPU thread pool.
"""
from __future__ import annotations

import itertools
import os
import sys
import warnings
from typing import Callable, Optional

import numpy as np

from . import _align, _checks, _columns, _devices, _estimators as est, _native, _schedule
from ._align import MiTask
from ._columns import ColsTask

DISCRETE_NORMALIZATION_WARNING = (
    "You have set normalize=True while at least one variable is discrete. "
    "The correlation coefficient formula assumes both variables to be continuous, "
    "and the results are not comparable across different discrete variables. "
    "Compare the raw MI against entropy of the discrete variable instead. "
    "If you really want to calculate correlation coefficients, you can suppress "
    "this warning by setting normalize=False and calling normalize_mi() on the results.")

PREPROCESS_CONSTANT_DATA_WARNING = _align.CONSTANT_DATA_WARNING


def _pandas():
    """pandas, but only if the caller has already imported it (as the reference does)."""
    return sys.modules.get("pandas")


# ---------------------------------------------------------------------------------------------
# normalisation
# ---------------------------------------------------------------------------------------------
def _to_corr(mi: float) -> float:
    # correlation coefficient of the bivariate Gaussian with this MI; non-positive values pass through
    return mi if mi <= 0.0 else np.sqrt(1 - np.exp(-2 * mi))


def normalize_mi(mi):
    """Map MI (nats) to the correlation-coefficient scale ``sqrt(1 - exp(-2 MI))``.

    Negative estimates are returned unchanged.  pandas inputs keep their index and columns.
    Same as passing ``normalize=True`` to the estimators.  (``_driver.py:43-81``)
    """
    pd = _pandas()
    if pd is not None and isinstance(mi, (pd.DataFrame, pd.Series)):
        return mi.map(_to_corr)
    return np.vectorize(_to_corr, otypes=[float])(mi)


# ---------------------------------------------------------------------------------------------
# entropy
# ---------------------------------------------------------------------------------------------
def estimate_entropy(x, *, k: int = 3, multidim: bool = False, discrete: bool = False,
                     mask=None, cond=None, drop_nan: bool = False):
    """Differential entropy (nats) of one or more continuous variables by the k-NN estimator.

    Columns of a 2-D ``x`` are separate variables unless ``multidim=True`` makes them one
    m-dimensional variable.  ``discrete=True`` switches to the plug-in entropy of the observed
    categories.  ``mask`` selects observations; ``cond`` gives H(X | cond) by the chain rule
    H(X, cond) - H(cond); ``drop_nan`` removes rows with NaNs.  pandas input gives a one-row
    ``DataFrame`` (except with ``multidim``).  (``_driver.py:84-226``)
    """
    x_arr = np.asarray(x)
    if mask is not None:
        mask = np.asarray(mask)
        _checks.mask_is_valid(mask, x_arr.shape[0])
    _checks.x_is_valid(x_arr)
    _checks.k_is_valid(k)

    if cond is None:
        result = _entropy_of(x_arr, k, multidim, mask, discrete, drop_nan)
    else:
        cond_arr = np.asarray(cond)
        _checks.cond_is_valid(cond_arr, x_arr.shape[0])
        # chain rule, no bias correction: H(X | C) = H(X, C) - H(C)
        if _cond_entropy_on_device(x_arr, cond_arr, mask, discrete, drop_nan):
            result = _cond_entropy_device(x_arr, cond_arr, k, multidim)
        else:
            h_cond = _entropy_of(cond_arr, k, True, mask, discrete, drop_nan)
            if multidim or x_arr.ndim == 1:
                joint = _entropy_rows(np.column_stack((x_arr, cond_arr)), k, mask, discrete, drop_nan)
                result = np.asarray(joint - h_cond)
            else:
                joint = np.empty(x_arr.shape[1])
                for j in range(x_arr.shape[1]):
                    joint[j] = _entropy_rows(np.column_stack((x_arr[:, j], cond_arr)), k, mask, discrete, drop_nan)
                result = joint - h_cond

    pd = _pandas()
    if not multidim and pd is not None:
        if isinstance(x, pd.DataFrame):
            return pd.DataFrame(np.atleast_2d(result), columns=x.columns)
        if isinstance(x, pd.Series):
            return pd.DataFrame(np.atleast_2d(result), columns=[x.name])
    return result


def _cond_entropy_on_device(x_arr, cond_arr, mask, discrete: bool, drop_nan: bool) -> bool:
    """Whether H(X | cond) runs on device-resident columns: continuous float64 data, every row used, large enough for
    the uploads to matter, one process per estimate (a row-sharded job keeps its own route)."""
    if discrete or mask is not None or x_arr.shape[0] < DEVICE_COLUMNS_MIN_ROWS:
        return False
    if x_arr.dtype != np.float64 or cond_arr.dtype != np.float64 or x_arr.ndim > 2 or cond_arr.ndim > 2:
        return False
    if drop_nan and (np.isnan(x_arr).any() or np.isnan(cond_arr).any()):
        return False                                  # (rows are dropped per variable: different row sets)
    from . import distributed
    return not distributed.row_sharding_enabled()


def _cond_entropy_device(x_arr: np.ndarray, cond_arr: np.ndarray, k: int, multidim: bool):
    """H(X_j | C) = H(X_j, C) - H(C) (``_driver.py:202-220``) with every column of X and C uploaded ONCE: the columns of C
    are named in both terms of every variable instead of being stacked and copied again per term.
    Same checks, in the reference's order (``_driver.py:180-200``)."""
    n = x_arr.shape[0]
    if k >= n:
        raise ValueError(_checks.MSG_K_TOO_LARGE)
    if np.isnan(cond_arr).any() or np.isnan(x_arr).any():
        raise ValueError(_checks.MSG_NANS_LEFT)
    dev = _devices.current()
    store = _columns.ColumnStore()
    try:
        xkeys = store.add_columns(x_arr)
        ckeys = store.add_columns(cond_arr)
        for key in xkeys + ckeys:
            store.ensure(dev, key)

        def desc(key):
            return _native.ColDesc(key, 0, 1, 0.0, 0.0, 0, 0, 1)          # std = 0: the values as they are

        cdesc = [desc(key) for key in ckeys]
        h_cond = _native.entropy_cols(cdesc, n, k, dev=dev)
        if multidim or x_arr.ndim == 1:
            return np.asarray(_native.entropy_cols([desc(key) for key in xkeys] + cdesc, n, k, dev=dev) - h_cond)
        return np.array([_native.entropy_cols([desc(key)] + cdesc, n, k, dev=dev) for key in xkeys]) - h_cond
    finally:
        store.close()


def _entropy_rows(rows: np.ndarray, k: int, mask, discrete: bool, drop_nan: bool) -> float:
    """Mask, drop NaNs, validate (``_driver.py:180-200``) and estimate one variable."""
    if mask is not None:
        rows = rows[np.asanyarray(mask)]
    if drop_nan and not discrete:
        bad = np.isnan(rows)
        rows = rows[~(np.max(bad, axis=1) if rows.ndim > 1 else bad)]
    if k >= rows.shape[0]:
        raise ValueError(_checks.MSG_K_TOO_LARGE)
    if not discrete and _block_entropy_on_device(rows):
        # a row-major (n, m) block goes up as it is (one copy, split into columns on the device) instead of being
        # transposed on the host first; NaNs are found by the device's own scan of the data
        try:
            return _entropy_block_device(rows, k)
        except _native.NonFiniteInput as e:
            if e.nan:
                raise ValueError(_checks.MSG_NANS_LEFT) from None
            raise
    if not discrete and np.any(np.isnan(rows)):
        raise ValueError(_checks.MSG_NANS_LEFT)
    if discrete:
        return est._estimate_discrete_entropy(rows)
    return est._estimate_single_entropy(rows, k)


def _block_entropy_on_device(rows: np.ndarray) -> bool:
    if rows.ndim != 2 or rows.shape[1] < 2 or rows.dtype != np.float64 or rows.shape[0] < DEVICE_COLUMNS_MIN_ROWS:
        return False
    if _native.block_layout(rows) is None:
        return False
    from . import distributed
    return not distributed.row_sharding_enabled()


def _entropy_block_device(rows: np.ndarray, k: int) -> float:
    dev = _devices.current()
    store = _columns.ColumnStore()
    try:
        keys = store.add_columns(rows)
        for key in keys:
            store.ensure(dev, key)
        return _native.entropy_cols([_native.ColDesc(key, 0, 1, 0.0, 0.0, 0, 0, 1) for key in keys], rows.shape[0], k, dev=dev)
    finally:
        store.close()


def _entropy_of(x: np.ndarray, k: int, multidim: bool, mask, discrete: bool, drop_nan: bool):
    if multidim or x.ndim == 1:
        return np.asarray(_entropy_rows(x, k, mask, discrete, drop_nan))
    if (not discrete and mask is None and x.shape[1] >= 2 and _block_entropy_on_device(x)
            and not (drop_nan and np.isnan(x).any())):
        return _entropy_columns_device(x, k)
    out = np.empty(x.shape[1])
    for j in range(x.shape[1]):
        out[j] = _entropy_rows(x[:, j], k, mask, discrete, drop_nan)
    return out


def _entropy_columns_device(x: np.ndarray, k: int) -> np.ndarray:
    """H(X_j) of every column of a row-major (n, nvar) array: ONE upload of the block (split into columns on the device),
    one estimate per column from the cached columns - instead of a strided host copy and an upload per variable.
    Checks as ``_entropy_rows`` makes them, column by column."""
    if k >= x.shape[0]:
        raise ValueError(_checks.MSG_K_TOO_LARGE)
    dev = _devices.current()
    store = _columns.ColumnStore()
    try:
        keys = store.add_columns(x)
        for key in keys:
            store.ensure(dev, key)
        out = np.empty(len(keys))
        for j, key in enumerate(keys):
            try:
                out[j] = _native.entropy_cols([_native.ColDesc(key, 0, 1, 0.0, 0.0, 0, 0, 1)], x.shape[0], k, dev=dev)
            except _native.NonFiniteInput as e:
                if e.nan:
                    raise ValueError(_checks.MSG_NANS_LEFT) from None
                raise
        return out
    finally:
        store.close()


# ---------------------------------------------------------------------------------------------
# mutual information
# ---------------------------------------------------------------------------------------------
# below this many rows the per-task host preparation is cheaper than caching columns on the device
DEVICE_COLUMNS_MIN_ROWS = int(os.environ.get("ENNEMI_B200_COLUMNS_MIN_ROWS", "20000"))


def _device_store(arrays, mask, drop_nan: bool, any_discrete: bool):
    """A :class:`_columns.ColumnStore` when the call qualifies for device-resident columns, else None."""
    if arrays[0].shape[0] < DEVICE_COLUMNS_MIN_ROWS:
        return None
    if not _columns.eligible(arrays, mask, drop_nan, any_discrete):
        return None
    return _columns.ColumnStore()


def _run_task(t) -> float:
    """Prepare one task on the host and dispatch to the estimator its variable types call for
    (``_driver.py:815-832``).  For a discrete variable the continuous one goes first."""
    if isinstance(t, ColsTask):          # continuous variables already resident on the GPU
        return t.run()
    xs, ys, zs = _align.prepare(t)
    if zs is None:
        if t.discrete_x and t.discrete_y:
            return est._estimate_discrete_mi(xs, ys)
        if t.discrete_x:
            return est._estimate_semidiscrete_mi(ys, xs, t.k)
        if t.discrete_y:
            return est._estimate_semidiscrete_mi(xs, ys, t.k)
        return est._estimate_single_mi(xs, ys, t.k)
    if t.discrete_x and t.discrete_y:
        return est._estimate_conditional_discrete_mi(xs, ys, zs)
    if t.discrete_x:
        return est._estimate_conditional_semidiscrete_mi(ys, xs, zs, t.k)
    if t.discrete_y:
        return est._estimate_conditional_semidiscrete_mi(xs, ys, zs, t.k)
    return est._estimate_conditional_mi(xs, ys, zs, t.k)


def estimate_mi(y, x, lag=0, *, k: int = 3, cond=None, cond_lag=0, mask=None,
                discrete_y: bool = False, discrete_x: bool = False, preprocess: bool = True,
                drop_nan: bool = False, normalize: bool = False, max_threads: Optional[int] = None,
                callback: Optional[Callable[[int, int], None]] = None):
    """Mutual information (nats) between ``y`` and every column of ``x`` at every ``lag``.

    Returns a ``(len(lag), n_variables)`` array — a ``DataFrame`` indexed by lag when ``x`` is
    pandas.  The model is ``y(t) ~ x(t - lag) | cond(t - cond_lag)``; ``y`` is cropped to
    ``y[max(max_lag, 0) : N + min(min_lag, 0)]`` so that it stays fixed across lags.

    ``cond`` switches to conditional MI (Frenzel-Pompe); ``discrete_x`` / ``discrete_y`` to the
    discrete-continuous estimator (Ross), or the plug-in estimator when both are set (``cond`` is
    then discrete too).  ``mask`` selects ``y`` observations (lags applied to ``x`` and ``cond``).
    ``preprocess`` rescales continuous variables to unit variance and adds fixed-seed 1e-10 noise;
    ``drop_nan`` removes rows with NaNs; ``normalize`` returns correlation coefficients
    (:func:`normalize_mi`).  ``max_threads`` bounds the number of concurrent tasks (GPU workers);
    ``callback(var_index, lag)`` is invoked once per task after it has finished.  (``_driver.py:229-361``)

    Callback timing differs from the reference's thread pool: tasks on device-resident columns are estimated in
    batches (all unconditional tasks of a GPU in ONE library call), so their callbacks arrive in bursts when a batch
    returns; with task fan-out over ranks (:mod:`ennemi_b200.distributed`) every rank sees the callbacks of its own
    share of the tasks only.
    """
    x_arr = np.asarray(x)
    y_arr = np.asarray(y)
    cond_arr = None
    n_cond = 1
    if cond is not None:
        cond_arr = np.column_stack((np.asarray(cond),))
        n_cond = cond_arr.shape[1]
    mask_arr = None if mask is None else np.asarray(mask)
    lags = np.atleast_1d(lag)
    cond_lags = np.broadcast_to(np.column_stack((cond_lag,)), (lags.shape[0], n_cond))

    _checks.mi_arguments(x_arr, y_arr, k, cond_arr, mask_arr)
    lo = min(np.min(lags), np.min(cond_lags))
    hi = max(np.max(lags), np.max(cond_lags))
    if hi - lo >= y_arr.size or hi >= y_arr.size or lo <= -y_arr.size:
        raise ValueError(_checks.MSG_LAG_TOO_LARGE)

    n_var = 1 if x_arr.ndim == 1 else x_arr.shape[1]
    cells = list(itertools.product(range(len(lags)), range(n_var)))
    x_cols = [x_arr if x_arr.ndim == 1 else x_arr[:, v] for v in range(n_var)]
    store = _device_store([x_arr, y_arr, cond_arr], mask_arr, drop_nan, discrete_x or discrete_y)
    if store is not None:
        # every variable goes to the GPU once; a task is a set of lag offsets into the cached columns
        # (a row-major (n, nvar) array travels as one block and is split into columns on the device)
        store.full_stats = bool(preprocess and hi == 0 and lo == 0 and n_var > 1)
        xkeys = store.add_columns(x_arr)
        ykey = store.add(y_arr)
        zkeys = [] if cond_arr is None else store.add_columns(cond_arr)
        tasks = [ColsTask(store, xkeys[v], ykey, zkeys, x_cols[v], y_arr, cond_arr, lags[li], hi, lo, cond_lags[li],
                          k, preprocess) for li, v in cells]
        if len(tasks) == 1:
            tasks[0].single_use = True
    else:
        tasks = [MiTask(x_cols[v], y_arr, lags[li], hi, lo, k, mask_arr, cond_arr,
                        cond_lags[li], discrete_x, discrete_y, preprocess, drop_nan) for li, v in cells]

    def done(i: int) -> None:
        if callback is not None:
            li, v = cells[i]
            callback(v, lags[li])

    per_task = _schedule.gpu_time_estimate(len(y_arr), 0 if cond_arr is None else n_cond, k)
    try:
        values = _schedule.run_tasks(_run_task, tasks, max_threads, per_task, done)
    finally:
        if store is not None:
            store.close()

    result = np.empty((len(lags), n_var))
    for cell, value in zip(cells, values):
        result[cell] = value

    if normalize:
        if discrete_x or discrete_y:
            warnings.warn(DISCRETE_NORMALIZATION_WARNING)
        result = normalize_mi(result)

    pd = _pandas()
    if pd is not None:
        if isinstance(x, pd.DataFrame):
            return pd.DataFrame(result, index=lags, columns=x.columns)
        if isinstance(x, pd.Series):
            return pd.DataFrame(result, index=lags, columns=[x.name])
    return result


def estimate_corr(y, x, lag=0, *, k: int = 3, cond=None, cond_lag=0, mask=None, preprocess: bool = True,
                  drop_nan: bool = False, max_threads: Optional[int] = None,
                  callback: Optional[Callable[[int, int], None]] = None):
    """:func:`estimate_mi` for continuous variables with ``normalize=True`` (``_driver.py:363-449``)."""
    return estimate_mi(y, x, lag, k=k, cond=cond, cond_lag=cond_lag, mask=mask, preprocess=preprocess,
                       drop_nan=drop_nan, normalize=True, max_threads=max_threads, callback=callback)


def pairwise_mi(data, *, k: int = 3, cond=None, mask=None, discrete=False, preprocess: bool = True,
                drop_nan: bool = False, normalize: bool = False, max_threads: Optional[int] = None,
                callback: Optional[Callable[[int, int], None]] = None):
    """Symmetric matrix of MI between every pair of columns of ``data``; NaN on the diagonal.

    ``discrete`` marks discrete columns (scalar or one flag per column); ``cond`` is allowed only
    when the data are all continuous or all discrete.  Other options as in :func:`estimate_mi`.
    ``callback(i, j)`` fires once per pair after it has finished (in bursts: the pairs of a GPU are estimated
    by one library call; with task fan-out over ranks each rank sees its own share only - see
    :func:`estimate_mi`).  A ``DataFrame`` in gives a ``DataFrame`` out.  (``_driver.py:540-624, 680-723``)
    """
    data_arr = np.asarray(data)
    cond_arr = None if cond is None else np.column_stack((np.asarray(cond),))
    mask_arr = None if mask is None else np.asarray(mask)
    if data_arr.ndim == 1 or data_arr.shape[1] == 1:
        return np.full((1, 1), np.nan)
    flags = np.broadcast_to(discrete, data_arr.shape[1])

    _checks.mi_arguments(data_arr, None, k, cond_arr, mask_arr)
    if cond_arr is not None and not (np.all(flags) or np.all(~flags)):
        raise ValueError("Conditioning is not supported with mixed discrete and continuous data. "
                         "This is a limitation that can be lifted in the future (see "
                         "https://github.com/polsys/ennemi/issues/87).")

    n_obs, n_var = data_arr.shape
    zero_lag = np.asarray(0) if cond_arr is None else np.full(cond_arr.shape[1], 0)
    pairs = [(i, j) for i in range(n_var) for j in range(i + 1, n_var)]
    store = _device_store([data_arr, cond_arr], mask_arr, drop_nan, bool(flags.any()))
    if store is not None:
        store.full_stats = bool(preprocess)
        store.sharded_upload = len(pairs) >= 64        # (every rank has tasks, so every rank reaches the upload)
        keys = store.add_columns(data_arr)
        zkeys = [] if cond_arr is None else store.add_columns(cond_arr)
        tasks = [ColsTask(store, keys[i], keys[j], zkeys, data_arr[:, i], data_arr[:, j], cond_arr, 0, 0, 0,
                          np.atleast_1d(zero_lag), k, preprocess) for i, j in pairs]
        for t in tasks:
            t.share_prepared = True
    else:
        tasks = [MiTask(data_arr[:, i], data_arr[:, j], 0, 0, 0, k, mask_arr, cond_arr, zero_lag,
                        flags[i], flags[j], preprocess, drop_nan) for i, j in pairs]

    def done(t: int) -> None:
        if callback is not None:
            callback(*pairs[t])

    per_task = _schedule.gpu_time_estimate(n_obs, 0 if cond_arr is None else cond_arr.shape[1], k)
    try:
        values = _schedule.run_tasks(_run_task, tasks, max_threads, per_task, done)
    finally:
        if store is not None:
            store.close()

    result = np.full((n_var, n_var), np.nan)
    for (i, j), value in zip(pairs, values):
        result[i, j] = value
        result[j, i] = value

    if normalize:
        if flags.any():
            warnings.warn(DISCRETE_NORMALIZATION_WARNING)
        result = normalize_mi(result)

    pd = _pandas()
    if pd is not None and isinstance(data, pd.DataFrame):
        return pd.DataFrame(result, index=data.columns, columns=data.columns)
    return result


def pairwise_corr(data, *, k: int = 3, cond=None, mask=None, preprocess: bool = True, drop_nan: bool = False,
                  max_threads: Optional[int] = None, callback: Optional[Callable[[int, int], None]] = None):
    """:func:`pairwise_mi` with ``normalize=True`` (``_driver.py:627-677``)."""
    return pairwise_mi(data, k=k, cond=cond, mask=mask, preprocess=preprocess, drop_nan=drop_nan,
                       normalize=True, max_threads=max_threads, callback=callback)
