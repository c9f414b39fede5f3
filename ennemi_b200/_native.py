"""ctypes binding of ``libennemi_b200.so`` (C ABI in ``include/ennemi_b200.h``).

There is deliberately no CPU fallback: if the shared library is missing or no CUDA device is
usable, every estimator raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libennemi_b200.so")

FLAG_DEVICE_INPUT = 1
FLAG_BRUTE_COUNT = 2
FLAG_NO_PRUNE = 4
FLAG_SINGLE_USE = 8
FLAG_DEVICE_STATS = 16

ERR_CUDA, ERR_ARG, ERR_NONFINITE, ERR_UNSUPPORTED, ERR_CONSTANT = 1, 2, 3, 4, 5
P_LEN = 16
P_SUM, P_ZERO_A, P_ZERO_B, P_ZERO_C, P_ROWS, P_PAIRS = 0, 1, 2, 3, 4, 5

# every symbol include/ennemi_b200.h declares (tests check the library exports all of them)
EXPORTS = (
    "eb2_init", "eb2_shutdown", "eb2_device_count", "eb2_last_error", "eb2_version",
    "eb2_ksg_mi", "eb2_ksg_mi_rows", "eb2_ksg_mi_finish",
    "eb2_cmi", "eb2_cmi_rows", "eb2_cmi_finish",
    "eb2_ross_mi", "eb2_ross_cmi",
    "eb2_entropy", "eb2_entropy_rows", "eb2_entropy_finish", "eb2_entropy_cols",
    "eb2_psi", "eb2_kth_distance", "eb2_ball_count", "eb2_last_timing", "eb2_measure_fp64_peak",
    "eb2_cache_put", "eb2_cache_drop", "eb2_ksg_mi_cols", "eb2_cmi_cols", "eb2_last_data_flags",
    "eb2_mi_cols_batch", "eb2_ksg_mi_cols_rows", "eb2_cmi_cols_rows", "eb2_cache_stats",
    "eb2_cache_put_block", "eb2_cache_stats_many", "eb2_ksg_mi_pairs", "eb2_last_pipeline", "eb2_cache_put_block_dev", "eb2_sharded_ksg_mi",
)

_lib = None
_lib_lock = threading.Lock()


class ColDesc(ctypes.Structure):
    """``eb2_col_t``: one coordinate of the joint space taken from the device column cache."""
    _fields_ = [("key", ctypes.c_uint64), ("off", ctypes.c_int64), ("stride", ctypes.c_int64),
                ("mean", ctypes.c_double), ("std", ctypes.c_double),
                ("nkey", ctypes.c_uint64), ("noff", ctypes.c_int64), ("nstride", ctypes.c_int64)]


_c_dp = ctypes.POINTER(ctypes.c_double)
_c_lp = ctypes.POINTER(ctypes.c_int64)
_c_ip = ctypes.POINTER(ctypes.c_int32)
_i64, _int, _u32 = ctypes.c_int64, ctypes.c_int, ctypes.c_uint32
_vp = ctypes.c_void_p


def load():
    """Loads the shared library (once) and declares the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"ennemi_b200: CUDA library not built ({LIB_PATH} is missing). "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C ennemi_b200/csrc`. "
                "There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        lib.eb2_last_error.restype = ctypes.c_char_p
        lib.eb2_version.restype = ctypes.c_char_p
        # pointers are passed as void* so that host (numpy) and device (integer address) buffers share one path
        lib.eb2_ksg_mi.argtypes = [_int, _vp, _i64, _int, _u32, _c_dp, _vp, _vp, _vp]
        lib.eb2_ksg_mi_rows.argtypes = [_int, _vp, _i64, _int, _u32, _i64, _i64, _c_dp, _vp, _vp, _vp]
        lib.eb2_ksg_mi_finish.argtypes = [_c_dp, _i64, _int, _c_dp]
        lib.eb2_cmi.argtypes = [_int, _vp, _i64, _int, _int, _u32, _c_dp, _vp, _vp, _vp, _vp]
        lib.eb2_cmi_rows.argtypes = [_int, _vp, _i64, _int, _int, _u32, _i64, _i64, _c_dp, _vp, _vp, _vp, _vp]
        lib.eb2_cmi_finish.argtypes = [_c_dp, _i64, _int, _c_dp]
        lib.eb2_ross_mi.argtypes = [_int, _vp, _vp, _i64, _int, _int, _u32, _c_dp, _vp, _vp]
        lib.eb2_ross_cmi.argtypes = [_int, _vp, _vp, _i64, _int, _int, _int, _u32, _c_dp, _vp, _vp, _vp, _vp]
        lib.eb2_entropy.argtypes = [_int, _vp, _i64, _int, _int, _u32, _c_dp, _vp]
        lib.eb2_entropy_rows.argtypes = [_int, _vp, _i64, _int, _int, _u32, _i64, _i64, _c_dp, _vp]
        lib.eb2_entropy_finish.argtypes = [_c_dp, _i64, _int, _int, _c_dp]
        lib.eb2_psi.argtypes = [_int, _vp, _i64, _vp]
        lib.eb2_kth_distance.argtypes = [_int, _vp, _vp, _i64, _int, _int, _int, _u32, _vp]
        lib.eb2_ball_count.argtypes = [_int, _vp, _vp, _i64, _int, _int, _int, _vp, _u32, _vp]
        lib.eb2_last_timing.argtypes = [_int, _c_dp, ctypes.POINTER(_int)]
        lib.eb2_measure_fp64_peak.argtypes = [_int, _c_dp]
        lib.eb2_cache_put.argtypes = [_int, ctypes.c_uint64, _vp, _i64]
        lib.eb2_cache_drop.argtypes = [_int, ctypes.c_uint64]
        lib.eb2_cache_put_block.argtypes = [_int, _vp, _int, _vp, _i64, _i64]
        lib.eb2_cache_put_block_dev.argtypes = [_int, _vp, _int, _vp, _i64, _i64]
        lib.eb2_sharded_ksg_mi.argtypes = [_int, _vp, _i64, _int, _u32, _c_dp]
        lib.eb2_cache_stats_many.argtypes = [_int, _vp, _vp, _int, _i64, _i64, _vp, _vp]
        lib.eb2_ksg_mi_cols.argtypes = [_int, ctypes.POINTER(ColDesc), _i64, _int, _u32, _c_dp]
        lib.eb2_cmi_cols.argtypes = [_int, ctypes.POINTER(ColDesc), _i64, _int, _int, _u32, _c_dp]
        lib.eb2_entropy_cols.argtypes = [_int, ctypes.POINTER(ColDesc), _i64, _int, _int, _u32, _c_dp]
        lib.eb2_ksg_mi_cols_rows.argtypes = [_int, ctypes.POINTER(ColDesc), _i64, _int, _u32, _i64, _i64, _c_dp]
        lib.eb2_cmi_cols_rows.argtypes = [_int, ctypes.POINTER(ColDesc), _i64, _int, _int, _u32, _i64, _i64, _c_dp]
        lib.eb2_cache_stats.argtypes = [_int, ctypes.c_uint64, _i64, _i64, _i64, _c_dp, _c_dp]
        lib.eb2_mi_cols_batch.argtypes = [_int, ctypes.POINTER(ColDesc), _i64, _int, _i64, _int, _u32, _c_dp, ctypes.POINTER(_int)]
        lib.eb2_ksg_mi_pairs.argtypes = [_int, ctypes.POINTER(ColDesc), _int, _vp, _i64, _i64, _int, _u32, _c_dp, ctypes.POINTER(_int)]
        for name in EXPORTS:
            getattr(lib, name)
        _lib = lib
    return _lib


def device_count() -> int:
    return int(load().eb2_device_count())


def require_device() -> int:
    n = device_count()
    if n <= 0:
        raise RuntimeError("ennemi_b200: no usable CUDA device (the estimators run on B200 GPUs only; "
                           "there is no CPU fallback)")
    return n


def _raise(rc: int):
    msg = load().eb2_last_error().decode("utf-8", "replace")
    if rc == ERR_NONFINITE:
        raise ValueError(msg)              # what cKDTree raises in the reference
    if rc == ERR_ARG:
        raise ValueError(f"ennemi_b200: {msg}")
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(f"ennemi_b200: {msg}")
    raise RuntimeError(f"ennemi_b200: {msg}")


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def pack_coords(columns: Sequence[np.ndarray]) -> np.ndarray:
    """Dimension-major (d, n) contiguous fp64 block from 1-D / (n, c) column groups."""
    rows = []
    for col in columns:
        col = np.asarray(col)
        if col.ndim == 1:
            rows.append(col)
        else:
            rows.extend(col[:, j] for j in range(col.shape[1]))
    out = np.empty((len(rows), len(rows[0])), dtype=np.float64)
    for t, r in enumerate(rows):
        out[t, :] = r
    return out


def _outs(n: int, want: bool, count: int):
    if not want:
        return None, [None] * count
    return np.empty(n), [np.empty(n, dtype=np.int64) for _ in range(count)]


def ksg_mi(coords: np.ndarray, k: int, dev: int = 0, flags: int = 0, details: bool = False):
    lib = load()
    n = coords.shape[1]
    value = ctypes.c_double()
    eps, (nx, ny) = _outs(n, details, 2)
    rc = lib.eb2_ksg_mi(dev, coords.ctypes.data, n, k, flags, ctypes.byref(value), _ptr(eps), _ptr(nx), _ptr(ny))
    if rc:
        _raise(rc)
    return (value.value, {"eps": eps, "nx": nx, "ny": ny}) if details else value.value


def sharded_ksg_mi(coords: np.ndarray, k: int, ngpu: int, flags: int = 0) -> float:
    """One estimate with its query rows sharded over the first ``ngpu`` GPUs of the box, from this one process
    (``eb2_sharded_ksg_mi``); bit-identical for every ``ngpu``."""
    value = ctypes.c_double()
    rc = load().eb2_sharded_ksg_mi(ngpu, coords.ctypes.data, coords.shape[1], k, flags, ctypes.byref(value))
    if rc:
        _raise(rc)
    return value.value


def ksg_mi_rows(coords_ptr: int, n: int, k: int, row_lo: int, row_hi: int, dev: int = 0, flags: int = 0) -> np.ndarray:
    """Raw partial block for rows [row_lo, row_hi); ``coords_ptr`` is a host or device address."""
    lib = load()
    partial = np.zeros(P_LEN)
    rc = lib.eb2_ksg_mi_rows(dev, coords_ptr, n, k, flags, row_lo, row_hi,
                             partial.ctypes.data_as(_c_dp), None, None, None)
    if rc:
        _raise(rc)
    return partial


def ksg_mi_finish(partial: np.ndarray, n: int, k: int) -> float:
    lib = load()
    partial = np.ascontiguousarray(partial, dtype=np.float64)
    value = ctypes.c_double()
    rc = lib.eb2_ksg_mi_finish(partial.ctypes.data_as(_c_dp), n, k, ctypes.byref(value))
    if rc:
        _raise(rc)
    return value.value


def cmi(coords: np.ndarray, k: int, dev: int = 0, flags: int = 0, details: bool = False):
    lib = load()
    d, n = coords.shape
    value = ctypes.c_double()
    eps, (nxz, nyz, nz) = _outs(n, details, 3)
    rc = lib.eb2_cmi(dev, coords.ctypes.data, n, d - 2, k, flags, ctypes.byref(value),
                     _ptr(eps), _ptr(nxz), _ptr(nyz), _ptr(nz))
    if rc:
        _raise(rc)
    return (value.value, {"eps": eps, "nxz": nxz, "nyz": nyz, "nz": nz}) if details else value.value


def cmi_rows(coords_ptr: int, n: int, c: int, k: int, row_lo: int, row_hi: int, dev: int = 0, flags: int = 0) -> np.ndarray:
    lib = load()
    partial = np.zeros(P_LEN)
    rc = lib.eb2_cmi_rows(dev, coords_ptr, n, c, k, flags, row_lo, row_hi,
                          partial.ctypes.data_as(_c_dp), None, None, None, None)
    if rc:
        _raise(rc)
    return partial


def cmi_finish(partial: np.ndarray, n: int, k: int) -> float:
    lib = load()
    partial = np.ascontiguousarray(partial, dtype=np.float64)
    value = ctypes.c_double()
    rc = lib.eb2_cmi_finish(partial.ctypes.data_as(_c_dp), n, k, ctypes.byref(value))
    if rc:
        _raise(rc)
    return value.value


def ross_mi(coords: np.ndarray, cls: np.ndarray, ncls: int, k: int, dev: int = 0, flags: int = 0, details: bool = False):
    lib = load()
    n = coords.shape[1]
    value = ctypes.c_double()
    eps, (nfull,) = _outs(n, details, 1)
    rc = lib.eb2_ross_mi(dev, coords.ctypes.data, cls.ctypes.data, n, ncls, k, flags, ctypes.byref(value),
                         _ptr(eps), _ptr(nfull))
    if rc:
        _raise(rc)
    return (value.value, {"eps": eps, "n_full": nfull}) if details else value.value


def ross_cmi(coords: np.ndarray, cls: np.ndarray, ncls: int, k: int, dev: int = 0, flags: int = 0, details: bool = False):
    lib = load()
    d, n = coords.shape
    value = ctypes.c_double()
    eps, (nxz, nyz, nz) = _outs(n, details, 3)
    rc = lib.eb2_ross_cmi(dev, coords.ctypes.data, cls.ctypes.data, n, d - 1, ncls, k, flags, ctypes.byref(value),
                          _ptr(eps), _ptr(nxz), _ptr(nyz), _ptr(nz))
    if rc:
        _raise(rc)
    return (value.value, {"eps": eps, "nxz": nxz, "nyz": nyz, "nz": nz}) if details else value.value


def entropy(coords: np.ndarray, k: int, dev: int = 0, flags: int = 0, details: bool = False):
    lib = load()
    m, n = coords.shape
    value = ctypes.c_double()
    dist = np.empty(n) if details else None
    rc = lib.eb2_entropy(dev, coords.ctypes.data, n, m, k, flags, ctypes.byref(value), _ptr(dist))
    if rc:
        _raise(rc)
    return (value.value, {"dist": dist}) if details else value.value


def entropy_rows(coords_ptr: int, n: int, m: int, k: int, row_lo: int, row_hi: int, dev: int = 0, flags: int = 0) -> np.ndarray:
    lib = load()
    partial = np.zeros(P_LEN)
    rc = lib.eb2_entropy_rows(dev, coords_ptr, n, m, k, flags, row_lo, row_hi, partial.ctypes.data_as(_c_dp), None)
    if rc:
        _raise(rc)
    return partial


def entropy_finish(partial: np.ndarray, n: int, m: int, k: int) -> float:
    lib = load()
    partial = np.ascontiguousarray(partial, dtype=np.float64)
    value = ctypes.c_double()
    rc = lib.eb2_entropy_finish(partial.ctypes.data_as(_c_dp), n, m, k, ctypes.byref(value))
    if rc:
        _raise(rc)
    return value.value


def psi(counts: np.ndarray, dev: int = 0) -> np.ndarray:
    lib = load()
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    out = np.empty(counts.shape)
    rc = lib.eb2_psi(dev, counts.ctypes.data, counts.size, out.ctypes.data)
    if rc:
        _raise(rc)
    return out


def kth_distance(coords: np.ndarray, k: int, cls: Optional[np.ndarray] = None, ncls: int = 0, dev: int = 0,
                 flags: int = 0) -> np.ndarray:
    lib = load()
    d, n = coords.shape
    out = np.empty(n)
    rc = lib.eb2_kth_distance(dev, coords.ctypes.data, _ptr(cls), n, d, ncls, k, flags, out.ctypes.data)
    if rc:
        _raise(rc)
    return out


def ball_count(coords: np.ndarray, radius: np.ndarray, cls: Optional[np.ndarray] = None, ncls: int = 0,
               within_class: bool = False, dev: int = 0, flags: int = 0) -> np.ndarray:
    lib = load()
    d, n = coords.shape
    radius = np.ascontiguousarray(radius, dtype=np.float64)
    out = np.empty(n, dtype=np.int64)
    rc = lib.eb2_ball_count(dev, coords.ctypes.data, _ptr(cls), n, d, ncls, int(within_class), radius.ctypes.data,
                            flags, out.ctypes.data)
    if rc:
        _raise(rc)
    return out


def last_timing(dev: int = 0) -> dict:
    lib = load()
    ms = (ctypes.c_double * 5)()
    launches = _int()
    rc = lib.eb2_last_timing(dev, ms, ctypes.byref(launches))
    if rc:
        _raise(rc)
    return {"total_ms": ms[0], "knn_ms": ms[1], "count_ms": ms[2], "psi_ms": ms[3], "layout_ms": ms[4],
            "launches": launches.value}


def last_pipeline(dev: int = 0) -> int:
    """1 if the last KSG call on ``dev`` ran on the bivariate pipeline, 0 if the general path took it."""
    return int(load().eb2_last_pipeline(dev))


def measure_fp64_peak(dev: int = 0) -> float:
    """FP64 (DADD) issue rate of the device in 1e12 instructions/s — the roofline denominator."""
    lib = load()
    out = ctypes.c_double()
    rc = lib.eb2_measure_fp64_peak(dev, ctypes.byref(out))
    if rc:
        _raise(rc)
    return out.value


class NonFiniteInput(ValueError):
    """Raised by the ``*_cols`` calls; ``nan`` tells NaN input from otherwise non-finite data."""

    def __init__(self, msg, nan):
        super().__init__(msg)
        self.nan = nan


class ConstantWindow(Exception):
    """``FLAG_DEVICE_STATS``: a window was constant; the task has to be repeated with host-checked statistics."""


def cache_put(key: int, column: np.ndarray, dev: int = 0) -> None:
    lib = load()
    column = np.ascontiguousarray(column, dtype=np.float64)
    rc = lib.eb2_cache_put(dev, key, column.ctypes.data, column.size)
    if rc:
        _raise(rc)


def block_layout(block: np.ndarray):
    """Row stride (in elements) of a 2-D float64 array whose rows are runs of consecutive doubles — the
    layout ``eb2_cache_put_block`` uploads with one copy — or None for any other layout."""
    if block.ndim != 2 or block.dtype != np.float64 or block.shape[0] == 0 or block.shape[1] < 2:
        return None
    s0, s1 = block.strides
    if s1 != 8 or s0 % 8 != 0 or s0 < 8 * block.shape[1]:
        return None
    return s0 // 8


def cache_put_block(keys: Sequence[int], block: np.ndarray, dev: int = 0) -> None:
    """Caches every column of a row-major ``(n, ncols)`` float64 block (see :func:`block_layout`) with one
    host-to-device copy; the de-interleave runs on the device."""
    lib = load()
    ld = block_layout(block)
    if ld is None or len(keys) != block.shape[1]:
        raise ValueError("cache_put_block: need a 2-D float64 array with unit column stride and one key per column")
    karr = np.asarray(keys, dtype=np.uint64)
    rc = lib.eb2_cache_put_block(dev, karr.ctypes.data, len(keys), block.ctypes.data, block.shape[0], ld)
    if rc:
        _raise(rc)


def cache_put_block_dev(keys: Sequence[int], ptr: int, n: int, ld: int, dev: int = 0) -> None:
    """:func:`cache_put_block` from a row-major block that already is in device memory at ``ptr`` on ``dev``."""
    karr = np.asarray(keys, dtype=np.uint64)
    rc = load().eb2_cache_put_block_dev(dev, karr.ctypes.data, len(keys), ptr, n, ld)
    if rc:
        _raise(rc)


def cache_stats_many(keys: Sequence[int], offs: Sequence[int], n: int, stride: int = 1, dev: int = 0):
    """(means, stds) arrays of several same-length cached column windows: one call, NumPy's bits."""
    lib = load()
    karr = np.asarray(keys, dtype=np.uint64)
    oarr = np.asarray(offs, dtype=np.int64)
    means, stds = np.empty(len(karr)), np.empty(len(karr))
    rc = lib.eb2_cache_stats_many(dev, karr.ctypes.data, oarr.ctypes.data, len(karr), stride, n,
                                  means.ctypes.data, stds.ctypes.data)
    if rc:
        _raise(rc)
    return means, stds


def cache_drop(key: int, dev: int = 0) -> None:
    lib = load()
    rc = lib.eb2_cache_drop(dev, key)
    if rc:
        _raise(rc)


def _cols_call(fn, cols, *args):
    arr = (ColDesc * len(cols))(*cols)
    value = ctypes.c_double()
    rc = fn(*args[:1], arr, *args[1:], ctypes.byref(value))
    if rc == ERR_NONFINITE:
        lib = load()
        raise NonFiniteInput(lib.eb2_last_error().decode(), bool(lib.eb2_last_data_flags() & 1))
    if rc == ERR_CONSTANT:
        raise ConstantWindow()
    if rc:
        _raise(rc)
    return value.value


def ksg_mi_cols(cols, n: int, k: int, dev: int = 0, flags: int = 0) -> float:
    """KSG MI of two cached device columns (``cols``: two :class:`ColDesc`)."""
    return _cols_call(load().eb2_ksg_mi_cols, cols, dev, n, k, flags)


def cmi_cols(cols, n: int, k: int, dev: int = 0, flags: int = 0) -> float:
    """Frenzel-Pompe CMI of cached device columns (``cols``: x, y, then the condition's columns)."""
    return _cols_call(load().eb2_cmi_cols, cols, dev, n, len(cols) - 2, k, flags)


def entropy_cols(cols, n: int, k: int, dev: int = 0, flags: int = 0) -> float:
    """k-NN entropy of the space spanned by cached device columns (``std = 0`` descriptors: values as they are)."""
    return _cols_call(load().eb2_entropy_cols, cols, dev, n, len(cols), k, flags)


def mi_cols_batch(tasks, n: int, k: int, dev: int = 0, flags: int = 0):
    """``tasks``: list of descriptor lists of equal length (2 + c).  Returns (values, status) arrays;
    ``status[t]`` is 0 or ``err | data_flags << 8``."""
    lib = load()
    d = len(tasks[0])
    flat = (ColDesc * (len(tasks) * d))(*[c for t in tasks for c in t])
    values = np.empty(len(tasks))
    status = (ctypes.c_int * len(tasks))()
    rc = lib.eb2_mi_cols_batch(dev, flat, len(tasks), d - 2, n, k, flags, values.ctypes.data_as(_c_dp), status)
    if rc:
        _raise(rc)
    return values, np.frombuffer(status, dtype=np.int32).copy()


def ksg_mi_pairs(cols, pairs: np.ndarray, n: int, k: int, dev: int = 0, flags: int = 0):
    """Every pair ``(cols[pairs[t, 0]], cols[pairs[t, 1]])`` of the prepared variables ``cols`` (descriptors) in one
    call (``eb2_ksg_mi_pairs``).  Returns (values, status) arrays as :func:`mi_cols_batch` does."""
    lib = load()
    arr = (ColDesc * len(cols))(*cols)
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    values = np.empty(len(pairs))
    status = (ctypes.c_int * max(len(pairs), 1))()
    rc = lib.eb2_ksg_mi_pairs(dev, arr, len(cols), pairs.ctypes.data, len(pairs), n, k, flags,
                              values.ctypes.data_as(_c_dp), status)
    if rc:
        _raise(rc)
    return values, np.frombuffer(status, dtype=np.int32)[:len(pairs)].copy()


def mi_cols_rows(cols, n: int, k: int, row_lo: int, row_hi: int, dev: int = 0, flags: int = 0) -> np.ndarray:
    """Partial block of a KSG (2 columns) or CMI (2 + c columns) estimate on cached columns for the
    query rows [row_lo, row_hi)."""
    lib = load()
    arr = (ColDesc * len(cols))(*cols)
    partial = np.zeros(P_LEN)
    if len(cols) == 2:
        rc = lib.eb2_ksg_mi_cols_rows(dev, arr, n, k, flags, row_lo, row_hi, partial.ctypes.data_as(_c_dp))
    else:
        rc = lib.eb2_cmi_cols_rows(dev, arr, n, len(cols) - 2, k, flags, row_lo, row_hi, partial.ctypes.data_as(_c_dp))
    if rc == ERR_NONFINITE:
        raise NonFiniteInput(lib.eb2_last_error().decode(), bool(lib.eb2_last_data_flags() & 1))
    if rc == ERR_CONSTANT:
        raise ConstantWindow()
    if rc:
        _raise(rc)
    return partial


def cache_stats(key: int, off: int, n: int, stride: int = 1, dev: int = 0):
    """(mean, std) of a cached column window with NumPy's summation order (bit-identical to
    ``view.mean()``, ``view.std()``), computed on the device."""
    lib = load()
    mean, std = ctypes.c_double(), ctypes.c_double()
    rc = lib.eb2_cache_stats(dev, key, off, stride, n, ctypes.byref(mean), ctypes.byref(std))
    if rc:
        _raise(rc)
    return mean.value, std.value
