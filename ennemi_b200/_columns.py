"""Device-resident columns for lag sweeps and ``pairwise_mi``.

The reference prepares every (variable, lag) task from scratch on the host: slice, rescale, add
noise, build trees.  Here every variable of a call is uploaded to the GPU ONCE; a task is then just
a list of column descriptors — which cached column, the lag offset, the window length, the mean and
standard deviation of that window (computed by NumPy on the host exactly as the reference does, so
the bits agree) and which cached noise vector to add.  The rescaling arithmetic itself runs on the
device (``prep_kernel``), bit-identically to ``ennemi/_driver.py:882-883``.

Eligible tasks: continuous float64 variables, no mask, and no NaN dropping (``drop_nan`` with data
that has no NaNs is a no-op and stays eligible).  Everything else takes the general host path.
"""
from __future__ import annotations

import itertools
import threading
import warnings
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _align, _checks, _devices, _native

_next_key = itertools.count(1)
_key_lock = threading.Lock()


def _new_key() -> int:
    with _key_lock:
        return next(_next_key)


class _Scratch(threading.local):
    buf: Optional[np.ndarray] = None


_scratch = _Scratch()


def window_stats(view: np.ndarray) -> Tuple[float, float]:
    """``(view.mean(), view.std())`` with NumPy's own operation sequence (``_methods._mean`` /
    ``_methods._var``: pairwise ``add.reduce``, subtract, multiply, ``add.reduce``, sqrt), hence the
    same bits, but into a reused per-thread scratch buffer instead of fresh temporaries."""
    n = view.shape[0]
    mean = np.add.reduce(view) / n
    buf = _scratch.buf
    if buf is None or buf.shape[0] < n:
        buf = _scratch.buf = np.empty(max(n, 1 << 16))
    dev = buf[:n]
    np.subtract(view, mean, out=dev)
    np.multiply(dev, dev, out=dev)
    std = np.sqrt(np.add.reduce(dev) / n)
    return float(mean), float(std)


class NoiseBank:
    """The reference's fixed-seed noise draws (``_driver.py:874,883``), memoised on the host by
    ``_align._NoiseStream`` and mirrored on each GPU under a cache key.  Every :class:`ColumnStore` pins the
    keys it was handed until it is closed (its memoised descriptors keep referring to them without coming back
    here), so least-recently-used entries are dropped from the devices only once more than ``MAX_ENTRIES`` draw
    sequences are live AND no running call holds them."""

    _lock = threading.Lock()
    _keys: "OrderedDict[tuple, list]" = OrderedDict()       # draw-shape sequence -> [cache key, pin count]
    _on_device: Dict[int, set] = {}
    MAX_ENTRIES = 64

    @classmethod
    def key_for(cls, shapes: tuple, values: np.ndarray, dev: int, store: "Optional[ColumnStore]" = None) -> int:
        ordinal = _devices.ordinal(dev)
        with cls._lock:
            entry = cls._keys.get(shapes)
            if entry is None:
                entry = cls._keys[shapes] = [_new_key(), 0]
                if len(cls._keys) > cls.MAX_ENTRIES:
                    for old_shapes in [sh for sh, e in cls._keys.items() if e[1] == 0 and e is not entry]:
                        if len(cls._keys) <= cls.MAX_ENTRIES:
                            break
                        old = cls._keys.pop(old_shapes)[0]
                        for d, have in cls._on_device.items():
                            if old in have:
                                have.discard(old)
                                _native.cache_drop(old, dev=d)
            else:
                cls._keys.move_to_end(shapes)
            key = entry[0]
            if store is not None and store._pin_noise(shapes):
                entry[1] += 1
            have = cls._on_device.setdefault(ordinal, set())
            if key not in have:
                _native.cache_put(key, np.ravel(values), dev=dev)
                have.add(key)
            return key

    @classmethod
    def release(cls, shapes_list) -> None:
        with cls._lock:
            for shapes in shapes_list:
                entry = cls._keys.get(shapes)
                if entry is not None and entry[1] > 0:
                    entry[1] -= 1


class ColumnStore:
    """The variables of one API call, uploaded lazily to each GPU that runs one of its tasks.

    Columns of a row-major 2-D array (the ``(n, nvar)`` layout ``pairwise_mi`` and multi-variable
    ``estimate_mi`` receive) are registered as one *block*: the first task that needs any of them on a GPU
    uploads the whole block with one copy and the device de-interleaves it (``eb2_cache_put_block``) — the
    host never gathers strided columns."""

    # a block is uploaded whole only if its columns fill at least this fraction of its rows' stride
    MIN_BLOCK_DENSITY = 0.25

    def __init__(self, full_stats: bool = False):
        self._cols: Dict[int, np.ndarray] = {}
        self._on_device: Dict[int, set] = {}
        self._lock = threading.Lock()
        self._stats: Dict[tuple, tuple] = {}
        self._blocks: List[Tuple[np.ndarray, List[int]]] = []
        self._block_of: Dict[int, int] = {}
        self._descs: Dict[tuple, tuple] = {}
        self._noise_pins: set = set()
        self._inflight: Dict[tuple, threading.Event] = {}
        # whole-column mean/std of every block column computed right after the upload, in one call
        # (the windows of every pairwise_mi task; lagged windows are computed on demand)
        self.full_stats = full_stats
        # set by the API when every rank of a fan-out job is known to reach the upload (a collective)
        self.sharded_upload = False

    def add(self, column: np.ndarray) -> int:
        key = _new_key()
        self._cols[key] = np.ascontiguousarray(column, dtype=np.float64)
        return key

    def add_columns(self, block: np.ndarray) -> List[int]:
        """One key per column of a 2-D array (or the single key of a 1-D one)."""
        if block.ndim == 1:
            return [self.add(block)]
        ld = _native.block_layout(block)
        if ld is None or block.shape[1] < self.MIN_BLOCK_DENSITY * ld:
            return [self.add(block[:, j]) for j in range(block.shape[1])]
        keys = [_new_key() for _ in range(block.shape[1])]
        self._blocks.append((block, keys))
        for j, key in enumerate(keys):
            self._cols[key] = block[:, j]
            self._block_of[key] = len(self._blocks) - 1
        return keys

    def _pin_noise(self, shapes: tuple) -> bool:
        """True the first time this store is handed the noise vector of ``shapes`` (NoiseBank then counts a pin)."""
        with self._lock:
            if shapes in self._noise_pins:
                return False
            self._noise_pins.add(shapes)
            return True

    def ensure(self, dev: int, key: int) -> None:
        """Uploads ``key`` (or the block it belongs to) to ``dev`` unless it is there already.  The store lock is
        held only to decide who uploads: the copy itself runs unlocked (two lanes may upload different variables
        side by side), and a second caller of the same variable waits for the first one's upload."""
        ordinal = _devices.ordinal(dev)
        bid = self._block_of.get(key)
        unit = (ordinal, "b", bid) if bid is not None else (ordinal, "k", key)
        while True:
            with self._lock:
                have = self._on_device.setdefault(ordinal, set())
                if key in have:
                    return
                waiter = self._inflight.get(unit)
                if waiter is None:
                    mine = self._inflight[unit] = threading.Event()
                    break
            waiter.wait()
        try:
            stats = {}
            if bid is None:
                _native.cache_put(key, self._cols[key], dev=dev)
                done = [key]
            else:
                block, keys = self._blocks[bid]
                if not self._put_block_sharded(keys, block, dev):
                    _native.cache_put_block(keys, block, dev=dev)
                done = keys
                n = block.shape[0]
                if self.full_stats and n >= DEVICE_STATS_MIN_ROWS:
                    means, stds = _native.cache_stats_many(keys, [0] * len(keys), n, dev=dev)
                    stats = {(k, 0, n): (float(m), float(sd)) for k, m, sd in zip(keys, means, stds)}
            with self._lock:
                self._on_device.setdefault(ordinal, set()).update(done)
                for tag, value in stats.items():
                    self._stats.setdefault(tag, value)
        finally:
            with self._lock:
                self._inflight.pop(unit, None)
            mine.set()

    def _put_block_sharded(self, keys, block: np.ndarray, dev: int) -> bool:
        """One process per GPU with task fan-out on (every rank makes the same call with the same data): each rank
        uploads only its 1/G of the block's rows and the slices are all-gathered over NVLink, so the host-to-device
        copy - the largest fixed cost of a multi-GPU ``pairwise_mi`` - is shared out.  False: not applicable."""
        if not self.sharded_upload or block.shape[0] < 8 * 1024:
            return False
        from . import distributed
        if not distributed.task_fanout_enabled() or distributed._dist().get_backend() != "nccl":
            return False
        import torch
        rank, size = distributed.world()
        n, ncols = block.shape
        ld = _native.block_layout(block)
        if ld != ncols:
            return False
        per = -(-n // size)
        tdev = torch.device("cuda", _devices.ordinal(dev))
        mine = torch.zeros((per, ncols), dtype=torch.float64, device=tdev)
        lo, hi = min(rank * per, n), min((rank + 1) * per, n)
        if hi > lo:
            mine[: hi - lo].copy_(torch.from_numpy(block[lo:hi]), non_blocking=False)
        full = torch.empty((per * size, ncols), dtype=torch.float64, device=tdev)
        distributed._dist().all_gather_into_tensor(full, mine)
        torch.cuda.current_stream(tdev).synchronize()
        _native.cache_put_block_dev(keys, int(full.data_ptr()), n, ncols, dev=dev)
        return True

    def cached_desc(self, tag: tuple):
        return self._descs.get(tag)

    def remember_desc(self, tag: tuple, desc, drew: bool) -> None:
        self._descs[tag] = (desc, drew)

    def missing(self, dev: int, *keys) -> bool:
        with self._lock:
            have = self._on_device.get(_devices.ordinal(dev), set())
            return any(k not in have for k in keys)

    def has_stats(self, tag: tuple) -> bool:
        with self._lock:
            return tag in self._stats

    def stats(self, tag: tuple, compute):
        with self._lock:
            hit = self._stats.get(tag)
        if hit is None:
            hit = compute()
            with self._lock:
                self._stats[tag] = hit
        return hit

    def close(self) -> None:
        for dev, have in self._on_device.items():
            for key in have:
                try:
                    _native.cache_drop(key, dev=dev)
                except Exception:
                    pass
        NoiseBank.release(self._noise_pins)
        self._noise_pins = set()
        self._on_device.clear()
        self._cols.clear()
        self._blocks.clear()
        self._block_of.clear()
        self._descs.clear()


def eligible(arrays, mask, drop_nan: bool, discrete_any: bool) -> bool:
    if discrete_any or mask is not None:
        return False
    for a in arrays:
        if a is None:
            continue
        if a.dtype != np.float64:
            return False
    if drop_nan:
        for a in arrays:
            if a is not None and np.isnan(a).any():
                return False
    return True


class ColsTask:
    """A continuous (x, y[, cond]) task on cached columns; ``run()`` returns the estimate."""

    __slots__ = ("store", "xkey", "ykey", "zkeys", "xview", "yview", "cond", "lag", "hi", "lo", "cond_lag", "k",
                 "preprocess", "n_total", "single_use", "share_prepared")

    def __init__(self, store, xkey, ykey, zkeys, xview, yview, cond, lag, hi, lo, cond_lag, k, preprocess):
        self.store, self.xkey, self.ykey, self.zkeys = store, xkey, ykey, zkeys
        self.xview, self.yview, self.cond = xview, yview, cond
        self.lag, self.hi, self.lo, self.cond_lag, self.k, self.preprocess = lag, hi, lo, cond_lag, k, preprocess
        self.n_total = len(yview)
        self.single_use = False      # set by the API when the call consists of this one task
        # prepared (rescaled + sorted) variables are kept on the device for the other tasks of the call only where
        # descriptors recur (pairwise_mi: every column 63 times); a lag sweep's x windows are used once each, and
        # caching them would grow device memory with nlags * nvar (set by the API)
        self.share_prepared = False

    def describe(self, dev: int, in_call_stats: bool = False):
        """(descriptors, n) of this task for device ``dev``: uploads what is missing, computes (cached)
        window statistics, raises the reference's errors for impossible windows.  ``in_call_stats``: leave
        the x / y window statistics to the library call itself (``FLAG_DEVICE_STATS``; descriptors carry
        mean = NaN) instead of asking for them first."""
        lo_pad, hi_pad = max(self.hi, 0), min(self.lo, 0)            # as _align.lagged_windows
        n_tot = self.n_total
        if not in_call_stats and not self.zkeys and isinstance(self.lag, (int, np.integer)):
            # both variables already described in these roles by an earlier task of the call (pairwise_mi asks for
            # each column 63 times): two dictionary look-ups instead of the work below
            n_fast = n_tot + hi_pad - lo_pad
            ordinal_fast = _devices.ordinal(dev)
            hx = self.store.cached_desc(((self.xkey, int(lo_pad - self.lag), n_fast), (), ordinal_fast, self.preprocess))
            if hx is not None:
                hy = self.store.cached_desc(((self.ykey, int(lo_pad), n_fast), ((n_fast,),) if hx[1] else (), ordinal_fast,
                                             self.preprocess))
                if hy is not None and n_fast > self.k:
                    return [hx[0], hy[0]], n_fast
        xs = self.xview[lo_pad - self.lag: n_tot - self.lag + hi_pad]   # views: non-integer lags raise here
        ys = self.yview[lo_pad: n_tot + hi_pad]
        n = len(ys)
        if n <= self.k:
            raise ValueError(_checks.MSG_K_TOO_LARGE)
        x_off = int(lo_pad - self.lag)
        y_off = int(lo_pad)
        z_offs = [int(lo_pad - cl) for cl in self.cond_lag] if self.zkeys else []

        shapes: tuple = ()
        stream = _align._NoiseStream()
        descs: List[_native.ColDesc] = []

        ordinal = _devices.ordinal(dev)

        def one(key, off, view, tag):
            # the descriptor of a variable in a given role (= position in the draw sequence) is the same for every
            # task of the call that uses it there: pairwise_mi asks for each column 63 times
            nonlocal shapes
            memo = (tag, shapes, ordinal, self.preprocess)
            hit = None if in_call_stats else self.store.cached_desc(memo)
            if hit is not None:
                if hit[1]:
                    stream.skip((n,))
                    shapes = shapes + ((n,),)
                return hit[0]
            self.store.ensure(dev, key)
            mean, std, nkey = 0.0, 0.0, 0
            if self.preprocess and in_call_stats and not self.store.has_stats(tag):
                values = stream.normal((n,))
                shapes = shapes + ((n,),)
                return _native.ColDesc(key, off, 1, float("nan"), 1.0, NoiseBank.key_for(shapes, values, dev, self.store), 0, 1)
            reusable = True
            if self.preprocess:
                mean, std = self.store.stats(tag, stats_of(key, off, view))
                if abs(std) < _align.CONSTANT_STD:
                    warnings.warn(_align.CONSTANT_DATA_WARNING)
                    std = 0.0
                    reusable = False                                # every task warns, as in the reference
                elif std == std:                                    # NaN std (NaN input): reported by the device
                    values = stream.normal((n,))
                    shapes = shapes + ((n,),)
                    nkey = NoiseBank.key_for(shapes, values, dev, self.store)
                else:
                    std = 0.0
            desc = _native.ColDesc(key, off, 1, mean, std, nkey, 0, 1)
            if reusable and not in_call_stats:
                self.store.remember_desc(memo, desc, nkey != 0)
            return desc

        device_stats = n >= DEVICE_STATS_MIN_ROWS
        in_call_stats = in_call_stats and device_stats

        def stats_of(key, off, view):
            if device_stats:      # NumPy-exact pairwise mean/std where the column already is (eb2_cache_stats)
                return lambda: _native.cache_stats(key, off, n, dev=dev)
            return lambda: window_stats(view)

        if n >= OVERLAP_MIN_ROWS and self.store.missing(dev, self.xkey, self.ykey):
            # large first-time windows: the two uploads (and statistics, unless the library call computes them itself)
            # run side by side on two stream lanes - from pageable memory each is bound by one thread's memcpy
            want_stats = self.preprocess and not in_call_stats

            def side(key, off, view, lane_dev):
                self.store.ensure(lane_dev, key)
                if want_stats:
                    self.store.stats((key, off, n), (lambda: _native.cache_stats(key, off, n, dev=lane_dev)) if device_stats
                                     else (lambda: window_stats(view)))
            other = _devices.with_lane(dev, ((dev >> 8) + 1) % 4)
            fut = _helper_pool().submit(side, self.ykey, y_off, ys, other)
            side(self.xkey, x_off, xs, dev)
            fut.result()
        descs.append(one(self.xkey, x_off, xs, (self.xkey, x_off, n)))
        descs.append(one(self.ykey, y_off, ys, (self.ykey, y_off, n)))
        if self.zkeys:
            c = len(self.zkeys)
            zstd = zmean = None
            nkey = 0
            if self.preprocess:
                def zstats():
                    zs = np.column_stack([self.cond[o: o + n, j] for j, o in enumerate(z_offs)])
                    return zs.mean(axis=0), zs.std(axis=0)       # axis-0 reductions, exactly as :895-899
                zmean, zstd = self.store.stats(("z", tuple(self.zkeys), tuple(z_offs), n), zstats)
                if np.any(np.abs(zstd) < _align.CONSTANT_STD):
                    warnings.warn(_align.CONSTANT_DATA_WARNING)
                    zstd = None
                elif not np.any(np.isnan(zstd)):
                    values = stream.normal((n, c))
                    shapes = shapes + ((n, c),)
                    nkey = NoiseBank.key_for(shapes, values, dev, self.store)
                else:
                    zstd = None
            for j, (key, off) in enumerate(zip(self.zkeys, z_offs)):
                self.store.ensure(dev, key)
                if zstd is None:
                    descs.append(_native.ColDesc(key, off, 1, 0.0, 0.0, 0, 0, 1))
                else:
                    descs.append(_native.ColDesc(key, off, 1, float(zmean[j]), float(zstd[j]), nkey, j, c))
        return descs, n

    def run(self) -> float:
        dev = _devices.current()
        from . import distributed
        sharded = distributed.row_sharding_enabled()
        # a one-task call on a large window: upload, statistics, rescaling and the estimate in ONE library call
        descs, n = self.describe(dev, self.single_use and self.preprocess and not sharded)
        if any(d.mean != d.mean for d in descs):
            try:
                return self._estimate(dev, descs, n, _native.FLAG_SINGLE_USE | _native.FLAG_DEVICE_STATS)
            except _native.ConstantWindow:       # rare: the reference's constant-data warning path needs the value of std
                descs, n = self.describe(dev, False)
        return self._estimate(dev, descs, n, 0 if self.share_prepared and not self.single_use else _native.FLAG_SINGLE_USE)

    def _estimate(self, dev: int, descs, n: int, flags: int) -> float:
        try:
            from . import distributed
            if distributed.row_sharding_enabled():
                # one process per GPU, every rank holds the columns: shard the query rows, one all-reduce
                rank, size = distributed.world()
                lo, hi = distributed.shard_bounds(n, rank, size)
                part = distributed._all_reduce_sum(_native.mi_cols_rows(descs, n, self.k, lo, hi, dev=dev))
                if self.zkeys:
                    return _native.cmi_finish(part, n, self.k)
                return _native.ksg_mi_finish(part, n, self.k)
            if self.zkeys:
                return _native.cmi_cols(descs, n, self.k, dev=dev, flags=flags)
            return _native.ksg_mi_cols(descs, n, self.k, dev=dev, flags=flags)
        except _native.NonFiniteInput as e:
            if e.nan:
                raise ValueError(_checks.MSG_NANS_LEFT) from None
            raise ValueError(str(e)) from None


OVERLAP_MIN_ROWS = 200_000
DEVICE_STATS_MIN_ROWS = 50_000     # above this the window statistics are computed on the device
_pool = None
_pool_lock = threading.Lock()


def _helper_pool():
    global _pool
    with _pool_lock:
        if _pool is None:
            import concurrent.futures
            _pool = concurrent.futures.ThreadPoolExecutor(2, "ennemi-b200-prep")
        return _pool


def _check_status(status) -> None:
    for st in status:
        if st:
            code, data_flags = int(st) & 0xFF, int(st) >> 8
            if code == _native.ERR_NONFINITE:
                if data_flags & 1:
                    raise ValueError(_checks.MSG_NANS_LEFT)
                raise ValueError("data must be finite, check for nan or inf values")
            raise RuntimeError(f"ennemi_b200: batched estimate failed with code {code}")


def run_pairs(tasks: List["ColsTask"]) -> List[float]:
    """All unconditional column tasks of one window length in ONE library call (``eb2_ksg_mi_pairs``): every distinct
    (variable, role) descriptor is rescaled and sorted once on the device, the pairs share them."""
    dev = _devices.current()
    out: List[float] = []
    index: Dict[int, int] = {}
    cols: List[_native.ColDesc] = []
    pairs: List[Tuple[int, int]] = []
    n = None

    def flush() -> None:
        nonlocal index, cols, pairs
        if pairs:
            values, status = _native.ksg_mi_pairs(cols, np.asarray(pairs, dtype=np.int32), n, tasks[0].k, dev=dev)
            _check_status(status)
            out.extend(float(v) for v in values)
        index, cols, pairs = {}, [], []

    for task in tasks:
        descs, n_t = task.describe(dev)
        if n is not None and n_t != n:
            flush()                                  # (windows of another length: a call of their own)
        n = n_t
        # the prepared variables of one call stay on the device together: bound them (~40 bytes per row each)
        if len(cols) + 2 > max(2, PAIR_CALL_BYTES // (40 * max(n, 1))):
            flush()
        ids = []
        for d in descs:
            # memoised descriptors are the same object for every task that uses a variable in a role
            ident = index.get(id(d))
            if ident is None:
                ident = index[id(d)] = len(cols)
                cols.append(d)
            ids.append(ident)
        pairs.append((ids[0], ids[1]))
    flush()
    return out


PAIR_CALL_BYTES = 8 << 30     # device memory the prepared variables of one eb2_ksg_mi_pairs call may take


BATCH = 8     # tasks per native call: one interpreter round trip (and one GIL release) per batch


def run_batch(tasks: List["ColsTask"]) -> List[float]:
    """Estimates a run of same-shaped column tasks with one call into the library."""
    dev = _devices.current()
    described = [t.describe(dev) for t in tasks]
    n = described[0][1]
    values, status = _native.mi_cols_batch([d for d, _ in described], n, tasks[0].k, dev=dev,
                                           flags=0 if tasks[0].share_prepared else _native.FLAG_SINGLE_USE)
    _check_status(status)
    return [float(v) for v in values]
