"""Fan-out of independent estimation tasks over the GPUs of the box.

Replaces the reference's CPU thread pool (``_map_maybe_parallel``, ``ennemi/_driver.py:736-785``)
and keeps its contract: results come back in task order, ``callback(i)`` fires once per finished
task (from the worker that ran it), ``max_threads`` bounds the concurrency, and a workload that is
too small to be worth it runs inline on the calling thread.

Each worker thread is bound to one GPU and one of that GPU's stream lanes.  Several workers share
a GPU so that one task's host-side preparation overlaps another's kernels and so that tasks too
small to fill 148 SMs run side by side on separate streams.  A task's value never depends on which worker ran it (the device
reduction is a fixed tree), so results are bitwise reproducible.
"""
from __future__ import annotations

import concurrent.futures
from typing import Callable, List, Optional, Sequence, TypeVar

from . import _devices

T = TypeVar("T")

WORKERS_PER_DEVICE = _devices.LANES
INLINE_BUDGET_S = 0.05   # run inline when the whole job is estimated below this


def gpu_time_estimate(n: int, n_cond: int, k: int) -> float:
    """Rough seconds per task on one B200 (launch-latency floor plus a small per-row cost)."""
    return 2.5e-4 + n * (1.5e-8 + 1.0e-8 * n_cond) * (1.0 + 0.02 * k)


def run_tasks(func: Callable[[T], float], params: Sequence[T], max_threads: Optional[int],
              time_estimate: float, callback: Callable[[int], None]) -> List[float]:
    from . import distributed
    if distributed.task_fanout_enabled():
        # one process per GPU: deal the tasks over the ranks, run this rank's share on its own GPU
        # (threads + stream lanes as below), then all-gather the scalars
        return distributed.fan_out(func, params, callback,
                                   runner=lambda f, p, cb: _run_local(f, p, max_threads, time_estimate, cb))
    return _run_local(func, params, max_threads, time_estimate, callback)


def _run_local(func: Callable[[T], float], params: Sequence[T], max_threads: Optional[int],
               time_estimate: float, callback: Callable[[int], None]) -> List[float]:
    from . import _columns
    from . import distributed as _dist
    if len(params) > 1 and not _dist.row_sharding_enabled() and all(isinstance(p, _columns.ColsTask) for p in params):
        return _run_column_batches(list(params), max_threads, time_estimate, callback)
    devices = _devices.visible()
    workers = max(1, len(devices)) * WORKERS_PER_DEVICE
    if len(params) * time_estimate < INLINE_BUDGET_S:
        workers = 1
    if max_threads is not None:
        workers = min(workers, max_threads)
    workers = min(workers, max(len(params), 1))
    if _dist.row_sharding_enabled():
        workers = 1          # every estimate is a collective: all ranks must run the tasks in the same order

    if workers <= 1:
        out = []
        for i, p in enumerate(params):
            out.append(func(p))
            callback(i)
        return out

    results: List[float] = [float("nan")] * len(params)

    def work(i: int, dev: int) -> None:
        with _devices.use(dev):
            results[i] = func(params[i])
        callback(i)

    with concurrent.futures.ThreadPoolExecutor(workers, "ennemi-b200-work") as pool:
        # task i -> device i mod G; consecutive tasks of one device rotate over its stream lanes
        futures = [pool.submit(work, i, _devices.with_lane(devices[i % len(devices)], (i // len(devices)) % _devices.LANES))
                   for i in range(len(params))]
        concurrent.futures.wait(futures)
        for f in futures:
            f.result()          # re-raise the first failure, like the reference's done-callback does
    return results


def _run_column_batches(tasks, max_threads: Optional[int], time_estimate: float,
                        callback: Callable[[int], None]) -> List[float]:
    """Column tasks (device-resident variables) go to the library in batches: consecutive batches
    rotate over the (GPU, lane) workers, one native call per batch."""
    from . import _columns
    devices = _devices.visible()
    if all(not t.zkeys for t in tasks):
        return _run_pair_calls(tasks, devices, max_threads, callback)
    workers = max(1, len(devices)) * WORKERS_PER_DEVICE
    if len(tasks) * time_estimate < INLINE_BUDGET_S:
        workers = 1
    if max_threads is not None:
        workers = min(workers, max_threads)
    size = max(1, min(_columns.BATCH, -(-len(tasks) // max(workers, 1))))
    batches = [list(range(lo, min(lo + size, len(tasks)))) for lo in range(0, len(tasks), size)]
    workers = max(1, min(workers, len(batches)))
    results: List[float] = [float("nan")] * len(tasks)

    def work(b: int, dev: int) -> None:
        with _devices.use(dev):
            values = _columns.run_batch([tasks[i] for i in batches[b]])
        for i, v in zip(batches[b], values):
            results[i] = v
            callback(i)

    if workers <= 1:
        for b in range(len(batches)):
            work(b, _devices.current())
        return results
    with concurrent.futures.ThreadPoolExecutor(workers, "ennemi-b200-work") as pool:
        futures = [pool.submit(work, b, _devices.with_lane(devices[b % len(devices)], (b // len(devices)) % _devices.LANES))
                   for b in range(len(batches))]
        concurrent.futures.wait(futures)
        for f in futures:
            f.result()
    return results


def _run_pair_calls(tasks, devices, max_threads: Optional[int], callback: Callable[[int], None]) -> List[float]:
    """Unconditional column tasks: each GPU takes one contiguous share of the task list and estimates it with ONE
    library call (``eb2_ksg_mi_pairs``: every variable rescaled and sorted once, all pairs batched per stage)."""
    from . import _columns
    parts = max(1, min(len(devices), len(tasks)))
    if max_threads is not None:
        parts = max(1, min(parts, max_threads))
    bounds = [(len(tasks) * p) // parts for p in range(parts + 1)]
    results: List[float] = [float("nan")] * len(tasks)

    def work(p: int, dev: int) -> None:
        lo, hi = bounds[p], bounds[p + 1]
        with _devices.use(dev):
            values = _columns.run_pairs(tasks[lo:hi])
        for i, v in zip(range(lo, hi), values):
            results[i] = v
            callback(i)

    if parts == 1:
        work(0, devices[0] if devices else _devices.current())
        return results
    with concurrent.futures.ThreadPoolExecutor(parts, "ennemi-b200-work") as pool:
        futures = [pool.submit(work, p, devices[p]) for p in range(parts)]
        concurrent.futures.wait(futures)
        for f in futures:
            f.result()
    return results
