"""Multi-GPU operation with one process per GPU (``torchrun`` / ``torch.distributed``).

Two partitions of the work, matching the two ways the path shards:

* **row sharding of one large estimate** — every rank holds the full point set, estimates the query
  rows of its own shard and contributes an 8-double block of raw sums; one ``all_reduce(SUM)``
  (NCCL over NVLink on GPUs, gloo in CPU tests) combines them and every rank finishes the same
  value.  This is the only collective on the data path.
* **task fan-out** — independent (variable pair / lag) tasks are dealt round-robin to the ranks, each
  rank runs its share on its own GPU, and one ``all_gather`` of the scalar results rebuilds the
  task-ordered list on every rank.  No data-path collective.

``torch`` is imported lazily: the single-process API does not need it.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _devices, _native

_enabled = False


def _dist():
    import torch.distributed as dist
    return dist


def is_initialized() -> bool:
    try:
        dist = _dist()
    except Exception:
        return False
    return dist.is_available() and dist.is_initialized()


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) outside ``torch.distributed``."""
    if not is_initialized():
        return 0, 1
    dist = _dist()
    return dist.get_rank(), dist.get_world_size()


def enable_task_fanout(on: bool = True) -> None:
    """Make ``estimate_mi`` / ``pairwise_mi`` deal their tasks across the ranks of the default
    process group (every rank must then make the same call with the same data)."""
    global _enabled
    _enabled = on


_row_sharding = False
_reduce_buffers: dict = {}


def enable_row_sharding(on: bool = True) -> None:
    """Make every single KSG / CMI / entropy estimate shard its query rows over the ranks (every
    rank must make the same call with the same data).  Use for a few very large estimates; for
    many small tasks prefer :func:`enable_task_fanout`."""
    global _row_sharding
    _row_sharding = on


def row_sharding_enabled() -> bool:
    return _row_sharding and is_initialized() and world()[1] > 1


def task_fanout_enabled() -> bool:
    return _enabled and is_initialized() and world()[1] > 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous row range of ``rank``; the ranges of all ranks tile [0, n) exactly."""
    return (n * rank) // world_size, (n * (rank + 1)) // world_size


def _all_reduce_sum(block: np.ndarray, group=None) -> np.ndarray:
    """Sum of an fp64 block over the ranks.  With the NCCL backend the block is reduced on the
    rank's GPU (over NVLink/NVSwitch); with gloo on the host."""
    if not is_initialized() or world()[1] == 1:
        return block
    import torch
    dist = _dist()
    backend = dist.get_backend(group)
    block = np.ascontiguousarray(block, dtype=np.float64)
    if backend == "nccl":
        # page-locked staging on both sides of the collective, kept across calls (an estimate is a few hundred
        # microseconds: allocations and pageable copies would show)
        ordinal = _devices.ordinal(_devices.current())
        bufs = _reduce_buffers.get((ordinal, block.size))
        if bufs is None:
            dev = torch.device("cuda", ordinal)
            bufs = _reduce_buffers[(ordinal, block.size)] = (
                torch.empty(block.size, dtype=torch.float64, pin_memory=True),
                torch.empty(block.size, dtype=torch.float64, device=dev))
        host, device = bufs
        host.numpy()[...] = block.ravel()
        device.copy_(host, non_blocking=True)
        dist.all_reduce(device, op=dist.ReduceOp.SUM, group=group)
        host.copy_(device, non_blocking=True)
        torch.cuda.current_stream(device.device).synchronize()
        return host.numpy().reshape(block.shape).copy()
    t = torch.from_numpy(block.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()


def _coords_ptr(coords) -> Tuple[int, int]:
    """(address, extra flags) of a (d, n) block given as a numpy array or a CUDA torch tensor."""
    if isinstance(coords, np.ndarray):
        return coords.ctypes.data, 0
    return int(coords.data_ptr()), _native.FLAG_DEVICE_INPUT      # torch tensor resident on this rank's GPU


def sharded_ksg_mi(coords, k: int = 3, flags: int = 0, group=None) -> float:
    """KSG MI of ``coords = [x; y]`` (2 x n, replicated on every rank) with query rows sharded
    over the ranks; one sum-allreduce of the partial block.  ``coords`` may be a contiguous numpy
    array (host) or a CUDA ``torch`` tensor on the rank's device.

    The partial block carries the digamma sum as exact integer limbs (``EB2_P_FIX0``), which add without
    rounding: the value is bit-identical for every number of ranks."""
    rank, size = world()
    n = int(coords.shape[1])
    lo, hi = shard_bounds(n, rank, size)
    ptr, extra = _coords_ptr(coords)
    part = _native.ksg_mi_rows(ptr, n, k, lo, hi, dev=_devices.current(), flags=flags | extra)
    return _native.ksg_mi_finish(_all_reduce_sum(part, group), n, k)


def sharded_cmi(coords, k: int = 3, flags: int = 0, group=None) -> float:
    """Frenzel-Pompe CMI of ``coords = [x; y; z...]`` with query rows sharded over the ranks."""
    rank, size = world()
    d, n = int(coords.shape[0]), int(coords.shape[1])
    lo, hi = shard_bounds(n, rank, size)
    ptr, extra = _coords_ptr(coords)
    part = _native.cmi_rows(ptr, n, d - 2, k, lo, hi, dev=_devices.current(), flags=flags | extra)
    return _native.cmi_finish(_all_reduce_sum(part, group), n, k)


def sharded_entropy(coords, k: int = 3, flags: int = 0, group=None) -> float:
    """k-NN entropy of ``coords`` (m x n) with query rows sharded over the ranks."""
    rank, size = world()
    m, n = int(coords.shape[0]), int(coords.shape[1])
    lo, hi = shard_bounds(n, rank, size)
    ptr, extra = _coords_ptr(coords)
    part = _native.entropy_rows(ptr, n, m, k, lo, hi, dev=_devices.current(), flags=flags | extra)
    return _native.entropy_finish(_all_reduce_sum(part, group), n, m, k)


def fan_out(func: Callable, params: Sequence, callback: Optional[Callable[[int], None]] = None,
            group=None, runner: Optional[Callable] = None) -> List[float]:
    """Runs ``func(params[i])`` for the tasks ``i = rank, rank + world, ...`` on this rank and returns
    the full task-ordered result list on every rank (one all_gather of fp64 scalars).  ``runner(func,
    params, callback)`` executes this rank's share (default: sequentially).

    Failures are agreed on before anybody raises: a rank whose share failed (a NaN column in one of its
    pairs, ``k`` too large for one window, a CUDA error) still enters the collective with an error flag,
    and every rank then raises — the failing rank its own exception, the others a copy of the first
    failing rank's — so that a data-dependent error is never a multi-rank hang."""
    rank, size = world()
    mine = list(range(rank, len(params), size))
    slots = len(range(0, len(params), size)) if size else 0
    local = np.full(slots + 1, np.nan)          # last slot: 1.0 when this rank's share failed

    def local_cb(slot: int) -> None:
        if callback is not None:
            callback(mine[slot])

    error: Optional[BaseException] = None
    try:
        if runner is None:
            values = []
            for slot, i in enumerate(mine):
                values.append(func(params[i]))
                local_cb(slot)
        else:
            values = runner(func, [params[i] for i in mine], local_cb)
        local[:len(mine)] = values
    except Exception as e:                       # noqa: BLE001 - re-raised below, after the ranks agreed
        if size == 1:
            raise
        error = e
    local[slots] = 0.0 if error is None else 1.0
    if size == 1:
        return [float(v) for v in local[:len(params)]]
    import torch
    dist = _dist()
    backend = dist.get_backend(group)
    t = torch.from_numpy(local)
    if backend == "nccl":
        t = t.to(torch.device("cuda", _devices.ordinal(_devices.current())))
    gathered = [torch.empty_like(t) for _ in range(size)]
    dist.all_gather(gathered, t, group=group)
    table = np.stack([g.cpu().numpy() for g in gathered])          # [rank][slot]
    failed = [r for r in range(size) if table[r, slots] != 0.0]
    if failed:
        # every rank reaches this point: exchange what went wrong, then raise everywhere
        info = [None] * size
        mine_info = None if error is None else (type(error).__name__, str(error))
        dist.all_gather_object(info, mine_info, group=group)
        if error is not None:
            raise error
        kind, message = info[failed[0]] or ("RuntimeError", "unknown failure")
        exc = ValueError if kind == "ValueError" else RuntimeError
        raise exc(f"{message} (raised as {kind} on rank {failed[0]})")
    return [float(table[i % size, i // size]) for i in range(len(params))]
