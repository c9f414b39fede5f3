"""Which GPU a piece of work runs on.

One process per GPU is the intended deployment (``torchrun``: the rank's ``LOCAL_RANK`` GPU).  A
single Python process may also drive several GPUs of the box: the task scheduler binds each of
its worker threads to one device through :func:`use`.
"""
from __future__ import annotations

import contextlib
import os
import threading
from typing import List

from . import _native

_tls = threading.local()


def default_device() -> int:
    """``ENNEMI_B200_DEVICE`` if set, else the torchrun ``LOCAL_RANK``, else 0."""
    env = os.environ.get("ENNEMI_B200_DEVICE")
    if env is not None:
        return int(env)
    if "LOCAL_RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        n = max(_native.device_count(), 1)
        return int(os.environ["LOCAL_RANK"]) % n
    return 0


def current() -> int:
    """Device of the calling thread."""
    return getattr(_tls, "dev", None) if getattr(_tls, "dev", None) is not None else default_device()


@contextlib.contextmanager
def use(dev: int):
    """Binds the calling thread to ``dev`` for the duration of the block."""
    prev = getattr(_tls, "dev", None)
    _tls.dev = dev
    try:
        yield
    finally:
        _tls.dev = prev


LANES = 4    # concurrent streams per GPU used by the task fan-out (all the library has: measured best for pairwise_mi)


def ordinal(dev: int) -> int:
    """CUDA device ordinal of a (device | lane << 8) id."""
    return dev & 0xFF


def with_lane(dev: int, lane: int) -> int:
    return ordinal(dev) | ((lane % 4) << 8)


def visible() -> List[int]:
    """Devices a task fan-out may use from this process.

    ``ENNEMI_B200_DEVICES=0,2,3`` selects explicitly.  Under ``torchrun`` (one process per GPU) a
    process owns only its own GPU; otherwise every device of the box.
    """
    env = os.environ.get("ENNEMI_B200_DEVICES")
    if env:
        return [int(t) for t in env.split(",") if t.strip() != ""]
    if getattr(_tls, "dev", None) is not None or "ENNEMI_B200_DEVICE" in os.environ:
        return [current()]
    if "LOCAL_RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return [default_device()]
    return list(range(max(_native.device_count(), 1)))
