"""The N > 1 host logic on CPU: two processes, ``gloo`` backend, world_size 2.

The per-rank device work is replaced by the CPU oracle (this is a test of the sharding / all-reduce
/ all-gather plumbing, not of the kernels): each rank produces the partial block for its row shard,
the blocks are summed with one all-reduce, and every rank must finish the same value as a single
process — and the task fan-out must rebuild the task-ordered result list on every rank."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import oracle
    from ennemi_b200 import _native as nat, distributed as ebd, _schedule, _devices
    import ennemi_b200 as eb

    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    n, k = 3_000, 3
    d = rng.multivariate_normal([0, 0], [[1, 0.6], [0.6, 1]], size=n)
    coords = nat.pack_coords([d[:, 0], d[:, 1]])
    full = oracle.ksg_mi(d[:, 0], d[:, 1], k, backend="scipy")

    def fake_rows(ptr, n_, k_, lo, hi, dev=0, flags=0):
        # what eb2_ksg_mi_rows returns for rows [lo, hi): raw digamma sum + zero counters + row count
        part = np.zeros(nat.P_LEN)
        part[nat.P_SUM] = np.sum(oracle.psi(full["nx"][lo:hi]) + oracle.psi(full["ny"][lo:hi]))
        part[nat.P_ROWS] = hi - lo
        return part

    nat.ksg_mi_rows = fake_rows
    nat.ksg_mi_finish = lambda part, n_, k_: float(oracle.psi(np.array([n_]))[0] + oracle.psi(np.array([k_]))[0]
                                                   - part[nat.P_SUM] / n_) if part[nat.P_ROWS] == n_ else float("nan")
    nat.device_count = lambda: 1
    sharded = ebd.sharded_ksg_mi(coords, k)

    # task fan-out through the public API: tasks dealt over the ranks, scalars all-gathered
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import OracleBackend
    fake = OracleBackend()
    for name in OracleBackend.PATCHED:
        setattr(nat, name, getattr(fake, name))
    _devices.visible = lambda: [0]
    x = rng.normal(size=(400, 5))
    seen = []
    ebd.enable_task_fanout(True)
    pw = eb.pairwise_mi(x, callback=lambda i, j: seen.append((i, j)))
    lagged = eb.estimate_mi(x[:, 0], x[:, 1:3], lag=[0, 1, 2])
    # the same calls on device-resident columns (the (n, nvar) array travels as one block per rank)
    from ennemi_b200 import api
    api.DEVICE_COLUMNS_MIN_ROWS = 0
    pw_cols = eb.pairwise_mi(x)
    lagged_cols = eb.estimate_mi(x[:, 0], x[:, 1:3], lag=[0, 1, 2])
    api.DEVICE_COLUMNS_MIN_ROWS = 10 ** 9
    ebd.enable_task_fanout(False)
    # a failure in ONE rank's share (task 3 runs on rank 1 only) must surface as the same exception type on
    # every rank after the collective, not as a hang of the ranks that had nothing to complain about
    def touchy(p):
        if p == 3:
            raise ValueError("input contains NaNs")
        return float(p)
    try:
        ebd.fan_out(touchy, list(range(6)))
        agreed = "no error"
    except ValueError as e:
        agreed = "ValueError:" + str(e)
    clean = ebd.fan_out(float, list(range(6)))          # the group is still usable afterwards
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), sharded=sharded, want=full["value"], pw=pw, lagged=lagged,
             pw_cols=pw_cols, lagged_cols=lagged_cols, block_puts=getattr(fake, "block_puts", 0),
             agreed=agreed, clean=np.array(clean), seen=np.array(seen), bounds=np.array(ebd.shard_bounds(n, rank, world)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_fanout(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    # row shards tile [0, n); the all-reduced value equals the single-process one on both ranks
    assert r0["bounds"][0] == 0 and r0["bounds"][1] == r1["bounds"][0] and r1["bounds"][1] == 3_000
    assert abs(float(r0["sharded"]) - float(r0["want"])) < 1e-12
    assert float(r0["sharded"]) == float(r1["sharded"])
    # fan-out: both ranks hold the full, identical, task-ordered results; each ran half of the tasks
    assert np.array_equal(r0["pw"], r1["pw"], equal_nan=True) and np.array_equal(r0["lagged"], r1["lagged"])
    assert not np.isnan(r0["pw"][np.triu_indices(5, 1)]).any() and not np.isnan(r0["lagged"]).any()
    assert len(r0["seen"]) + len(r1["seen"]) == 10 and len(r0["seen"]) == 5
    # the column path (block upload per rank) gives the same bits as the general host path, on both ranks
    for r in (r0, r1):
        assert np.array_equal(r["pw_cols"], r0["pw"], equal_nan=True) and np.array_equal(r["lagged_cols"], r0["lagged"])
        assert int(r["block_puts"]) >= 1
    # a one-rank failure is raised on both ranks (rank 1 its own exception, rank 0 a copy naming rank 1)
    assert str(r1["agreed"]) == "ValueError:input contains NaNs"
    assert str(r0["agreed"]).startswith("ValueError:input contains NaNs") and "rank 1" in str(r0["agreed"])
    assert np.array_equal(r0["clean"], np.arange(6.0)) and np.array_equal(r1["clean"], np.arange(6.0))


def test_single_process_reference_for_fanout(oracle_backend, tmp_path):
    """Same data as the two-rank run, one process: the gathered results must be these values."""
    import ennemi_b200 as eb
    rng = np.random.default_rng(0)
    rng.multivariate_normal([0, 0], [[1, 0.6], [0.6, 1]], size=3_000)
    x = rng.normal(size=(400, 5))
    pw = eb.pairwise_mi(x)
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    assert np.array_equal(r0["pw"], pw, equal_nan=True)
    assert np.array_equal(r0["lagged"], eb.estimate_mi(x[:, 0], x[:, 1:3], lag=[0, 1, 2]))


def test_shard_bounds_tile_exactly():
    from ennemi_b200 import distributed as ebd
    for n in (1, 7, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            b = [ebd.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
