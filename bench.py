#!/usr/bin/env python
"""Headline benchmark: KSG MI estimates/s at N = 10^6, k = 3 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one bivariate KSG estimate over a full 10^6-row synthetic Gaussian batch per GPU.  The metric is
a throughput, so with N > 1 (``torchrun``, one rank per GPU) independent same-shape estimates are fanned out,
one per GPU per step, with no data-path collective ("weak" scaling; value = N estimates / max-over-ranks
step time).  The other partition BASELINE.json configs[1] names — ONE estimate with its query rows sharded
over the GPUs and one NCCL all-reduce of the partial digamma sums — is timed in the same run and reported
under ``sharded``.

Numbers on the JSON line:
  value      estimates/s with the coordinates already resident in HBM (C ABI, EB2_FLAG_DEVICE_INPUT)
  e2e        the same metric through the public API ``ennemi_b200.estimate_mi(y, x, k=3)`` on HOST
             numpy buffers: host preprocessing + H2D + kernels + D2H inside the timed region
  roofline   the dominant kernel (knn_kernel2<4> + its leftover kernel, the k-NN search of the bivariate pipeline)
             against the MEASURED FP64 issue rate
  brute_force  the same step with EB2_FLAG_NO_PRUNE (every candidate tile visited): the kernel the
             FP64 roofline in SURVEY.md §8(d) is defined on
  cpu_baseline  the reference's CPU path (same SciPy cKDTree calls, via oracle/) on a bounded sample
``--impl reference`` times only that CPU path and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# rank 0 prints exactly ONE line on stdout: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "KSG MI estimates/sec at N=10^6, k=3"
UNIT = "estimates/s"
N_ROWS = 1_000_000
K_NEIGH = 3
RHO = 0.6
FP64_OPS_PER_PAIR = 4          # 2-D space: 2 subtractions + 2 compares (2d per pair)
SAMPLE_STRIDE = 50             # CPU baseline: every 50th row is queried, trees hold all rows
# dram__bytes_read.sum + dram__bytes_write.sum of knn_kernel2<4> at this workload, one ncu --set full capture
# (profiles/ncu_r02_summary.md: 22.0 MB read, 0.5 kB written; leftover_kernel2<4> adds 3.0 MB): the slot-ordered point
# set (16 MB) and the cell table (4 MB) are read once, everything else stays in the 126 MB L2
NCU_DRAM_BYTES_PER_LAUNCH = 22.0e6
# the workload both arms (`--impl ours` / `--impl reference`) run: identical `config` on both JSON lines
CONFIG = {"workload": "estimate_mi bivariate Gaussian rho=0.6, N=1,000,000, k=3 (BASELINE.json configs[1]); a step = one "
                      "complete estimate per GPU", "n": N_ROWS, "k": K_NEIGH, "rho": RHO, "seed": 0}


def make_data(n=N_ROWS, seed=0):
    rng = np.random.default_rng(seed)
    d = rng.multivariate_normal([0, 0], [[1, RHO], [RHO, 1]], size=n)
    return np.ascontiguousarray(d[:, 1]), np.ascontiguousarray(d[:, 0])     # estimate_mi(d[:,1], d[:,0])


_orig_make_data = make_data


def preprocessed(y, x):
    """The buffers the estimator sees inside estimate_mi(y, x): rescaled + fixed-seed noise."""
    from ennemi_b200 import _align
    xs, ys, _ = _align.rescaled(x, y, None, False, False)
    return xs, ys


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own SciPy path (oracle "scipy" backend) on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_step(xs, ys, k=K_NEIGH, stride=SAMPLE_STRIDE):
    """Extrapolated seconds of one full estimate on the reference's SciPy path (oracle/timing.py)."""
    from oracle import timing
    return timing.ksg_mi_seconds(xs, ys, k, stride)


def cpu_baseline_block(xs, ys, steps=1):
    secs = [cpu_step(xs, ys) for _ in range(steps)]
    best = min(secs)
    return {"value": 1.0 / best, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle SciPy backend = the reference's cKDTree calls (_entropy_estimators.py:100-110): "
                      f"3 trees on all {len(xs):,} rows, k-NN query + both marginal counts for every "
                      f"{SAMPLE_STRIDE}th row, query time scaled x{SAMPLE_STRIDE}; a single estimate is one thread "
                      f"in the reference (benchmarks/bench_large_sample_mi.py:6-7); host has {os.cpu_count()} cores",
            "seconds_per_estimate": best}


def cpu_pairwise_block(data, k=K_NEIGH, pairs_timed=2):
    """The reference's CPU path for the pairwise_mi leg, on a bounded sample: `pairs_timed` of the 2,016 variable pairs
    through the oracle's SciPy backend (= the reference's cKDTree calls on the preprocessed columns), scaled to the
    whole matrix on one core and on all the box's cores (the reference fans pairs out over cpu_count threads,
    ennemi/_driver.py:749)."""
    import oracle
    from ennemi_b200 import _align
    secs = []
    for j in range(1, pairs_timed + 1):
        xs, ys, _ = _align.rescaled(data[:, 0].copy(), data[:, j].copy(), None, False, False)
        t0 = time.perf_counter()
        oracle.ksg_mi(xs, ys, k, backend="scipy")
        secs.append(time.perf_counter() - t0)
    per_pair = min(secs)
    n_pairs = data.shape[1] * (data.shape[1] - 1) // 2
    cores = os.cpu_count() or 1
    return {"kind": "port", "seconds_per_pair": per_pair, "pairs_timed": pairs_timed, "cores": cores,
            "estimated_total_s_one_core": per_pair * n_pairs, "estimated_total_s_all_cores": per_pair * n_pairs / cores,
            "sample": f"{pairs_timed} of {n_pairs} pairs timed on one thread (SciPy cKDTree calls of the reference), scaled linearly"}


def pairwise_parity(data, pw, k=K_NEIGH, pairs=((0, 1), (5, 40), (62, 63))):
    """A few of the 2,016 pairs against the reference's SciPy calls on the same preprocessed columns."""
    import oracle
    from ennemi_b200 import _align
    worst = 0.0
    for i, j in pairs:
        xs, ys, _ = _align.rescaled(data[:, i].copy(), data[:, j].copy(), None, False, False)
        worst = max(worst, abs(oracle.ksg_mi(xs, ys, k, backend="scipy")["value"] - pw[i, j]))
    return {"pairs_checked": len(pairs), "max_abs_dmi": worst}


def _timed(fn, reps=5):
    fn()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def _steady(call, nat, dev, reps=3):
    """Device phase times of one native call in steady state: the call repeated, the fastest repetition's phases (the
    first call of a new shape also sizes the lane's workspace)."""
    best = None
    for _ in range(reps):
        part = call()
        ph = nat.last_timing(dev)
        if best is None or ph["total_ms"] < best[1]["total_ms"]:
            best = (part, ph)
    return best


def other_configs(nat, eb, dev, with_cpu):
    """BASELINE.json configs[0], [2], [4]: wall time of the public API call on host arrays (best of 5), device phase
    times of the estimator, the work done against the FP64 issue peak, and - on a bounded sample - the reference's
    SciPy calls on the box's host with eps / count / value parity."""
    import oracle
    from ennemi_b200 import _align
    rng = np.random.default_rng(0)
    peak = nat.measure_fp64_peak(dev)
    out = {}

    def roof(pairs, ops_per_pair, ms, brute_pairs):
        ach = pairs * ops_per_pair / (ms * 1e-3) * 1e-12
        return {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "ops_per_pair": ops_per_pair, "pairs_evaluated": pairs,
                "survey_8d_frac": brute_pairs * ops_per_pair / (ms * 1e-3) * 1e-12 / peak}

    # configs[0]: the reference's own CPU-runnable case, N = 10,000
    n = 10_000
    d = rng.multivariate_normal([0, 0], [[1, RHO], [RHO, 1]], size=n)
    y0, x0 = np.ascontiguousarray(d[:, 1]), np.ascontiguousarray(d[:, 0])
    t_api, mi0 = _timed(lambda: float(eb.estimate_mi(y0, x0, k=K_NEIGH)[0, 0]))
    leg = {"workload": "estimate_mi bivariate Gaussian rho=0.6, N=10,000, k=3", "api_ms": t_api * 1e3, "mi": mi0,
           "phase_ms": nat.last_timing(dev)}
    if with_cpu:
        xs, ys, _ = _align.rescaled(x0, y0, None, False, False)
        t0 = time.perf_counter(); want = oracle.ksg_mi(xs, ys, K_NEIGH, backend="scipy"); t1 = time.perf_counter()
        v, got = nat.ksg_mi(nat.pack_coords([xs, ys]), K_NEIGH, dev=dev, details=True)
        leg["cpu_reference"] = {"kind": "port", "seconds": t1 - t0, "cores": 1, "sample": "the whole estimate"}
        leg["parity"] = {"max_abs_deps": float(np.max(np.abs(got["eps"] - want["eps"]))),
                         "count_mismatches": int(np.sum(got["nx"] != want["nx"]) + np.sum(got["ny"] != want["ny"])),
                         "abs_dmi": abs(v - want["value"])}
    out["cfg0_n1e4"] = leg

    # configs[2]: Frenzel-Pompe CMI, 3-D condition, N = 200,000, 50 lags
    n = 200_000
    z = rng.normal(size=(n, 3)); x = rng.normal(size=n) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=n)
    lags = list(range(50))
    t_api, mi_l = _timed(lambda: eb.estimate_mi(y, x, lag=lags, k=K_NEIGH, cond=z), reps=2)
    co = nat.pack_coords([x, y, z])
    part, ph = _steady(lambda: nat.cmi_rows(co.ctypes.data, n, 3, K_NEIGH, 0, n, dev=dev), nat, dev)
    leg = {"workload": "estimate_mi(y, x, lag=range(50), k=3, cond=z) N=200,000, 3-D condition", "api_s": t_api,
           "api_ms_per_lag": t_api * 1e3 / len(lags), "mi_lag0": float(mi_l[0, 0]), "one_lag_phase_ms": ph,
           "roofline": roof(part[nat.P_PAIRS], 2 * 5, ph["knn_ms"] + ph["count_ms"], float(n) * n * 2)}
    if with_cpu:
        # one lag through the reference's SciPy calls: the four trees on all rows, every 16th row queried (scaled x16)
        stride = 16
        rows = np.arange(0, n, stride)
        xyz = np.column_stack((x, y, z)); xz = np.column_stack((x, z)); yz = np.column_stack((y, z))
        t0 = time.perf_counter()
        eps_w = oracle.kth_distance(xyz, K_NEIGH, query=xyz[rows], backend="scipy")
        rad = eps_w - 1e-12
        nxz_w = oracle.ball_count(xz, rad, query=xz[rows], backend="scipy")
        nyz_w = oracle.ball_count(yz, rad, query=yz[rows], backend="scipy")
        nz_w = oracle.ball_count(z, rad, query=z[rows], backend="scipy")
        t1 = time.perf_counter()
        v, got = nat.cmi(co, K_NEIGH, dev=dev, details=True)
        leg["cpu_reference"] = {"kind": "port", "seconds_per_lag": (t1 - t0) * stride, "cores": 1,
                                "estimated_sweep_s_all_cores": (t1 - t0) * stride * len(lags) / (os.cpu_count() or 1),
                                "sample": f"one lag, trees on all rows, every {stride}th row queried, time scaled x{stride} "
                                          "(tree builds included in the scaled time: an upper bound)"}
        leg["parity"] = {"rows_checked": int(len(rows)), "max_abs_deps": float(np.max(np.abs(got["eps"][rows] - eps_w))),
                         "count_mismatches": int(np.sum(got["nxz"][rows] != nxz_w) + np.sum(got["nyz"][rows] != nyz_w)
                                                 + np.sum(got["nz"][rows] != nz_w))}
    out["cfg2_cmi_50lags"] = leg

    # configs[4a]: Ross discrete-continuous MI, 16 classes, N = 500,000, k = 5
    n = 500_000
    yd = rng.integers(0, 16, n); xc = rng.normal(size=n) + 0.25 * yd
    t_api, mi_r = _timed(lambda: float(eb.estimate_mi(xc, yd, discrete_x=True, k=5)[0, 0]), reps=3)
    leg = {"workload": "estimate_mi(xc, yd, discrete_x=True, k=5) N=500,000, 16 classes", "api_ms": t_api * 1e3, "mi": mi_r,
           "phase_ms": nat.last_timing(dev)}
    if with_cpu:
        xs, _, _ = _align.rescaled(xc, yd, None, False, True)
        t0 = time.perf_counter(); want = oracle.semidiscrete_mi(xs, yd, 5, backend="scipy"); t1 = time.perf_counter()
        labels, inv = np.unique(yd, return_inverse=True)
        v, got = nat.ross_mi(nat.pack_coords([xs]), np.ascontiguousarray(inv, dtype=np.int32), len(labels), 5, dev=dev, details=True)
        leg["cpu_reference"] = {"kind": "port", "seconds": t1 - t0, "cores": 1, "sample": "the whole estimate"}
        leg["parity"] = {"max_abs_deps": float(np.max(np.abs(got["eps"] - want["eps"]))),
                         "count_mismatches": int(np.sum(got["n_full"] != want["n_full"])), "abs_dmi": abs(v - want["value"])}
    out["cfg4_ross"] = leg

    # configs[4b]: 4-D k-NN entropy, N = 500,000, k = 5
    cov = np.array([[1.0, 0.5, 0.2, 0.1], [0.5, 1.0, 0.3, 0.0], [0.2, 0.3, 1.0, -0.4], [0.1, 0.0, -0.4, 1.0]])
    d4 = rng.multivariate_normal(np.zeros(4), cov, size=n)
    t_api, h4 = _timed(lambda: float(eb.estimate_entropy(d4, k=5, multidim=True)), reps=3)
    co4 = nat.pack_coords([d4])
    part, ph = _steady(lambda: nat.entropy_rows(co4.ctypes.data, n, 4, 5, 0, n, dev=dev), nat, dev)
    leg = {"workload": "estimate_entropy(4-D Gaussian, k=5, multidim=True) N=500,000", "api_ms": t_api * 1e3, "entropy": h4,
           "phase_ms": ph, "roofline": roof(part[nat.P_PAIRS], 2 * 4, ph["knn_ms"], float(n) * n)}
    if with_cpu:
        t0 = time.perf_counter(); want = oracle.knn_entropy(d4, 5, backend="scipy"); t1 = time.perf_counter()
        v, got = nat.entropy(co4, 5, dev=dev, details=True)
        leg["cpu_reference"] = {"kind": "port", "seconds": t1 - t0, "cores": 1, "sample": "the whole estimate"}
        leg["parity"] = {"max_abs_deps": float(np.max(np.abs(got["dist"] - want["dist"]))), "count_mismatches": 0,
                         "abs_dvalue": abs(v - want["value"])}
    out["cfg4_entropy4d"] = leg
    return out


def run_reference(args):
    """The reference's CPU path, MEASURED: every step is one complete estimate (three cKDTrees on all 10^6 rows, the
    k-NN query and both ball counts for EVERY row — stride 1, nothing extrapolated), one thread, as the reference
    runs a single estimate.  ~9 s per step on the GPU box's host: `--steps 25` is ~4 minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    y, x = make_data()
    xs, ys = preprocessed(y, x)
    for _ in range(min(args.warmup, 1)):          # one warm-up estimate is enough for a CPU path (page-in, allocator)
        cpu_step(xs, ys, stride=1)
    secs = [cpu_step(xs, ys, stride=1) for _ in range(args.steps)]
    per = sum(secs) / len(secs)
    value = 1.0 / per
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(CONFIG),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "oracle SciPy backend = the reference's own cKDTree calls (_entropy_estimators.py:100-110) on ALL "
                                   f"{N_ROWS:,} rows per step (no sub-sampling, no scaling); a single estimate is one thread in the "
                                   f"reference (benchmarks/bench_large_sample_mi.py:6-7); warm-up steps run: {min(args.warmup, 1)}; "
                                   f"host has {os.cpu_count()} cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(", ") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, reasons, mx, pw = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx = float(r[1]); pw.append(float(r[2]))
            except ValueError:
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            busy = [v for v in sm if v > 0.5 * max(sm)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw) if pw else None)
        return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from ennemi_b200 import _native as nat, distributed as ebd, _devices
    import ennemi_b200 as eb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nat.require_device()
    # one process = one GPU, always: at N = 1 on a multi-GPU box the pairwise leg must not fan out over the other GPUs
    os.environ.setdefault("ENNEMI_B200_DEVICES", str(local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ebd.enable_row_sharding(True)

    # N = 1: the named batch.  N > 1: every rank gets its own batch of the same shape (seed = rank) for the
    # task fan-out legs; rank 0's batch (replicated) is the one that is row-sharded in the `sharded` leg.
    y, x = make_data()
    xs, ys = preprocessed(y, x)
    coords_host = nat.pack_coords([xs, ys])
    if world > 1:
        y_own, x_own = make_data(seed=rank)
        xs_own, ys_own = preprocessed(y_own, x_own)
        coords_own = torch.from_numpy(nat.pack_coords([xs_own, ys_own])).to(dev)
    else:
        y_own, x_own = y, x
    # e2e inputs: NumPy arrays in page-locked host memory (the bench contract's "pinned host memory"); the
    # public API takes them like any other array and the library's H2D copies run at DMA speed
    pinned_keep = []

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        v = t.numpy()
        v[...] = a
        pinned_keep.append(t)
        return v

    y_own, x_own = pinned(y_own), pinned(x_own)
    coords_dev = torch.from_numpy(coords_host).to(dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)            # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(fn, steps, warmup):
        """Each step bracketed by CUDA events on the current stream (every library call completes
        before it returns), L2 flushed between steps outside the brackets; max over ranks."""
        for _ in range(warmup):
            fn()
        total_ms, launches, knn_ms, pairs = 0.0, 0, 0.0, 0.0
        for _ in range(steps):
            flush.zero_()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            extra = fn()
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
            t = nat.last_timing(local)
            launches += t["launches"]
            knn_ms += t["knn_ms"]
            last["phases"] = t
            if extra is not None:
                pairs += extra
        barrier()
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, launches, knn_ms, pairs

    last = {}

    def step_sharded(flags=0):
        """ONE estimate, query rows sharded over the ranks, one all-reduce of the partial block."""
        rank_, size_ = ebd.world()
        lo, hi = ebd.shard_bounds(N_ROWS, rank_, size_)
        part = nat.ksg_mi_rows(int(coords_dev.data_ptr()), N_ROWS, K_NEIGH, lo, hi, dev=local,
                               flags=flags | nat.FLAG_DEVICE_INPUT)
        pairs = part[nat.P_PAIRS]
        total = ebd._all_reduce_sum(part)
        last["value"] = nat.ksg_mi_finish(total, N_ROWS, K_NEIGH)
        return pairs

    def step_fanout():
        """One full estimate per rank on its own resident batch, no collective (task fan-out)."""
        part = nat.ksg_mi_rows(int(coords_own.data_ptr()), N_ROWS, K_NEIGH, 0, N_ROWS, dev=local, flags=nat.FLAG_DEVICE_INPUT)
        last["value_own"] = nat.ksg_mi_finish(part, N_ROWS, K_NEIGH)
        return part[nat.P_PAIRS]

    def step_e2e():
        last["e2e"] = float(eb.estimate_mi(y_own, x_own, k=K_NEIGH)[0, 0])
        return None

    y_page, x_page = np.array(y_own), np.array(x_own)          # ordinary (pageable) NumPy arrays: what users pass

    def step_e2e_pageable():
        last["e2e_pageable"] = float(eb.estimate_mi(y_page, x_page, k=K_NEIGH)[0, 0])
        return None

    sampler = ClockSampler(local) if rank == 0 else None
    sharded = None
    e2e_extra = None
    if world == 1:
        ms_res, launches, knn_ms, pairs = timed_steps(step_sharded, args.steps, args.warmup)
        phases = dict(last["phases"])
        # a COLD call first: fresh pageable arrays of this shape for the first time in the process (noise vectors drawn and
        # uploaded, workspaces grown, nothing memoised), then the steady states from page-locked and from pageable memory
        yc, xc = make_data(seed=12345)
        t0 = time.perf_counter()
        cold_mi = float(eb.estimate_mi(yc, xc, k=K_NEIGH)[0, 0])
        cold_ms = (time.perf_counter() - t0) * 1e3
        ms_e2e, launches_e2e, _, _ = timed_steps(step_e2e, args.steps, args.warmup)
        ms_page, _, _, _ = timed_steps(step_e2e_pageable, args.steps, args.warmup)
        e2e_extra = {"pageable_ms_per_step": ms_page / args.steps, "pageable_value": 1e3 / (ms_page / args.steps),
                     "cold_first_call_ms": cold_ms, "cold_mi": cold_mi,
                     "note": "pageable: the same call on ordinary NumPy arrays (uploads staged through a page-locked ring by "
                             "the library); cold: the first call of the process on fresh arrays (includes drawing 2e6 PCG64 "
                             "normals for the reference's fixed-seed noise, workspace growth, first-touch of every buffer)"}
    else:
        # the metric is a throughput: on N GPUs independent estimates are fanned out, one per GPU per step
        # (weak scaling, no data-path collective); the row-sharded single estimate of configs[1] is timed beside it
        ebd.enable_row_sharding(False)
        ms_res, launches, knn_ms, pairs = timed_steps(step_fanout, args.steps, args.warmup)
        phases = dict(last["phases"])
        ms_e2e, launches_e2e, _, _ = timed_steps(step_e2e, args.steps, args.warmup)
        ebd.enable_row_sharding(True)
        ms_s, _, knn_s, _ = timed_steps(step_sharded, args.steps, args.warmup)
        sharded_mi, sharded_phases = last["value"], dict(last["phases"])
        # what the exchange alone costs: the same all-reduce of a partial block, nothing else
        blk = np.zeros(nat.P_LEN)
        for _ in range(5):
            ebd._all_reduce_sum(blk)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            ebd._all_reduce_sum(blk)
        reduce_ms = (time.perf_counter() - t0) / 20 * 1e3
        # the brute-force kernel (every pair evaluated: the FP64-roofline statement of the north star) row-sharded the same way
        brute_sh = None
        if not args.no_brute:
            ms_bs, _, knn_bs, _ = timed_steps(lambda: step_sharded(nat.FLAG_NO_PRUNE), 2, 1)
            brute_sh = {"ms_per_step": ms_bs / 2, "knn_ms": knn_bs / 2, "mi": last["value"],
                        "note": "EB2_FLAG_NO_PRUNE, query rows sharded over the ranks: 4e12 FP64 instructions / G per GPU"}
        sharded = {"value": 1e3 / (ms_s / args.steps), "unit": UNIT, "scaling": "strong", "ms_per_step": ms_s / args.steps,
                   "brute_force": brute_sh,
                   "knn_ms": knn_s / args.steps, "mi": sharded_mi, "phase_ms": sharded_phases,
                   "all_reduce_ms": reduce_ms,
                   "note": "configs[1]: ONE N=1e6 estimate, the x-buckets (and with them the query rows) sharded over the GPUs; every "
                           "rank builds the whole grid (the point set is replicated), searches and counts its own buckets, and one "
                           "NCCL all-reduce sums the partial blocks, whose digamma sum is an exact integer: the value is bit-identical "
                           "for every number of GPUs.  Latency of a single estimate; the replicated grid build bounds it below"}
    units = world          # estimates per step

    # second half of BASELINE.json's metric: pairwise_mi wall time, 64 variables x N = 100,000 (configs[3]),
    # through the public API on host arrays; with N > 1 the 2,016 pair tasks are dealt over the ranks
    pairwise = None
    if not args.no_pairwise:
        ebd.enable_row_sharding(False)
        ebd.enable_task_fanout(world > 1)
        data = np.random.default_rng(0).normal(size=(100_000, 64))
        eb.pairwise_mi(data[:, :8], k=K_NEIGH)                                   # warm-up (28 pairs)
        pw_s = float("inf")
        for _ in range(3):                                                       # best of three complete calls
            barrier()
            t0 = time.perf_counter()
            pw = eb.pairwise_mi(data, k=K_NEIGH)
            barrier()
            pw_s = min(pw_s, time.perf_counter() - t0)
        if world > 1:
            t = torch.tensor([pw_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pw_s = float(t.item())
        pairwise = {"metric": "pairwise_mi wall time, 64 vars", "value": pw_s, "unit": "s", "higher_is_better": False,
                    "pairs": 2016, "n": 100_000, "k": K_NEIGH, "scaling": "strong",
                    "max_offdiag_mi": float(np.nanmax(pw)),
                    "note": "ennemi_b200.pairwise_mi(data, k=3) on a pageable host (100000, 64) array, best of 3 calls: one "
                            "block upload, every variable rescaled and gridded once, all pairs of a rank in ONE library call "
                            "(eb2_ksg_mi_pairs: each stage one launch per batch of pairs)"
                            + (", pairs dealt over the ranks + one all_gather" if world > 1 else "")}
        ebd.enable_task_fanout(False)
        ebd.enable_row_sharding(True)
        if rank == 0 and world == 1 and not args.no_cpu:
            pairwise["cpu_reference"] = cpu_pairwise_block(data)
            pairwise["parity"] = pairwise_parity(data, pw)

    others = None
    if world == 1 and rank == 0 and not args.no_others:
        ebd.enable_row_sharding(False)
        others = other_configs(nat, eb, local, not args.no_cpu)
        ebd.enable_row_sharding(True)

    brute = None
    if world == 1 and not args.no_brute:
        bsteps = max(2, min(args.steps, 3))
        ms_b, _, knn_b, pairs_b = timed_steps(lambda: step_sharded(nat.FLAG_NO_PRUNE), bsteps, 1)
        brute = (ms_b / bsteps, knn_b / bsteps, pairs_b / bsteps, last["value"])

    clocks = sampler.stop() if sampler else None      # sampled across every timed leg above
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak = nat.measure_fp64_peak(local)                                      # 1e12 FP64 instr/s, measured live
    per_ms = ms_res / args.steps
    knn_per_ms = knn_ms / args.steps
    ops = pairs / args.steps * FP64_OPS_PER_PAIR                            # this rank's shard
    achieved = ops / (knn_per_ms * 1e-3) * 1e-12
    line = {
        "metric": METRIC, "value": units * 1e3 / per_ms, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict(CONFIG),
        "details": {"step": "N > 1: independent same-shape batches fanned out, one estimate per GPU per step, no collective; the "
                            "row-sharded single estimate is in `sharded`",
                    "algorithm": "exact search on a sort-free adaptive grid (bit-exact eps and counts); brute force in brute_force",
                    "l2": "flushed between steps (512 MiB write); inputs are 16 MB",
                    "parallelism": "single GPU" if world == 1 else f"task fan-out x{world} (+ rows/{world} in `sharded`)"},
        "mi": last["value"] if world == 1 else last["value_own"],
        "e2e": {"value": units * 1e3 / (ms_e2e / args.steps), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(coords_host.nbytes), "d2h_bytes_per_step": 44,
                "api": "ennemi_b200.estimate_mi(y, x, k=3) on host NumPy arrays in pinned memory (upload, mean/std, rescale+noise, estimate, result read-back)"
                       + ("" if world == 1 else "; one call per rank per step on its own arrays"),
                "mi": last.get("e2e")},
        "gpu_launches": int(launches),
        "phase_ms": phases,
        "layout_roofline": layout_roofline(phases),
        "roofline": {"bound": "fp64", "kernel": "knn_kernel2<4> + leftover_kernel2<4>", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                     "ops_per_launch": ops, "ms_per_launch": knn_per_ms,
                     # what actually bounds the kernel, from the ncu capture of this workload (profiles/ncu_r02_summary.md)
                     "ncu": {"issue_active_pct": 62.1, "lanes_active_per_instruction": 15.5, "warps_active_pct": 42.2,
                             "dram_read_mb": 22.0, "registers": 64},
                     "survey_8d_frac": (float(N_ROWS) * N_ROWS * FP64_OPS_PER_PAIR) / (knn_per_ms * 1e-3) * 1e-12 / peak,
                     "note": "FP64 CUDA-core issue bound (DADD+DSETP, 1 op = 1 FP64 instruction per lane); peak = DADD "
                             "issue rate measured live by eb2_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 figure); "
                             "ops = candidate pairs actually examined x 4 (search + leftover kernels, ms_per_launch = both, CUDA events "
                             "on the launching stream); the grid search examines ~25 candidates per row, one thread per query, and is "
                             "bound by instruction issue of the walk itself (62 % issue-active, 15.5 of 32 lanes active: "
                             "profiles/ncu_r02_summary.md), not by FP64 throughput - survey_8d_frac is the same time against SURVEY.md "
                             "8(d)'s brute-force count N^2 x 4, which a pruned algorithm legitimately exceeds; the FP64 roofline proper "
                             "(every pair evaluated) is in brute_force; traffic = dram bytes of knn_kernel2<4> per launch (ncu)"},
        "clocks": clocks,
    }
    if e2e_extra:
        line["e2e"].update(e2e_extra)
    if sharded:
        line["sharded"] = sharded
    if pairwise:
        line["pairwise"] = pairwise
    if others:
        line["configs"] = others
    if brute:
        b_ms, b_knn, b_pairs, b_val = brute
        b_ops = float(N_ROWS) * N_ROWS * FP64_OPS_PER_PAIR
        line["brute_force"] = {"value": 1e3 / b_ms, "ms_per_step": b_ms, "knn_ms": b_knn, "mi": b_val,
                               "roofline": {"bound": "fp64", "achieved": b_ops / (b_knn * 1e-3) * 1e-12, "peak": peak,
                                            "unit": "TFLOP/s", "frac": b_ops / (b_knn * 1e-3) * 1e-12 / peak,
                                            "ops_per_launch": b_ops}}
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_block(xs, ys, 1)
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def layout_roofline(phases):
    """The grid-build phase on the critical path against the HBM roofline.  Algorithmic bytes of one step on the main
    stream: x column into buckets (histogram: 8 B value read + 2 B bucket id written; scatter: 2 + 8 read, 4 B row + 8 B
    value written = 32 B/row) and the layout kernel (4 B row + 8 B x + 8 B y read, x, y, row, bucket id and cell table
    written = 46 B/row).  The y column's buckets and both columns' fine cells (another 72 B/row) are built on the second
    stream underneath the search and are not in `layout_ms`."""
    n = float(N_ROWS)
    bytes_step = (32 + 46) * n
    peak, src = 6528.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, KeyError, ValueError):
        pass
    ms = phases.get("layout_ms") or float("nan")
    achieved = bytes_step / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "phase": "grid build on the critical path (x buckets: sample, rank, splitters, histogram, scatter; layout)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "bytes_per_step": bytes_step, "ms": ms,
            "peak_source": src,
            "note": "seven short kernels (5-64 us each, the 16 MB working set lives in L2): bound by launch latency and the "
                    "dependent chain sample -> splitters -> histogram -> scatter -> layout, not by HBM"}


_REAL_STDOUT = None


def own_stdout():
    """Rank 0 prints exactly ONE line on stdout.  Libraries write there too (NCCL prints its version banner
    from C code at communicator creation whatever NCCL_DEBUG says), so file descriptor 1 is pointed at stderr for
    the duration of the run and the JSON line goes to the saved original descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    payload = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(payload.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, payload)


def main():
    own_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-brute", action="store_true", help="skip the brute-force (EB2_FLAG_NO_PRUNE) leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-pairwise", action="store_true", help="skip the pairwise_mi (64 variables) leg")
    ap.add_argument("--no-others", action="store_true", help="skip the legs of BASELINE.json configs[0], [2] and [4]")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
