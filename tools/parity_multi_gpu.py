"""N-GPU parity run (torchrun, one rank per GPU, NCCL): every multi-GPU mode of the path against the same estimate on
one GPU.  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/parity_multi_gpu.py > profiles/parity_r02_n2.json

Checked: (a) row-sharded KSG estimate of BASELINE.json configs[1] (N = 10^6): BITWISE equal to the unsharded value (the
digamma sum travels as exact integer limbs); (b) row-sharded Frenzel-Pompe CMI and 4-D entropy: within 1e-10 of the
unsharded values (their partial sums are doubles); (c) pairwise_mi with the pairs fanned out over the ranks and the
block upload shared out + all-gathered over NVLink: bitwise equal to the single-GPU matrix; (d) a lag sweep fanned
out; (e) error agreement: a NaN column raises the same ValueError on every rank."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG", "WARN")


def main():
    import torch
    import torch.distributed as dist
    from ennemi_b200 import _native as nat, distributed as ebd
    import ennemi_b200 as eb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    rng = np.random.default_rng(0)

    # (a) configs[1], row-sharded
    n = 1_000_000
    d = rng.multivariate_normal([0, 0], [[1, 0.6], [0.6, 1]], size=n)
    co = nat.pack_coords([d[:, 0], d[:, 1]])
    co_dev = torch.from_numpy(co).to(dev)
    single = nat.ksg_mi(co, 3, dev=local)
    sharded_host = ebd.sharded_ksg_mi(co, 3)
    sharded_dev = ebd.sharded_ksg_mi(co_dev, 3)
    out["ksg_n1e6"] = {"single": single, "sharded_host_input": sharded_host, "sharded_device_input": sharded_dev,
                       "bitwise_equal": bool(single == sharded_host == sharded_dev)}

    # (b) CMI and 4-D entropy, row-sharded
    m = 100_000
    z = rng.normal(size=(m, 3)); x = rng.normal(size=m) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=m)
    cc = nat.pack_coords([x, y, z])
    s_c, p_c = nat.cmi(cc, 3, dev=local), ebd.sharded_cmi(cc, 3)
    e4 = nat.pack_coords([rng.normal(size=(m, 4))])
    s_e, p_e = nat.entropy(e4, 3, dev=local), ebd.sharded_entropy(e4, 3)
    out["cmi_n1e5"] = {"single": s_c, "sharded": p_c, "abs_diff": abs(s_c - p_c)}
    out["entropy4d_n1e5"] = {"single": s_e, "sharded": p_e, "abs_diff": abs(s_e - p_e)}

    # (c) pairwise_mi fanned out (sharded block upload + all-gather) vs one GPU
    data = rng.normal(size=(50_000, 16)) @ rng.normal(size=(16, 16))
    ebd.enable_task_fanout(False)
    pw_single = eb.pairwise_mi(data)
    ebd.enable_task_fanout(True)
    pw_fan = eb.pairwise_mi(data)
    lag_fan = eb.estimate_mi(data[:, 0], data[:, 1:4], lag=[0, 1, 2, 5])
    ebd.enable_task_fanout(False)
    lag_single = eb.estimate_mi(data[:, 0], data[:, 1:4], lag=[0, 1, 2, 5])
    out["pairwise_16x5e4"] = {"bitwise_equal": bool(np.array_equal(pw_single, pw_fan, equal_nan=True)),
                              "max_abs_diff": float(np.nanmax(np.abs(pw_single - pw_fan)))}
    out["lag_sweep"] = {"bitwise_equal": bool(np.array_equal(lag_single, lag_fan))}

    # (e) a NaN column: the same ValueError on every rank, no hang
    bad = data.copy(); bad[7, 3] = np.nan
    ebd.enable_task_fanout(True)
    try:
        eb.pairwise_mi(bad)
        verdict = "no error"
    except ValueError as e:
        verdict = "ValueError: " + str(e)[:60]
    ebd.enable_task_fanout(False)
    verdicts = [None] * world
    if world > 1:
        dist.all_gather_object(verdicts, verdict)
    else:
        verdicts = [verdict]
    out["nan_column"] = {"verdicts": verdicts, "all_value_errors": all(v.startswith("ValueError") for v in verdicts)}

    ok = (out["ksg_n1e6"]["bitwise_equal"] and out["cmi_n1e5"]["abs_diff"] <= 1e-10 and out["entropy4d_n1e5"]["abs_diff"] <= 1e-10
          and out["pairwise_16x5e4"]["bitwise_equal"] and out["lag_sweep"]["bitwise_equal"] and out["nan_column"]["all_value_errors"])
    out["ok"] = bool(ok)
    if world > 1:
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        out["ok_all_ranks"] = all(flags)
        dist.barrier()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
