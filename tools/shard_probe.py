"""Where the time of a row-sharded estimate goes (torchrun, 2+ ranks): library call vs exchange."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from ennemi_b200 import _native as nat, distributed as ebd
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 1_000_000
d = np.random.default_rng(0).multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=n)
co = torch.from_numpy(nat.pack_coords([d[:, 0], d[:, 1]])).to(dev)
lo, hi = ebd.shard_bounds(n, rank, world)
for _ in range(6):
    part = nat.ksg_mi_rows(int(co.data_ptr()), n, 3, lo, hi, dev=local, flags=nat.FLAG_DEVICE_INPUT); ebd._all_reduce_sum(part)
for label, full in (("shard", False), ("whole", True)):
    a, b = (0, n) if full else (lo, hi)
    for _ in range(4):
        nat.ksg_mi_rows(int(co.data_ptr()), n, 3, a, b, dev=local, flags=nat.FLAG_DEVICE_INPUT)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        part = nat.ksg_mi_rows(int(co.data_ptr()), n, 3, a, b, dev=local, flags=nat.FLAG_DEVICE_INPUT)
    t1 = time.perf_counter()
    print(rank, label, "library call ms", (t1 - t0) / 20 * 1e3, nat.last_timing(local))
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    dist.barrier()
t1 = time.perf_counter()
print(rank, "barrier ms", (t1 - t0) / 20 * 1e3)
t0 = time.perf_counter()
for _ in range(20):
    part = nat.ksg_mi_rows(int(co.data_ptr()), n, 3, lo, hi, dev=local, flags=nat.FLAG_DEVICE_INPUT)
    tot = ebd._all_reduce_sum(part)
t1 = time.perf_counter()
print(rank, "call + exchange ms", (t1 - t0) / 20 * 1e3, nat.ksg_mi_finish(tot, n, 3))
dist.destroy_process_group()
