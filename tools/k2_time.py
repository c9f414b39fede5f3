"""Developer timing of the bivariate pipeline at N = 1e6 and 1e5 with resident inputs (run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ennemi_b200 import _native as nat
if os.environ.get("EB2_LIB"):
    nat.LIB_PATH = os.path.abspath(os.environ["EB2_LIB"])          # a build variant (developer experiments)

rng = np.random.default_rng(0)
for n in (1_000_000, 100_000):
    d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=n)
    co = torch.from_numpy(nat.pack_coords([d[:, 0], d[:, 1]])).cuda()
    for env in [""]:
        for _ in range(5):
            part = nat.ksg_mi_rows(int(co.data_ptr()), n, 3, 0, n, flags=nat.FLAG_DEVICE_INPUT)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            part = nat.ksg_mi_rows(int(co.data_ptr()), n, 3, 0, n, flags=nat.FLAG_DEVICE_INPUT)
        t1 = time.perf_counter()
        print("N", n, env, "mi", nat.ksg_mi_finish(part, n, 3), "ms/step", round((t1 - t0) / reps * 1e3, 4), nat.last_timing(),
              "pairs/row", part[nat.P_PAIRS] / n)
