"""One resident bivariate-pipeline step at N = 1e6 without graph replay (profiling target: run under ncu)."""
import os, sys
os.environ["EB2_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ennemi_b200 import _native as nat
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(0)
d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=n)
co = torch.from_numpy(nat.pack_coords([d[:, 0], d[:, 1]])).cuda()
for _ in range(3):
    part = nat.ksg_mi_rows(int(co.data_ptr()), n, 3, 0, n, flags=nat.FLAG_DEVICE_INPUT)
print(nat.ksg_mi_finish(part, n, 3), nat.last_timing())
