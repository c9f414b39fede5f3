"""Developer timing of the bivariate pipeline on heavy-tailed inputs (N = 1e6, resident)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ennemi_b200 import _native as nat
rng = np.random.default_rng(0)
n = 1_000_000
cases = {"gauss": rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=n), "student-t3": rng.standard_t(3, size=(n, 2)),
         "student-t2": rng.standard_t(2, size=(n, 2)), "cauchy": rng.standard_cauchy(size=(n, 2)),
         "lognormal": rng.lognormal(size=(n, 2)), "outliers": np.concatenate([rng.normal(size=(n - 50, 2)), rng.normal(size=(50, 2)) * 1e4])}
for name, d in cases.items():
    co = torch.from_numpy(nat.pack_coords([np.ascontiguousarray(d[:, 0]), np.ascontiguousarray(d[:, 1])])).cuda()
    best = None
    for _ in range(6):
        part = nat.ksg_mi_rows(int(co.data_ptr()), n, 3, 0, n, flags=nat.FLAG_DEVICE_INPUT)
        t = nat.last_timing()
        if best is None or t["total_ms"] < best["total_ms"]:
            best = t
    print(name, "pipeline", nat.last_pipeline(), {k: round(v, 3) for k, v in best.items()}, "pairs/row", round(part[nat.P_PAIRS] / n, 1))
