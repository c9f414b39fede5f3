"""Phase times of single shards of the N = 1e6 estimate on ONE GPU (which part of the row range costs what)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ennemi_b200 import _native as nat
n = 1_000_000
d = np.random.default_rng(0).multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=n)
co = torch.from_numpy(nat.pack_coords([d[:, 0], d[:, 1]])).cuda()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for g in list(range(G)) + [-1]:
    lo, hi = (0, n) if g < 0 else (n * g // G, n * (g + 1) // G)
    for _ in range(5):
        part = nat.ksg_mi_rows(int(co.data_ptr()), n, 3, lo, hi, flags=nat.FLAG_DEVICE_INPUT)
    t = nat.last_timing()
    print("shard", g, "rows", int(part[nat.P_ROWS]), "knn", round(t["knn_ms"], 4), "count", round(t["count_ms"], 4), "layout", round(t["layout_ms"], 4),
          "total", round(t["total_ms"], 4), "pairs/row", round(part[nat.P_PAIRS] / max(part[nat.P_ROWS], 1), 1))
