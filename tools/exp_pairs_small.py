"""A few pairwise_mi-style pair tasks (N=1e5, cached columns, prepared variables) for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat, _align
n, nvar = 100_000, 4
data = np.random.default_rng(0).normal(size=(n, nvar))
keys = list(range(990_000, 990_000 + nvar))
nat.cache_put_block(keys, data)
means, stds = nat.cache_stats_many(keys, [0] * nvar, n)
stream = _align._NoiseStream()
nat.cache_put(989_001, stream.normal((n,))); nat.cache_put(989_002, stream.normal((n,)))
pairs = [(i, j) for i in range(nvar) for j in range(i + 1, nvar)]
tasks = [[nat.ColDesc(keys[i], 0, 1, means[i], stds[i], 989_001, 0, 1), nat.ColDesc(keys[j], 0, 1, means[j], stds[j], 989_002, 0, 1)]
         for i, j in pairs]
for _ in range(2):
    vals, st = nat.mi_cols_batch(tasks, n, 3)
print(vals, nat.last_timing())
