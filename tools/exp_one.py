"""One KSG estimate at N (argv[1], default 1e6) for profiling: `ncu -k regex:knn ... python tools/exp_one.py 1000000`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = np.random.default_rng(0).multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=N)
co = nat.pack_coords([d[:, 0], d[:, 1]])
for _ in range(reps):
    v = nat.ksg_mi(co, 3)
print(v, nat.last_timing())
