"""Two 4-D k-NN entropy estimates at N = 5e5, k = 5 through the three-level grid (profiling target: run under ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat
rng = np.random.default_rng(0)
N = 500_000
cov = np.array([[1.0, 0.5, 0.2, 0.1], [0.5, 1.0, 0.3, 0.0], [0.2, 0.3, 1.0, -0.4], [0.1, 0.0, -0.4, 1.0]])
d4 = nat.pack_coords([rng.standard_t(2, size=(N, 4)) if "t2" in sys.argv else rng.multivariate_normal(np.zeros(4), cov, size=N)])
for _ in range(2):
    p = nat.entropy_rows(d4.ctypes.data, N, 4, 5, 0, N)
print("pairs/query", p[nat.P_PAIRS] / N, nat.last_timing(), nat.last_pipeline())
