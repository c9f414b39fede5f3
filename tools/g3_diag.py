"""Developer diagnostics of the three-level grid: candidates examined per query (work counter of the partial block)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat
rng = np.random.default_rng(0)
N = 500_000
cov = np.array([[1.0, 0.5, 0.2, 0.1], [0.5, 1.0, 0.3, 0.0], [0.2, 0.3, 1.0, -0.4], [0.1, 0.0, -0.4, 1.0]])
d4 = nat.pack_coords([rng.multivariate_normal(np.zeros(4), cov, size=N)])
for _ in range(2):
    p = nat.entropy_rows(d4.ctypes.data, N, 4, 5, 0, N)
print("entropy4d pairs/query", p[nat.P_PAIRS] / N, nat.last_timing(), nat.last_pipeline())
v, d = nat.entropy(d4, 5, details=True)
print("eps quantiles", np.quantile(d["dist"], [0.01, 0.5, 0.99, 1.0]))
N = 200_000
z = rng.normal(size=(N, 3)); x = rng.normal(size=N) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=N)
co = nat.pack_coords([x, y, z])
for _ in range(2):
    p = nat.cmi_rows(co.ctypes.data, N, 3, 3, 0, N)
print("cmi c=3 pairs/query", p[nat.P_PAIRS] / N, nat.last_timing(), nat.last_pipeline())
