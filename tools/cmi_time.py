"""Developer timing of one Frenzel-Pompe estimate (N = 2e5, 3-D condition: BASELINE.json configs[2]) and of the 50-lag sweep."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat
import ennemi_b200 as eb
rng = np.random.default_rng(0)
N = 200_000
z = rng.normal(size=(N, 3)); x = rng.normal(size=N) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=N)
co = nat.pack_coords([x, y, z])
best = None
for _ in range(4):
    v = nat.cmi(co, 3)
    t = nat.last_timing()
    if best is None or t["total_ms"] < best["total_ms"]:
        best = t
print("one lag:", v, {k: round(val, 3) for k, val in best.items()})
if "--sweep" in sys.argv:
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); eb.estimate_mi(y, x, lag=list(range(50)), k=3, cond=z); ts.append(time.perf_counter() - t0)
    print("50-lag sweep:", round(min(ts), 4), "s")
