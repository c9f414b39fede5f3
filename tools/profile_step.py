"""One resident KSG step per mode at N = 1e6 (profiling target: run under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat
mode = sys.argv[1] if len(sys.argv) > 1 else "pruned"
rng = np.random.default_rng(0)
if mode in ("pruned", "brute"):
    d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=1_000_000)
    co = nat.pack_coords([d[:, 1], d[:, 0]])
    flags = nat.FLAG_NO_PRUNE if mode == "brute" else 0
    for _ in range(3):
        print(nat.ksg_mi(co, 3, flags=flags), nat.last_timing())
else:   # cmi: configs[2] one lag
    N = 200_000
    z = rng.normal(size=(N, 3)); x = rng.normal(size=N) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=N)
    co = nat.pack_coords([x, y, z])
    for _ in range(3):
        print(nat.cmi(co, 3), nat.last_timing())
