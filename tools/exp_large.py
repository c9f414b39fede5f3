"""One-off robustness check at N = 10^7: pruned vs brute-force k-NN distances and counts must agree bit for bit."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rng = np.random.default_rng(0)
d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=n)
co = nat.pack_coords([d[:, 0], d[:, 1]])
t0 = time.perf_counter(); v, p = nat.ksg_mi(co, 3, details=True); t1 = time.perf_counter()
print("pruned: mi", v, "wall %.1f ms" % ((t1 - t0) * 1e3), nat.last_timing())
if "--brute" in sys.argv:
    t0 = time.perf_counter(); vb, pb = nat.ksg_mi(co, 3, flags=nat.FLAG_NO_PRUNE, details=True); t1 = time.perf_counter()
    print("brute: mi", vb, "wall %.1f s" % (t1 - t0), "eps equal", np.array_equal(p["eps"], pb["eps"]),
          "nx equal", np.array_equal(p["nx"], pb["nx"]), "ny equal", np.array_equal(p["ny"], pb["ny"]), "dMI", abs(v - vb))
