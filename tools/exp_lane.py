"""A/B of the per-lane window scan (EB2_LANE_SCAN) in the two-level k-NN kernel: bit-equality of eps and k-NN time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat

rng = np.random.default_rng(0)
cases = []
for N in (100_000, 1_000_000):
    d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=N)
    cases.append(("ksg N=%d k=3" % N, "ksg", nat.pack_coords([d[:, 0], d[:, 1]]), 3))
d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=300_000)
cases.append(("ksg N=3e5 k=6", "ksg", nat.pack_coords([d[:, 0], d[:, 1]]), 6))
t = rng.standard_t(2, size=(200_000, 2))
cases.append(("ksg heavy tails N=2e5", "ksg", nat.pack_coords([t[:, 0], t[:, 1]]), 3))
N = 200_000
z = rng.normal(size=(N, 3)); x = rng.normal(size=N) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=N)
cases.append(("cmi N=2e5 c=3", "cmi", nat.pack_coords([x, y, z]), 3))
cov = np.array([[1.0, 0.5, 0.6, -0.2], [0.5, 1.0, 0.7, -0.5], [0.6, 0.7, 2.0, -0.1], [-0.2, -0.5, -0.1, 0.5]])
x4 = rng.multivariate_normal([0, 0, 0, 0], cov, size=500_000)
cases.append(("entropy 4-D N=5e5 k=5", "ent", nat.pack_coords([x4]), 5))
x7 = rng.normal(size=(100_000, 7))
cases.append(("entropy 7-D N=1e5 k=3", "ent", nat.pack_coords([x7]), 3))

def run(kind, co, k):
    if kind == "ksg":
        v, d = nat.ksg_mi(co, k, details=True)
    elif kind == "cmi":
        v, d = nat.cmi(co, k, details=True)
    else:
        v, d = nat.entropy(co, k, details=True)
    return v, d

def timed(kind, co, k):
    fn = {"ksg": nat.ksg_mi, "cmi": nat.cmi, "ent": nat.entropy}[kind]
    fn(co, k)
    best = None
    for _ in range(5):
        fn(co, k)
        t = nat.last_timing()
        if best is None or t["knn_ms"] < best["knn_ms"]:
            best = t
    return best

x3 = rng.normal(size=(300_000, 3)) @ np.array([[1, .5, 0], [0, 1, .3], [0, 0, 1]])
cases.append(("entropy 3-D N=3e5 k=3", "ent", nat.pack_coords([x3]), 3))
modes = sys.argv[1:] or ["0", "48", "96", "160"]
for name, kind, co, k in cases:
    res = {}
    for mode in modes:
        os.environ["EB2_LANE_SCAN"] = mode
        v, d = run(kind, co, k)
        res[mode] = (v, d, timed(kind, co, k))
    v0, d0, t0 = res[modes[0]]
    same = all(all(np.array_equal(d0[kk], res[m][1][kk]) for kk in d0) and res[m][0] == v0 for m in modes)
    print("%-26s same=%s value %.12g | knn_ms " % (name, same, v0) + "  ".join("%s: %.3f" % (m, res[m][2]["knn_ms"]) for m in modes)
          + " | total " + "  ".join("%.3f" % res[m][2]["total_ms"] for m in modes))
