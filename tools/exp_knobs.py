"""k-NN phase time of bivariate KSG at N=1e5 / 1e6 under tuning knobs given as KEY=v1,v2,... arguments
(cartesian product), e.g. `python tools/exp_knobs.py EB2_DEFER=0,4,8,16 EB2_LANE_SCAN=0,160`."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat
knobs = [a.split("=") for a in sys.argv[1:]]
names = [k for k, _ in knobs]
rng = np.random.default_rng(0)
sets = {}
for N in (100_000, 1_000_000):
    d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=N)
    sets[N] = nat.pack_coords([d[:, 0], d[:, 1]])
ref = {}
for combo in itertools.product(*[v.split(",") for _, v in knobs]):
    for k, v in zip(names, combo):
        os.environ[k] = v
    out = []
    for N, co in sets.items():
        v, det = nat.ksg_mi(co, 3, details=True)
        if N not in ref:
            ref[N] = det["eps"]
        ok = np.array_equal(ref[N], det["eps"])
        best = None
        for _ in range(6):
            nat.ksg_mi(co, 3)
            t = nat.last_timing()
            if best is None or t["knn_ms"] < best["knn_ms"]:
                best = t
        out.append("N=%d knn %.3f total %.3f %s" % (N, best["knn_ms"], best["total_ms"], "" if ok else "EPS-MISMATCH"))
    print(" ".join("%s=%s" % kv for kv in zip(names, combo)), "|", " | ".join(out), flush=True)
