import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat
rng = np.random.default_rng(0)
d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=1_000_000)
co = nat.pack_coords([d[:, 0], d[:, 1]])
ref = None
for thr in [int(a) for a in sys.argv[1:]] or (0, 4, 16, 64):
    os.environ["EB2_DEFER"] = str(thr)
    v, parts = nat.ksg_mi(co, 3, details=True)
    if ref is None: ref = parts
    ok = all(np.array_equal(parts[k], ref[k]) for k in parts)
    part = nat.ksg_mi_rows(co.ctypes.data, 1_000_000, 3, 0, 1_000_000)
    t = nat.last_timing()
    print("defer", thr, "knn_ms %.3f total_ms %.3f" % (t["knn_ms"], t["total_ms"]), "pairs %.3e" % part[nat.P_PAIRS], "same_as_first", ok, v)
