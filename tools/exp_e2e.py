import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ennemi_b200 as eb
from ennemi_b200 import _native as nat, _columns, _align
rng = np.random.default_rng(0)
d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=1_000_000)
y, x = np.ascontiguousarray(d[:, 1]), np.ascontiguousarray(d[:, 0])
def t(f, n=7):
    f(); f(); b = 1e9
    for _ in range(n):
        t0 = time.perf_counter(); f(); b = min(b, time.perf_counter() - t0)
    return b * 1e3
print("estimate_mi            %.3f ms" % t(lambda: eb.estimate_mi(y, x, k=3)))
print("cache_put 8 MB         %.3f ms" % t(lambda: nat.cache_put(991, x)))
print("cache_stats            %.3f ms" % t(lambda: nat.cache_stats(991, 0, 1_000_000)))
nat.cache_put(992, y)
mx, sx = nat.cache_stats(991, 0, 1_000_000); my, sy = nat.cache_stats(992, 0, 1_000_000)
st = _align._NoiseStream(); nx_ = st.normal((1_000_000,)); ny_ = st.normal((1_000_000,))
nat.cache_put(993, nx_); nat.cache_put(994, ny_)
descs = [nat.ColDesc(991, 0, 1, mx, sx, 993, 0, 1), nat.ColDesc(992, 0, 1, my, sy, 994, 0, 1)]
print("ksg_mi_cols (cached)   %.3f ms" % t(lambda: nat.ksg_mi_cols(descs, 1_000_000, 3)), nat.last_timing())
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): eb.estimate_mi(y, x, k=3)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
