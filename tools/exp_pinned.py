"""e2e estimate_mi with pageable vs pinned host arrays (the API takes NumPy arrays either way)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ennemi_b200 as eb
from ennemi_b200 import _native as nat, _columns
rng = np.random.default_rng(0)
d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=1_000_000)
y, x = np.ascontiguousarray(d[:, 1]), np.ascontiguousarray(d[:, 0])
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    v = t.numpy(); v[...] = a
    return v, t
yp, _ty = pinned(y); xp, _tx = pinned(x)
def t(f, n=9):
    f(); f(); f(); b = 1e9
    for _ in range(n):
        t0 = time.perf_counter(); f(); b = min(b, time.perf_counter() - t0)
    return b * 1e3
print("estimate_mi pageable   %.3f ms" % t(lambda: eb.estimate_mi(y, x, k=3)))
print("estimate_mi pinned     %.3f ms" % t(lambda: eb.estimate_mi(yp, xp, k=3)), nat.last_timing())
print("cache_put pageable     %.3f ms" % t(lambda: nat.cache_put(991, x)))
print("cache_put pinned       %.3f ms" % t(lambda: nat.cache_put(991, xp)))
print("cache_stats            %.3f ms" % t(lambda: nat.cache_stats(991, 0, 1_000_000)))
nat.cache_put(992, yp)
from ennemi_b200 import _align
st = _align._NoiseStream(); nx_ = st.normal((1_000_000,)); ny_ = st.normal((1_000_000,))
nat.cache_put(993, nx_); nat.cache_put(994, ny_)
nan = float("nan")
descs = [nat.ColDesc(991, 0, 1, nan, 1.0, 993, 0, 1), nat.ColDesc(992, 0, 1, nan, 1.0, 994, 0, 1)]
F = nat.FLAG_SINGLE_USE | nat.FLAG_DEVICE_STATS
print("ksg_mi_cols in-call    %.3f ms" % t(lambda: nat.ksg_mi_cols(descs, 1_000_000, 3, flags=F)), nat.last_timing())
mx, sx = nat.cache_stats(991, 0, 1_000_000); my, sy = nat.cache_stats(992, 0, 1_000_000)
descs2 = [nat.ColDesc(991, 0, 1, mx, sx, 993, 0, 1), nat.ColDesc(992, 0, 1, my, sy, 994, 0, 1)]
print("ksg_mi_cols host stats %.3f ms" % t(lambda: nat.ksg_mi_cols(descs2, 1_000_000, 3, flags=nat.FLAG_SINGLE_USE)), nat.last_timing())
def both():
    nat.cache_put(991, xp); nat.cache_put(992, yp); return nat.ksg_mi_cols(descs, 1_000_000, 3, flags=F)
print("2 puts + in-call       %.3f ms" % t(both))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): eb.estimate_mi(yp, xp, k=3)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(16)
