"""Developer timing of the five BASELINE.json configs (single GPU), native + public API."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat
import ennemi_b200 as eb

def t(fn, reps=3):
    fn(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); best = min(best, time.perf_counter() - t0)
    return best, r

rng = np.random.default_rng(0)
d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=10_000)
print("cfg1 estimate_mi N=1e4: %.2f ms" % (1e3 * t(lambda: eb.estimate_mi(d[:, 1], d[:, 0], k=3))[0]))
d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=1_000_000)
print("cfg2 estimate_mi N=1e6: %.2f ms" % (1e3 * t(lambda: eb.estimate_mi(d[:, 1], d[:, 0], k=3))[0]), nat.last_timing())
N = 200_000
z = rng.normal(size=(N, 3)); x = rng.normal(size=N) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=N)
co = nat.pack_coords([x, y, z])
print("cfg3 cmi one lag native: %.2f ms" % (1e3 * t(lambda: nat.cmi(co, 3))[0]), nat.last_timing())
part = nat.cmi_rows(co.ctypes.data, N, 3, 3, 0, N); print("   pairs %.3e" % part[nat.P_PAIRS])
print("cfg3 cmi brute one lag native: %.2f ms" % (1e3 * t(lambda: nat.cmi(co, 3, flags=nat.FLAG_NO_PRUNE), 1)[0]), nat.last_timing())
tt, r = t(lambda: eb.estimate_mi(y, x, lag=range(8), k=3, cond=z), 1)
print("cfg3 estimate_mi 8 lags: %.1f ms -> 50 lags ~ %.1f ms" % (1e3 * tt, 1e3 * tt * 50 / 8))
data = rng.normal(size=(100_000, 16))
tt, r = t(lambda: eb.pairwise_mi(data, k=3), 1)
print("cfg4 pairwise 16 vars (120 pairs): %.1f ms -> 2016 pairs ~ %.1f ms" % (1e3 * tt, 1e3 * tt * 2016 / 120))
N = 500_000
yd = rng.integers(0, 16, N); xc = rng.normal(size=N) + 0.25 * yd
print("cfg5a ross N=5e5: %.2f ms" % (1e3 * t(lambda: eb.estimate_mi(xc, yd, discrete_x=True, k=5))[0]), nat.last_timing())
cov = np.array([[1.0, 0.5, 0.6, -0.2], [0.5, 1.0, 0.7, -0.5], [0.6, 0.7, 2.0, -0.1], [-0.2, -0.5, -0.1, 0.5]])
x4 = rng.multivariate_normal([0, 0, 0, 0], cov, size=N)
print("cfg5b entropy 4-D N=5e5: %.2f ms" % (1e3 * t(lambda: eb.estimate_entropy(x4, k=5, multidim=True))[0]), nat.last_timing())
co4 = nat.pack_coords([x4]); part = nat.entropy_rows(co4.ctypes.data, N, 4, 5, 0, N); print("   pairs %.3e" % part[nat.P_PAIRS])
