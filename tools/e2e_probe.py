"""Developer probe: where the time of one estimate_mi(y, x, k=3) call on page-locked host arrays goes (N = 1e6)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ennemi_b200 as eb
from ennemi_b200 import _native as nat, _columns

rng = np.random.default_rng(0)
n = 1_000_000
d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=n)
keep = []
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True); t.numpy()[...] = a; keep.append(t); return t.numpy()
y, x = pinned(d[:, 1]), pinned(d[:, 0])

def timed(f, reps=30):
    for _ in range(5): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3

print("estimate_mi, pinned:", round(timed(lambda: eb.estimate_mi(y, x, k=3)), 4), "ms", nat.last_timing())
# uploads alone, one after the other and from two threads
k1, k2 = 777001, 777002
def up_seq():
    nat.cache_put(k1, x); nat.cache_put(k2, y)
print("two uploads, sequential, one lane:", round(timed(up_seq), 4), "ms")
from concurrent.futures import ThreadPoolExecutor
pool = ThreadPoolExecutor(2)
def up_par():
    f = pool.submit(nat.cache_put, k2, y, 0 | (1 << 8)); nat.cache_put(k1, x); f.result()
print("two uploads, two threads / lanes:", round(timed(up_par), 4), "ms")
cols = [nat.ColDesc(k1, 0, 1, float("nan"), 1.0, 0, 0, 1), nat.ColDesc(k2, 0, 1, float("nan"), 1.0, 0, 0, 1)]
print("ksg_mi_cols on resident columns (device stats + rescale + pipeline):", round(timed(lambda: nat.ksg_mi_cols(cols, n, 3, flags=nat.FLAG_DEVICE_STATS)), 4), "ms", nat.last_timing())
cols0 = [nat.ColDesc(k1, 0, 1, 0.0, 0.0, 0, 0, 1), nat.ColDesc(k2, 0, 1, 0.0, 0.0, 0, 0, 1)]
print("ksg_mi_cols, pass-through columns:", round(timed(lambda: nat.ksg_mi_cols(cols0, n, 3)), 4), "ms", nat.last_timing())
nat.cache_drop(k1); nat.cache_drop(k2)
