"""Where the time of pairwise_mi (64 x 1e5) goes: host profile + device phases of the native call."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ENNEMI_B200_DEVICES", "0")
import numpy as np
import ennemi_b200 as eb
from ennemi_b200 import _native as nat
data = np.random.default_rng(0).normal(size=(100_000, 64))
eb.pairwise_mi(data[:, :8])
for _ in range(2):
    t0 = time.perf_counter(); eb.pairwise_mi(data); print("wall s", time.perf_counter() - t0, nat.last_timing())
pr = cProfile.Profile(); pr.enable(); eb.pairwise_mi(data); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:4000])
