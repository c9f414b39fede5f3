import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ennemi_b200 as eb
data = np.random.default_rng(0).normal(size=(100_000, 64))
eb.pairwise_mi(data[:, :8])
for rep in range(3):
    t0 = time.perf_counter(); pw = eb.pairwise_mi(data); t1 = time.perf_counter()
    print("pairwise 64:", t1 - t0, float(np.nanmax(pw)))
