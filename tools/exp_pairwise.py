"""pairwise_mi 64 x 1e5 (BASELINE configs[3]): wall time through the public API and where it goes —
block upload, whole-column statistics, and the 2,016 pair tasks driven (a) by one native batch call on one
stream lane, (b) by three threads on three lanes without any Python per task, (c) by the API's scheduler."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ennemi_b200 as eb
from ennemi_b200 import _native as nat, _columns, _devices, _align

data = np.random.default_rng(0).normal(size=(100_000, 64))
n, nvar = data.shape
eb.pairwise_mi(data[:, :8])
for rep in range(3):
    t0 = time.perf_counter(); pw = eb.pairwise_mi(data); t1 = time.perf_counter()
    print("pairwise_mi 64 vars: %.4f s  (max off-diagonal %.6f)" % (t1 - t0, float(np.nanmax(pw))))

keys = list(range(990_000, 990_000 + nvar))
t0 = time.perf_counter(); nat.cache_put_block(keys, data); t1 = time.perf_counter()
print("block upload + de-interleave (51 MB pageable): %.2f ms" % ((t1 - t0) * 1e3))
t0 = time.perf_counter(); means, stds = nat.cache_stats_many(keys, [0] * nvar, n); t1 = time.perf_counter()
print("64 whole-column mean/std: %.2f ms" % ((t1 - t0) * 1e3))
assert all(means[j] == data[:, j].mean() and stds[j] == data[:, j].std() for j in range(nvar))
stream = _align._NoiseStream()
nx = stream.normal((n,)); ny = stream.normal((n,))
nat.cache_put(989_001, nx); nat.cache_put(989_002, ny)
pairs = [(i, j) for i in range(nvar) for j in range(i + 1, nvar)]
tasks = [[nat.ColDesc(keys[i], 0, 1, means[i], stds[i], 989_001, 0, 1), nat.ColDesc(keys[j], 0, 1, means[j], stds[j], 989_002, 0, 1)]
         for i, j in pairs]
nat.mi_cols_batch(tasks[:64], n, 3)
t0 = time.perf_counter(); vals, st = nat.mi_cols_batch(tasks, n, 3); t1 = time.perf_counter()
print("2016 pairs, ONE native call, one lane: %.4f s (%.1f us/pair)" % (t1 - t0, (t1 - t0) / len(pairs) * 1e6))
assert not st.any() and np.array_equal(vals, np.array([pw[i, j] for i, j in pairs]))
for lanes in (2, 3, 4):
    out = [None] * lanes
    def work(l):
        out[l] = nat.mi_cols_batch(tasks[l::lanes], n, 3, dev=_devices.with_lane(0, l))
    th = [threading.Thread(target=work, args=(l,)) for l in range(lanes)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    t1 = time.perf_counter()
    print("2016 pairs, %d native calls on %d lanes: %.4f s" % (lanes, lanes, t1 - t0))
one = nat.last_timing(0)
print("last task on lane 0:", one)
for key in keys + [989_001, 989_002]:
    nat.cache_drop(key)
