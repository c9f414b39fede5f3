"""Small inputs through every kernel family (run under compute-sanitizer: memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat

rng = np.random.default_rng(0)
n = 20_000
d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=n)
co = nat.pack_coords([d[:, 0], d[:, 1]])
print("ksg k=3", nat.ksg_mi(co, 3, details=True)[0], nat.last_pipeline())
print("ksg k=6", nat.ksg_mi(co, 6), nat.last_pipeline())
print("ksg shard", nat.ksg_mi_finish(nat.ksg_mi_rows(co.ctypes.data, n, 3, 0, n // 2) + nat.ksg_mi_rows(co.ctypes.data, n, 3, n // 2, n), n, 3))
t = rng.standard_t(2, size=(n, 2))
print("heavy tails", nat.ksg_mi(nat.pack_coords([t[:, 0], t[:, 1]]), 3))
keys = [7001, 7002, 7003, 7004]
for j, key in enumerate(keys):
    nat.cache_put(key, np.ascontiguousarray(rng.normal(size=n)))
cols = [nat.ColDesc(key, 0, 1, 0.0, 0.0, 0, 0, 1) for key in keys]
pairs = np.array([(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)], dtype=np.int32)
print("pairs", nat.ksg_mi_pairs(cols, pairs, n, 3))
for key in keys:
    nat.cache_drop(key)
print("general path k=9", nat.ksg_mi(co, 9), nat.last_pipeline())
z = rng.normal(size=(n, 2))
print("cmi", nat.cmi(nat.pack_coords([d[:, 0], d[:, 1], z]), 3))
print("entropy", nat.entropy(nat.pack_coords([z]), 3))
cls = rng.integers(0, 5, n).astype(np.int32)
print("ross", nat.ross_mi(nat.pack_coords([d[:, 0]]), cls, 5, 3))
print("ross cmi", nat.ross_cmi(nat.pack_coords([d[:, 0], z]), cls, 5, 3))
print("entropy 1-d (two-pointer k-NN)", nat.entropy(nat.pack_coords([d[:, 0]]), 3))
# three-level grid (forced onto a small input), entropy and the opt-in Frenzel-Pompe variant
os.environ["EB2_G3_MIN"] = "2"; os.environ["EB2_G3_CMI"] = "1"
x4 = rng.standard_t(3, size=(n, 4))
print("grid entropy 4-d", nat.entropy(nat.pack_coords([x4]), 5, details=True)[0], nat.last_pipeline())
print("grid entropy 3-d", nat.entropy(nat.pack_coords([x4[:, :3]]), 3), nat.last_pipeline())
z3 = rng.normal(size=(n, 3))
print("grid cmi c=3", nat.cmi(nat.pack_coords([d[:, 0], d[:, 1], z3]), 3, details=True)[0], nat.last_pipeline())
del os.environ["EB2_G3_MIN"]; del os.environ["EB2_G3_CMI"]
# a full deferral list in the bivariate pipeline
os.environ["EB2_K2_LEFTCAP"] = "40"
c2 = rng.standard_cauchy(size=(6_000, 2))
print("full deferral list", nat.ksg_mi(nat.pack_coords([c2[:, 0], c2[:, 0] + c2[:, 1]]), 3, details=True)[0], nat.last_pipeline())
del os.environ["EB2_K2_LEFTCAP"]
