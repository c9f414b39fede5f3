"""Developer check of the three-level grid (entropy / Frenzel-Pompe in >= 3 dimensions) against the brute-force kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
os.environ.setdefault("EB2_G3_CMI", "1")
from ennemi_b200 import _native as nat
BRUTE = nat.FLAG_NO_PRUNE | nat.FLAG_BRUTE_COUNT
rng = np.random.default_rng(0)
bad = 0

def ent(name, x, k):
    global bad
    co = nat.pack_coords([x])
    v, d = nat.entropy(co, k, details=True)
    pipe = nat.last_pipeline()
    vb, b = nat.entropy(co, k, flags=BRUTE, details=True)
    ok = np.array_equal(d["dist"], b["dist"]) and (v == vb or abs(v - vb) <= 1e-10 or (np.isinf(v) and v == vb))
    print("ok " if ok else "MISMATCH ", "entropy", name, x.shape, "k", k, v, int(np.sum(d["dist"] != b["dist"])), "pipe", pipe)
    bad += 0 if ok else 1

def cmi(name, x, y, z, k):
    global bad
    co = nat.pack_coords([x, y, z])
    v, d = nat.cmi(co, k, details=True)
    vb, b = nat.cmi(co, k, flags=BRUTE, details=True)
    mism = {key: int(np.sum(d[key] != b[key])) for key in ("eps", "nxz", "nyz", "nz")}
    ok = not any(mism.values()) and (v == vb or abs(v - vb) <= 1e-10 or (np.isnan(v) and np.isnan(vb)) or v == vb)
    print("ok " if ok else "MISMATCH ", "cmi", name, z.shape, "k", k, v, mism)
    bad += 0 if ok else 1

n = 20_000
for D in (3, 4, 5, 7):
    for k in (1, 3, 6):
        ent("gauss", rng.normal(size=(n, D)) @ rng.normal(size=(D, D)), k)
ent("t2", rng.standard_t(2, size=(30_000, 4)), 3)
ent("ties", np.round(rng.normal(size=(20_000, 3)), 1), 3)
ent("dups", np.repeat(rng.normal(size=(5_000, 4)), 4, axis=0), 3)
for c in (2, 3, 4, 6):
    z = rng.normal(size=(n, c)); x = rng.normal(size=n) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=n)
    for k in (1, 3, 5):
        cmi("gauss", x, y, z, k)
z = rng.standard_t(2, size=(30_000, 3)); x = rng.normal(size=30_000) + z[:, 0]; y = rng.normal(size=30_000)
cmi("t2", x, y, z, 3)
z = np.round(rng.normal(size=(20_000, 2)), 1); x = rng.normal(size=20_000); y = x + z[:, 0]
cmi("ties", x, y, z, 3)
print("g3_check: bad =", bad)
if "--no-timing" not in sys.argv:
    N = 500_000
    cov = np.array([[1.0, 0.5, 0.2, 0.1], [0.5, 1.0, 0.3, 0.0], [0.2, 0.3, 1.0, -0.4], [0.1, 0.0, -0.4, 1.0]])
    d4 = nat.pack_coords([rng.multivariate_normal(np.zeros(4), cov, size=N)])
    for env in ({}, {"EB2_NO_G3": "1"}):
        os.environ.pop("EB2_NO_G3", None); os.environ.update(env)
        for _ in range(3): v = nat.entropy(d4, 5)
        print("entropy4d", env, v, nat.last_timing(), nat.last_pipeline())
    N = 200_000
    z = rng.normal(size=(N, 3)); x = rng.normal(size=N) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=N)
    co = nat.pack_coords([x, y, z])
    for env in ({}, {"EB2_NO_G3": "1"}):
        os.environ.pop("EB2_NO_G3", None); os.environ.update(env)
        for _ in range(3): v = nat.cmi(co, 3)
        print("cmi c=3", env, v, nat.last_timing(), nat.last_pipeline())
    os.environ.pop("EB2_NO_G3", None)
sys.exit(1 if bad else 0)
