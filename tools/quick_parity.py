"""Developer check: every golden estimator case through the C ABI in every mode (run under gpurun)."""
import sys, os, warnings, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat

g = np.load("tests/golden/estimators.npz")
names = sorted({k.split("/")[0] for k in g.files})
def case(n): return {k.split("/")[1]: g[k] for k in g.files if k.startswith(n + "/")}
def same(a, b):
    return (a == b) or (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-10
bad = 0
for flags in (0, nat.FLAG_NO_PRUNE, nat.FLAG_BRUTE_COUNT, nat.FLAG_NO_PRUNE | nat.FLAG_BRUTE_COUNT):
    for n in names:
        c = case(n)
        if n == "psi":
            out = nat.psi(c["n"])
            err = np.max(np.abs(out - c["value"]))
            if err > 1e-13: bad += 1; print("psi err", err)
            continue
        k = int(c["k"])
        try:
            if n.startswith("ksg"):
                v, d = nat.ksg_mi(nat.pack_coords([c["x"], c["y"]]), k, flags=flags, details=True)
            elif n.startswith("cmi"):
                v, d = nat.cmi(nat.pack_coords([c["x"], c["y"], c["z"]]), k, flags=flags, details=True)
            elif n.startswith("ross"):
                labels, inv = np.unique(c["y"], return_inverse=True)
                v, d = nat.ross_mi(nat.pack_coords([c["x"]]), inv.astype(np.int32), len(labels), k, flags=flags, details=True)
            elif n.startswith("cross"):
                labels, inv = np.unique(c["y"], return_inverse=True)
                v, d = nat.ross_cmi(nat.pack_coords([c["x"], c["z"]]), inv.astype(np.int32), len(labels), k, flags=flags, details=True)
            elif n.startswith("ent"):
                v, d = nat.entropy(nat.pack_coords([c["x"]]), k, flags=flags, details=True)
        except Exception as e:
            bad += 1; print("EXC", n, flags, type(e).__name__, e); continue
        msgs = []
        for key, arr in d.items():
            if not np.array_equal(arr, c[key]):
                msgs.append(f"{key}: {int(np.sum(arr != c[key]))} mismatches")
        if not same(v, float(c["value"])):
            msgs.append(f"value {v!r} vs {float(c['value'])!r}")
        if msgs:
            bad += 1; print("MISMATCH", n, "flags", flags, msgs)
print("quick_parity: bad =", bad)
if "--no-timing" in sys.argv:
    sys.exit(1 if bad else 0)
# a first timing: N=1e6 bivariate Gaussian
rng = np.random.default_rng(0)
for N in (100_000, 1_000_000):
    d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=N)
    co = nat.pack_coords([d[:, 0], d[:, 1]])
    for flags in (0, nat.FLAG_NO_PRUNE):
        if flags and N > 300_000 and "--brute" not in sys.argv: continue
        nat.ksg_mi(co, 3, flags=flags)
        t0 = time.perf_counter(); v = nat.ksg_mi(co, 3, flags=flags); t1 = time.perf_counter()
        print("N", N, "flags", flags, "mi", v, "wall_ms", (t1 - t0) * 1e3, nat.last_timing())
sys.exit(1 if bad else 0)
