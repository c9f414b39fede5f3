"""Developer check of the bivariate pipeline (run under gpurun): eps / n_x / n_y of the pipeline against the
brute-force kernels of the same library on data sets that stress it, then phase timings."""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from ennemi_b200 import _native as nat  # noqa: E402

BRUTE = nat.FLAG_NO_PRUNE | nat.FLAG_BRUTE_COUNT


def compare(name, x, y, k):
    co = nat.pack_coords([x, y])
    try:
        v, d = nat.ksg_mi(co, k, details=True)
        vb, b = nat.ksg_mi(co, k, flags=BRUTE, details=True)
    except Exception as e:           # noqa: BLE001
        print("EXC", name, type(e).__name__, e)
        return 1
    msgs = [f"{key}: {int(np.sum(d[key] != b[key]))} mismatches (first at {np.flatnonzero(d[key] != b[key])[:5]})"
            for key in ("eps", "nx", "ny") if not np.array_equal(d[key], b[key])]
    if not ((v == vb) or (np.isnan(v) and np.isnan(vb)) or abs(v - vb) <= 1e-10):
        msgs.append(f"value {v!r} vs {vb!r}")
    print(("MISMATCH " if msgs else "ok ") + name, "n", len(x), "k", k, "mi", v, msgs, nat.last_timing()["launches"])
    return 1 if msgs else 0


def main():
    rng = np.random.default_rng(0)
    bad = 0
    for n in (3000, 20_000, 150_000):
        d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=n)
        for k in (1, 3, 5, 7):
            bad += compare("gauss", d[:, 0], d[:, 1], k)
    t = rng.standard_t(2, size=(100_000, 2))
    bad += compare("student-t2", t[:, 0], t[:, 1], 3)
    u = rng.uniform(size=(50_000, 2)); u[:, 1] = u[:, 0] + 1e-3 * u[:, 1]
    bad += compare("thin-band", u[:, 0], u[:, 1], 3)
    g = np.round(rng.normal(size=(30_000, 2)), 1)                        # heavy ties (a few dozen distinct values)
    bad += compare("ties", g[:, 0], g[:, 1], 3)
    g0 = np.round(rng.normal(size=(30_000, 2)), 0)                        # buckets overflow: general path takes over
    bad += compare("overflow", g0[:, 0] + 1e-9 * rng.normal(size=30_000), g0[:, 1], 3)
    g2 = np.round(rng.normal(size=(30_000, 2)), 3)
    bad += compare("some-ties", g2[:, 0], g2[:, 1], 3)
    c = rng.normal(size=(40_000, 2)); c[::7] = c[0]                       # many exact duplicates of one point
    bad += compare("duplicates", c[:, 0], c[:, 1], 3)
    s = np.sort(rng.normal(size=60_000)); bad += compare("sorted-x", s, rng.normal(size=60_000), 3)
    o = rng.normal(size=(80_000, 2)); o[5] = (40.0, -35.0); o[6] = (0.0, 30.0)   # far outliers
    bad += compare("outliers", o[:, 0], o[:, 1], 3)
    bm = np.concatenate([rng.normal(-5, 0.1, size=(30_000, 2)), rng.normal(5, 0.1, size=(30_000, 2)), [[0.0, 0.0]]])
    bad += compare("bimodal+lonely", bm[:, 0], bm[:, 1], 3)
    big = rng.normal(size=(70_000, 2)) * 1e15 + 1e15
    bad += compare("huge-offset", big[:, 0], big[:, 1], 3)
    print("k2_check: bad =", bad)
    if "--no-timing" not in sys.argv:
        for n in (100_000, 1_000_000):
            d = rng.multivariate_normal([0, 0], [[1, .6], [.6, 1]], size=n)
            co = nat.pack_coords([d[:, 0], d[:, 1]])
            for env in ({}, {"EB2_NO_K2": "1"}):
                os.environ.pop("EB2_NO_K2", None)
                os.environ.update(env)
                for _ in range(3):
                    nat.ksg_mi(co, 3)
                t0 = time.perf_counter(); v = nat.ksg_mi(co, 3); t1 = time.perf_counter()
                print("N", n, env, "mi", v, "wall_ms", round((t1 - t0) * 1e3, 3), nat.last_timing())
            os.environ.pop("EB2_NO_K2", None)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
