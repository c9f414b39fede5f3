"""Developer run for `compute-sanitizer --tool memcheck`: the bivariate pipeline on the heavy-tailed inputs of
tests/test_gpu_parity.py::test_hard_distributions, repeated (the slot order inside a bucket differs from run to run)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat
loops = int(sys.argv[1]) if len(sys.argv) > 1 else 50
for kind in ("cauchy", "outlier", "clusters"):
    rng = np.random.default_rng(len(kind))
    n = 6_000
    if kind == "cauchy":
        x = rng.standard_cauchy(n); y = x + rng.standard_cauchy(n)
    elif kind == "outlier":
        x = rng.normal(size=n); y = rng.normal(size=n); x[7] = 1e6; y[11] = -1e7
    else:
        c = rng.integers(0, 5, n); x = c * 100.0 + rng.normal(size=n) * 1e-3; y = c * -50.0 + rng.normal(size=n) * 1e-3
    co = nat.pack_coords([x, y])
    for i in range(loops):
        for k in (1, 3):
            try:
                v, d = nat.ksg_mi(co, k, details=True)
            except RuntimeError as e:
                print("FAILED", kind, i, k, str(e)[:200]); sys.exit(1)
    print(kind, "done", nat.last_pipeline())
