"""Developer timing of the three-level grid against the general path: k-NN entropy in D dimensions."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from ennemi_b200 import _native as nat
    if os.environ.get("EB2_LIB"):
        nat.LIB_PATH = os.path.abspath(os.environ["EB2_LIB"])          # a build variant (developer experiments)
    rng = np.random.default_rng(0)
    for N, D, k in ((500_000, 4, 5), (500_000, 3, 3), (500_000, 5, 3), (200_000, 8, 3), (100_000, 4, 3), (20_000, 4, 3), (2_000_000, 4, 3)):
        x = rng.normal(size=(N, D)) @ rng.normal(size=(D, D))
        co = nat.pack_coords([x])
        best = None
        for _ in range(4):
            v = nat.entropy(co, k)
            t = nat.last_timing()
            if best is None or t["total_ms"] < best["total_ms"]:
                best = t
        print(f"N={N} D={D} k={k} pipeline={nat.last_pipeline()} total={best['total_ms']:.2f} knn={best['knn_ms']:.2f} layout={best['layout_ms']:.2f} v={v!r}")
    for name, x, k in (("student-t2", rng.standard_t(2, size=(500_000, 4)), 3), ("rounded", np.round(rng.normal(size=(500_000, 4)), 2), 3),
                       ("clusters", np.concatenate([rng.normal(m, 0.05, size=(125_000, 4)) for m in (-6, -2, 2, 6)]), 3),
                       ("outliers", np.concatenate([rng.normal(size=(499_950, 4)), rng.normal(size=(50, 4)) * 1e4]), 3),
                       ("lognormal", rng.lognormal(size=(500_000, 4)), 3), ("student-t3", rng.standard_t(3, size=(500_000, 4)), 3)):
        co = nat.pack_coords([np.ascontiguousarray(x)])
        best = None
        for _ in range(3):
            v = nat.entropy(co, k)
            t = nat.last_timing()
            if best is None or t["total_ms"] < best["total_ms"]:
                best = t
        print(f"{name} N={len(x)} D=4 k={k} pipeline={nat.last_pipeline()} total={best['total_ms']:.2f} knn={best['knn_ms']:.2f} v={v!r}")
else:
    for env in ({},) if os.environ.get("EB2_LIB") else ({}, {"EB2_NO_G3": "1"}):
        print("==", env)
        sys.stdout.flush()
        subprocess.run([sys.executable, __file__, "child"], env={**os.environ, **env})
