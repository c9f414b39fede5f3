"""One call per general-path estimator on small inputs, for `compute-sanitizer --tool initcheck` (fresh process: the lane
workspace is one allocation reused across calls, so only first touches are checked)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat
which = sys.argv[1]
rng = np.random.default_rng(6)
n = 6_000
x = rng.standard_cauchy(n); y = x + rng.standard_cauchy(n)
if which == "cmi2":
    z = np.column_stack((y, x * 0.5 + rng.normal(size=n)))
    print(nat.cmi(nat.pack_coords([x, y, z]), 3, details=True)[0])
elif which == "cmi3":
    z = rng.normal(size=(n, 3))
    print(nat.cmi(nat.pack_coords([x, y, z]), 3, details=True)[0])
elif which == "cmi1":
    z = rng.normal(size=(n, 1))
    print(nat.cmi(nat.pack_coords([x, y, z]), 3, details=True)[0])
elif which == "ksg":
    print(nat.ksg_mi(nat.pack_coords([x[:1500], y[:1500]]), 3, details=True)[0])
elif which == "ksg20":
    print(nat.ksg_mi(nat.pack_coords([x, y]), 20, details=True)[0])
elif which == "ent4":
    print(nat.entropy(nat.pack_coords([rng.normal(size=(n, 4))]), 3, details=True)[0])
elif which == "k2":
    print(nat.ksg_mi(nat.pack_coords([x, y]), 3, details=True)[0])
elif which == "g3":
    os.environ["EB2_G3_MIN"] = "2"
    print(nat.entropy(nat.pack_coords([rng.normal(size=(n, 4))]), 3, details=True)[0])
