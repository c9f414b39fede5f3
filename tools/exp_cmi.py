import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ennemi_b200 import _native as nat
rng = np.random.default_rng(0)
N = 200_000
z = rng.normal(size=(N, 3)); x = rng.normal(size=N) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=N)
co = nat.pack_coords([x, y, z])
for _ in range(2):
    v = nat.cmi(co, 3)
    print(v, nat.last_timing())
