"""Single process driving every GPU of the box: pairwise fan-out over worker threads / lanes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ennemi_b200 as eb
from ennemi_b200 import _native as nat, _devices
print("devices:", nat.device_count(), "visible:", _devices.visible())
data = np.random.default_rng(0).normal(size=(100_000, 32))
for devs in ("0", "1", None, None, "0"):
    if devs is None:
        os.environ.pop("ENNEMI_B200_DEVICES", None)
    else:
        os.environ["ENNEMI_B200_DEVICES"] = devs
    eb.pairwise_mi(data[:, :6])
    for rep in range(2):
        t0 = time.perf_counter(); out = eb.pairwise_mi(data); t1 = time.perf_counter()
        print("devices", devs, "rep", rep, "%.3f s" % (t1 - t0))
