"""Profiling driver: one warm round + one profiled round of a BASELINE workload (used under ncu, see profiles/README.md)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import xerus_b200 as xb

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
d, n, r, target = {"c3": (32, 2, 256, 128), "c1": (8, 4, 32, 16), "c3s": (14, 2, 256, 128)}[wl]
xb.init(0)
rng = np.random.default_rng(0)
base = xb.TTTensor.random([n] * d, r, rng)
for rep in range(2):
    t = base.copy()
    xb.synchronize()
    n0 = xb.kernel_launch_count()
    t.round(target)
    xb.synchronize()
    print("round", rep, "launches", xb.kernel_launch_count() - n0, "ranks", t.ranks()[:10])
