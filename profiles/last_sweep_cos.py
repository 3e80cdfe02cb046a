import sys, os; sys.path.insert(0, os.getcwd())
import numpy as np, xerus_b200 as xb
xb.init(0)
rng = np.random.default_rng(0)
base = xb.TTTensor.random([2]*32, 256, rng)
import time
for rep in range(3):
    t = base.copy(); xb.synchronize(); t0=time.perf_counter(); sv = t.round(128); xb.synchronize(); dt=time.perf_counter()-t0
print("round ms %.2f"%(dt*1e3), t.ranks()[:8])
ref = base.copy(); xb.set_option("svd_last_sweep_cos", 1e-7); ref.round(128)
print("rel distance to the 1e-7 result: %.2e" % (t.distance(ref)/ref.frob_norm()), " ||x-round(x)||/||x|| %.6e"%(t.distance(base)/base.frob_norm()))
