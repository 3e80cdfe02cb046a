import sys, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import xerus_b200 as xb
xb.init(0)
xb.set_option("round_plans", 0)
rng = np.random.default_rng(1)
base = xb.TTTensor.random([4] * 8, 32, rng)
for i in range(2):
    c = base.copy(); c.round(16); xb.synchronize()
