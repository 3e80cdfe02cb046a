"""One config-5 item (x = random({4}x12, 64); y = A x; y.round(64)) twice on the ordinary path: the ncu target of r2_launches_c5_item.txt."""
import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np
import xerus_b200 as xb
xb.init(0)
xb.set_option("round_plans", 0)
rng = np.random.default_rng(5)
A = xb.TTOperator.laplace(12, 4)
for i in range(2):
    x = xb.TTTensor.random([4] * 12, 64, rng)
    y = A.apply(x)
    y.round(64)
    xb.synchronize()
