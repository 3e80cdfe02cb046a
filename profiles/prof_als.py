"""Profiling driver: one ALS_SPD sweep at BASELINE config 2 shape (rank from argv) — used under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, xerus_b200 as xb
r = int(sys.argv[1]) if len(sys.argv) > 1 else 50
xb.init(0)
d, n = 16, 10
rng = np.random.default_rng(16)
A, b = xb.TTOperator.laplace(d, n), xb.TTTensor.ones([n] * d)
x = xb.TTTensor.random([n] * d, r, rng)
v = xb.ALSVariant(1, 0, True)
e = v(A, x, b, 2)
print("energy", e, "cg its", v.last_local_iterations)
