"""Profiling driver: SVDs of one matrix shape through the C ABI (used under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import xerus_b200 as xb
m = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
xb.init(0)
rng = np.random.default_rng(0)
A = rng.standard_normal((m, n))
for _ in range(reps):
    U, S, Vt = xb.blasWrapper.svd(A)
print("ok", S[:3], np.abs((U * S) @ Vt - A).max())
