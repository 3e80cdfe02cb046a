"""One ALS_SPD full sweep at BASELINE config 2 with the library's CUDA-event classes switched on: where a sweep's time goes."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, xerus_b200 as xb
xb.init(0)
d, n, r = 16, 10, int(sys.argv[1]) if len(sys.argv) > 1 else 50
rng = np.random.default_rng(16)
A, b = xb.TTOperator.laplace(d, n), xb.TTTensor.ones([n] * d)
x0 = xb.TTTensor.random([n] * d, r, rng)
for rep in range(2):
    x = x0.copy(); v = xb.ALSVariant(1, 0, True)
    if rep == 1: xb.profile_enable(True)
    xb.synchronize(); t0 = time.perf_counter()
    e = v(A, x, b, 2)
    xb.synchronize(); dt = time.perf_counter() - t0
print("r %d sweep %.1f ms (with class timers on), cg its %d" % (r, dt * 1e3, v.last_local_iterations))
for c in ["als_local_step", "als_cg_kernel", "als_move_to_next", "als_energy", "gemm", "qr", "svd", "mid_apply"]:
    sc, ln, ms = xb.profile_get(c)
    print("  %-18s scopes %5d launches %6d  %8.3f ms" % (c, sc, ln, ms))
