import sys, time, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import xerus_b200 as xb
xb.init(0)
def bench(dims, r, target, reps=6):
    rng = np.random.default_rng(1)
    base = xb.TTTensor.random(dims, r, rng)
    out = {}
    for plans in (0, 1):
        xb.set_option("round_plans", plans)
        ts = []
        for i in range(reps):
            c = base.copy(); xb.synchronize()
            l0 = xb.kernel_launch_count()
            t0 = time.perf_counter(); c.round(target); xb.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
            l1 = xb.kernel_launch_count()
        out[plans] = (ts, l1 - l0, c)
        print("dims %dx%d r=%d->%d plans=%d  ms per round:" % (dims[0], len(dims), r, target, plans), ["%.3f" % t for t in ts], "launches", l1 - l0, flush=True)
    a, b = out[0][2].cores(), out[1][2].cores()
    print("   identical:", all(np.array_equal(x, y) for x, y in zip(a, b)))
bench([4] * 8, 32, 16)
bench([4] * 12, 128, 64)
bench([2] * 32, 256, 128, reps=4)
