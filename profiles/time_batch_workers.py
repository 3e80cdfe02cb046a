import sys, time, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import xerus_b200 as xb
xb.init(0)
d, n, r = 12, 4, 64
A = xb.TTOperator.laplace(d, n)
rng = np.random.default_rng(5)
xs = [xb.TTTensor.random([n] * d, r, rng) for _ in range(64)]
ys0 = [A.apply(x) for x in xs]
ysc = []
for y in ys0:
    c = y.copy(); c.move_core(0); ysc.append(c)
print("ranks raw", ys0[0].ranks(), "canon", ysc[0].ranks())
for plans in (0, 1):
  xb.set_option("round_plans", plans)
  for w in (1, 2, 4, 8, 16):
    xb.set_option("batch_workers", w)
    for what, src in (("round raw", ys0), ("round canon", ysc)):
        for rep in range(3):
            ys = [y.copy() for y in src]
            xb.synchronize()
            t0 = time.perf_counter()
            xb.round_batched(ys, r)
            xb.synchronize()
            dt = time.perf_counter() - t0
        print("plans=%d workers=%2d %-12s %.2f ms/item -> %.0f items/s" % (plans, w, what, dt / len(ys) * 1e3, len(ys) / dt), flush=True)
