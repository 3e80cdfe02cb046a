import sys, time, threading, ctypes as C, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import xerus_b200 as xb
from xerus_b200 import _lib
import torch
xb.init(0)
rng = np.random.default_rng(3)
def run(m, n, nthreads, reps, what):
    k = min(m, n)
    bufs = []
    for t in range(nthreads):
        A = torch.from_numpy(rng.standard_normal((m, n))).cuda()
        bufs.append((A, torch.empty(m, k, dtype=torch.float64, device="cuda"), torch.empty(k, dtype=torch.float64, device="cuda"), torch.empty(k, n, dtype=torch.float64, device="cuda")))
    torch.cuda.synchronize()
    def work(t):
        xb.worker_select(t + 1)
        A, U, S, Vt = bufs[t]
        sw = C.c_int()
        for r in range(reps):
            if what == "svd":
                _lib.call("xb_dev_svd", C.c_void_p(U.data_ptr()), C.c_void_p(S.data_ptr()), C.c_void_p(Vt.data_ptr()), C.c_void_p(A.data_ptr()), m, n, k, 0, 0, C.byref(sw))
            else:
                _lib.call("xb_dev_qr", C.c_void_p(U.data_ptr()), C.c_void_p(Vt.data_ptr()), C.c_void_p(A.data_ptr()), m, n)
        xb.synchronize()
    for warm in range(2):
        ths = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
        t0 = time.perf_counter()
        for th in ths: th.start()
        for th in ths: th.join()
        dt = time.perf_counter() - t0
    print("%s %dx%d threads=%d: %.2f ms per op per thread, %.0f ops/s total" % (what, m, n, nthreads, dt / reps * 1e3, nthreads * reps / dt), flush=True)
for nt in (1, 2, 4, 8, 16):
    run(256, 128, nt, 20, "svd")
for nt in (1, 2, 4, 8, 16):
    run(512, 128, nt, 20, "qr")
