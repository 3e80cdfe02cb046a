import sys, time, ctypes as C, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import xerus_b200 as xb
from xerus_b200 import _lib
import torch
xb.init(0)
rng = np.random.default_rng(3)
for (m, n) in [(32, 64), (64, 32), (32, 32), (16, 64), (4, 64), (16, 16), (128, 32), (256, 32)]:
    for small in (1, 0):
        xb.set_option("small_kernels", small)
        A = rng.standard_normal((m, n))
        k = min(m, n)
        dA = torch.from_numpy(A).cuda(); U = torch.empty(m, k, dtype=torch.float64, device="cuda"); S = torch.empty(k, dtype=torch.float64, device="cuda"); Vt = torch.empty(k, n, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        sw = C.c_int()
        ts = []
        for rep in range(5):
            t0 = time.perf_counter()
            _lib.call("xb_dev_svd", C.c_void_p(U.data_ptr()), C.c_void_p(S.data_ptr()), C.c_void_p(Vt.data_ptr()), C.c_void_p(dA.data_ptr()), m, n, k, 0, 0, C.byref(sw))
            xb.synchronize(); ts.append((time.perf_counter() - t0) * 1e6)
        Un, Sn, Vn = U.cpu().numpy(), S.cpu().numpy(), Vt.cpu().numpy()
        s_ref = np.linalg.svd(A, compute_uv=False)
        rec = np.linalg.norm(Un * Sn @ Vn - A) / np.linalg.norm(A)
        print("%dx%d small=%d sweeps=%d wall us=%.0f  sv err %.1e recon %.1e orthU %.1e orthV %.1e" % (m, n, small, sw.value, min(ts), np.max(np.abs(Sn - s_ref)) / s_ref[0], rec,
              np.linalg.norm(Un.T @ Un - np.eye(k)), np.linalg.norm(Vn @ Vn.T - np.eye(k))), flush=True)
