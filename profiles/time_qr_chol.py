"""Cholesky-QR2 (qr_chol=1) against the Householder path (qr_chol=0) on the tall shapes of configs 2, 3 and 5: time per call
(CUDA events inside the library, class "qr"), launches, accuracy.  Usage: python profiles/time_qr_chol.py [MxN ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import xerus_b200 as xb

xb.init(0)
xb.set_option("qr_chol_min_rows", 0)          # every eligible shape, to find the crossover
rng = np.random.default_rng(0)
shapes = [tuple(map(int, a.split("x"))) for a in sys.argv[1:]] or [(512, 128), (768, 128), (1024, 128), (2048, 128), (256, 64), (512, 64), (1024, 64),
                                                                     (2048, 64), (500, 50), (1000, 50), (600, 33), (1500, 33), (4096, 16), (20000, 100)]
for (m, n) in shapes:
    A = rng.standard_normal((m, n))
    for on in (0, 1):
        xb.set_option("qr_chol", on)
        xb.blasWrapper.qr(A)
        xb.profile_enable(True)
        for _ in range(10):
            Q, R = xb.blasWrapper.qr(A)
        sc, l, ms = xb.profile_get("qr")
        tk = xb.profile_get("qr_chol")[0]
        parts = {k: xb.profile_get(k) for k in ("chol_gram", "chol_fact", "chol_apply", "chol_rr")}
        xb.profile_enable(False)
        err = np.linalg.norm(Q @ R - A) / np.linalg.norm(A)
        print("%5d x %3d  qr_chol %d  %.3f ms  %3d launches  taken %2d  recon %.1e  orth %.1e" %
              (m, n, on, ms / sc, l // sc, tk, err, np.linalg.norm(Q.T @ Q - np.eye(n))), flush=True)
        if on and tk:
            print("      per call, us: " + "  ".join("%s %.1f" % (k[5:], 1e3 * v[2] / tk) for k, v in parts.items()), flush=True)
xb.set_option("qr_chol", 1)
