"""CPU (numpy) study behind DESIGN.md 3.8, on the reference's own config-3 input (`oracle/_ref/ref_bench round 32 2 256 128 1 dump`):

  1. conditioning of the left interfaces X_<k and of the unfoldings at every bond (what a Gram / Cholesky formulation would face);
  2. the *simultaneous* rounding that VERDICT round 1 proposed (all bonds truncated from the spectra of the ORIGINAL tensor, so that
     the 31 decompositions are independent) against the reference's sequential sweep: distance between the two results.

    python profiles/conditioning_c3.py          # needs oracle/_ref/ref_bench; a few minutes
"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.make_golden import read_container
from oracle import tt_oracle as O
from conftest import golden_tt

with tempfile.TemporaryDirectory() as td:
    dump = os.path.join(td, "c3.bin")
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_bench"), "round", "32", "2", "256", "128", "1", dump], check=True, capture_output=True)
    rec = read_container(dump)
cin, _ = golden_tt(rec, "in")
cref, _ = golden_tt(rec, "out")
d = len(cin)
R = np.ones((1, 1)); Rs = [None] * (d + 1)
for k in range(d - 1):                                   # X_<k+1 = Q R_{k+1}
    c = np.tensordot(R, cin[k], axes=([1], [0]))
    _, R = np.linalg.qr(c.reshape(-1, c.shape[-1]))
    Rs[k + 1] = R
L = np.ones((1, 1)); Ls = [None] * (d + 1)
for k in range(d - 1, 0, -1):                            # Z_k = L_k Q'
    c = np.tensordot(cin[k], L, axes=([2], [0]))
    _, Rt = np.linalg.qr(c.reshape(c.shape[0], -1).T)
    L = Rt.T
    Ls[k] = L
print("bond  r   kappa(X_<k)  kappa(unfolding)  s[r/2-1]/s0  s_min/s0")
proj = {}
for k in range(1, d):
    sR = np.linalg.svd(Rs[k], compute_uv=False)
    U, sU, _ = np.linalg.svd(Rs[k] @ Ls[k])
    r = len(sU)
    print("%4d %4d  %10.2e  %10.2e  %10.3f  %9.2e" % (k, r, sR[0] / sR[-1], sU[0] / sU[-1], sU[r // 2 - 1] / sU[0], sU[-1] / sU[0]))
    if r > 128:                                          # dominant left subspace of the ORIGINAL unfolding at this bond
        A = np.linalg.solve(Rs[k], U[:, :128])           # X_<k A has orthonormal columns spanning it
        B = U[:, :128].T @ Rs[k]                         # B A = I
        proj[k] = (A, B)
sim = []
for k in range(d):
    c = cin[k]
    if k in proj:
        c = np.tensordot(proj[k][1], c, axes=([1], [0]))
    if k + 1 in proj:
        c = np.tensordot(c, proj[k + 1][0], axes=([2], [0]))
    sim.append(c)
seq = O.TT(cref, core_position=0)
simtt = O.TT(sim)
full = O.TT(cin, core_position=0)
print("ranks of the simultaneous result:", simtt.ranks()[:10], "...")
print("|| simultaneous - sequential (reference) || / || reference ||  = %.3e" % O.tt_distance_rel(simtt, seq))
print("|| A - sequential || / ||A|| = %.6f   || A - simultaneous || / ||A|| = %.6f" % (O.tt_distance_rel(seq, full), O.tt_distance_rel(simtt, full)))
