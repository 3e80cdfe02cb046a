"""Jacobi block width experiment (svd_max_bw): columns per block, i.e. warps per CTA, against SVD time."""
import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, xerus_b200 as xb
xb.init(0)
rng=np.random.default_rng(0)
shapes = [tuple(map(int, a.split('x'))) for a in sys.argv[1:]] or [(512,256),(256,128),(128,64)]
for (m,n) in shapes:
    A=rng.standard_normal((m,n))
    for bw in [2,4,8,16]:
        xb.set_option("svd_max_bw",bw)
        xb.blasWrapper.svd(A)
        xb.profile_enable(True)
        for _ in range(3): U,S,Vt=xb.blasWrapper.svd(A)
        sc,l,ms=xb.profile_get("svd_jacobi")
        xb.profile_enable(False)
        err=np.linalg.norm((U*S)@Vt-A)/np.linalg.norm(A)
        print(m,n,'max_bw',bw,'jacobi ms/svd %.3f'%(ms/sc),'recon %.1e'%err,flush=True)
