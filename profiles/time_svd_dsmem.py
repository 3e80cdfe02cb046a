"""Jacobi SVD: travelling-block hand-over through DSMEM (X workers in one cluster) against global memory + flags.
XB_JACOBI_TIMING=1 adds the kernel's clock64 phase counters."""
import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, xerus_b200 as xb
xb.init(0)
rng=np.random.default_rng(0)
shapes = [tuple(map(int, a.split('x'))) for a in sys.argv[1:]] or [(256,256),(512,256),(128,128),(64,64),(300,100),(512,512)]
for (m,n) in shapes:
    A=rng.standard_normal((m,n))
    for ds in [0,1]:
        xb.set_option("svd_dsmem",ds)
        xb.blasWrapper.svd(A)
        xb.profile_enable(True)
        for _ in range(3): U,S,Vt=xb.blasWrapper.svd(A)
        sc,l,ms=xb.profile_get("svd_jacobi"); sc2,l2,ms2=xb.profile_get("svd")
        xb.profile_enable(False)
        err=np.linalg.norm((U*S)@Vt-A)/np.linalg.norm(A)
        print(m,n,'dsmem',ds,'jacobi ms/svd %.3f  svd total %.3f'%(ms/sc,ms2/sc2),'recon %.1e'%err,'S err %.1e'%(np.abs(S-np.linalg.svd(A,compute_uv=False)).max()/S[0]), flush=True)
xb.set_option("svd_dsmem",1)
