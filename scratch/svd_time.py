import sys, time; sys.path.insert(0,'.')
import numpy as np, xerus_b200 as xb
xb.init(0)
rng=np.random.default_rng(0)
for (m,n) in [(256,256),(512,512),(128,128),(64,64)]:
    A=rng.standard_normal((m,n))
    for bw in [0,16,8,4]:
        xb.set_option("svd_max_bw",bw)
        xb.blasWrapper.svd(A)
        xb.profile_enable(True)
        for _ in range(3): U,S,Vt=xb.blasWrapper.svd(A)
        sc,l,ms=xb.profile_get("svd_jacobi")
        xb.profile_enable(False)
        err=np.abs((U*S)@Vt-A).max()
        print(m,n,'bw',bw,'jacobi ms/svd %.3f'%(ms/sc),'err %.1e'%err)
xb.set_option("svd_max_bw",0)
