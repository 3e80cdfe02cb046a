import sys, time; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, xerus_b200 as xb
xb.init(0)
rng=np.random.default_rng(0)
for (m,n) in [(256,256),(512,512),(128,128),(64,64),(32,32)]:
    A=rng.standard_normal((m,n))
    for bw in [0,16,8,4]:
        xb.set_option("svd_max_bw",bw); wpp=1
        try:
            xb.blasWrapper.svd(A)
            xb.profile_enable(True)
            for _ in range(3): U,S,Vt=xb.blasWrapper.svd(A)
            sc,l,ms=xb.profile_get("svd_jacobi"); sc2,l2,ms2=xb.profile_get("svd")
            xb.profile_enable(False)
            err=np.linalg.norm((U*S)@Vt-A)/np.linalg.norm(A)
            print(m,n,'bw',bw,'wpp',wpp,'jacobi ms/svd %.3f  svd total %.3f'%(ms/sc,ms2/sc2),'err %.1e'%err, flush=True)
        except Exception as e:
            print(m,n,bw,wpp,'ERR',e)
xb.set_option("svd_max_bw",0)
