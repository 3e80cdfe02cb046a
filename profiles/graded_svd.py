import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, xerus_b200 as xb
xb.init(0)
rng=np.random.default_rng(0)
xb.set_option("svd_mixed",0)
for n in [96,256]:
    Q1,_=np.linalg.qr(rng.standard_normal((n,n))); Q2,_=np.linalg.qr(rng.standard_normal((n,n)))
    for decades in [6,10,14]:
        sig=np.logspace(0,-decades,n); A=(Q1*sig)@Q2.T
        for pers in [0,1]:
            xb.set_option("svd_persistent",pers)
            xb.set_option("svd_max_sweeps",60)
            try:
                U,S,Vt=xb.blasWrapper.svd(A)
                print(n,decades,'persistent',pers,'ok  S abs err %.1e'%np.abs(S-sig).max(),'recon %.1e'%(np.linalg.norm((U*S)@Vt-A)/np.linalg.norm(A)),flush=True)
            except Exception as e:
                print(n,decades,'persistent',pers,'FAILED',str(e)[:60],flush=True)
