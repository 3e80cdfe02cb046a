import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, xerus_b200 as xb
xb.init(0)
rng=np.random.default_rng(0)
for (m,n) in [(512,512),(256,256),(64,64),(16,16),(300,70),(70,300),(8,8)]:
    A=rng.standard_normal((m,n))
    for pol in [0,1]:
        xb.set_option("svd_polish",pol)
        U,S,Vt=xb.blasWrapper.svd(A)
        k=min(m,n)
        rec=np.linalg.norm((U*S)@Vt-A)/np.linalg.norm(A)
        print(m,n,'polish',pol,'recon rel %.2e'%rec,'U orth %.2e'%np.linalg.norm(U.T@U-np.eye(k)),'V orth %.2e'%np.linalg.norm(Vt@Vt.T-np.eye(k)),'S err %.2e'%(np.abs(S-np.linalg.svd(A,compute_uv=False)).max()/S[0]))
