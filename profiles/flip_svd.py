import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, xerus_b200 as xb, subprocess
xb.init(0)
rng=np.random.default_rng(0)
for (m,n,kind) in [(512,256,'gauss'),(512,256,'decay'),(256,256,'gauss'),(256,256,'decay'),(128,128,'gauss'),(128,128,'decay'),(64,64,'decay')]:
    A=rng.standard_normal((m,n))
    if kind=='decay':
        k=min(m,n); A=rng.standard_normal((m,k))@np.diag(np.logspace(0,-10,k))@rng.standard_normal((k,n))
    for flip in [0,1]:
        xb.set_option("svd_flip",flip)  # flip=0 also disables the QR step of square inputs
        xb.blasWrapper.svd(A)
        xb.profile_enable(True)
        for _ in range(3): U,S,Vt=xb.blasWrapper.svd(A)
        sc,l,ms=xb.profile_get("svd")
        xb.profile_enable(False)
        err=np.linalg.norm((U*S)@Vt-A)/np.linalg.norm(A)
        Sref=np.linalg.svd(A,compute_uv=False)
        print(m,n,kind,'flip',flip,'svd ms %.3f'%(ms/sc),'recon %.1e'%err,'S err %.1e'%(np.abs(S-Sref).max()/S[0]),'orthU %.1e orthV %.1e'%(np.abs(U.T@U-np.eye(len(S))).max(),np.abs(Vt@Vt.T-np.eye(len(S))).max()),flush=True)
