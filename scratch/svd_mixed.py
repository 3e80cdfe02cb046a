import sys; sys.path.insert(0,'.')
import numpy as np, xerus_b200 as xb
xb.init(0)
rng=np.random.default_rng(0)
def run(A,label):
    for mixed in [0,1]:
        xb.set_option("svd_mixed",mixed)
        U,S,Vt=xb.blasWrapper.svd(A)
        xb.profile_enable(True)
        for _ in range(3): U,S,Vt=xb.blasWrapper.svd(A)
        sc,l,ms=xb.profile_get("svd"); xb.profile_enable(False)
        k=min(A.shape); Sl=np.linalg.svd(A,compute_uv=False)
        print(label,A.shape,'mixed',mixed,'ms %.3f'%(ms/sc),'recon %.1e'%(np.linalg.norm((U*S)@Vt-A)/np.linalg.norm(A)),'Uorth %.1e Vorth %.1e'%(np.linalg.norm(U.T@U-np.eye(k)),np.linalg.norm(Vt@Vt.T-np.eye(k))),'S err %.1e'%(np.abs(S-Sl).max()/Sl[0]),flush=True)
for (m,n) in [(256,256),(512,512),(128,128),(64,64),(300,100)]:
    run(rng.standard_normal((m,n)),'gauss')
run(rng.standard_normal((256,256))*1e33,'huge')
Q1,_=np.linalg.qr(rng.standard_normal((256,256))); Q2,_=np.linalg.qr(rng.standard_normal((256,256)))
run((Q1*np.logspace(0,-14,256))@Q2.T,'graded')
run(rng.standard_normal((256,40))@rng.standard_normal((40,256)),'rank40')
xb.set_option("svd_mixed",1)
