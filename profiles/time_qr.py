"""QR timing / accuracy: cluster panel kernel (registers + DSMEM) against the one-CTA shared-memory panel kernel."""
import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, xerus_b200 as xb
xb.init(0)
rng=np.random.default_rng(0)
shapes = [tuple(map(int, a.split('x'))) for a in sys.argv[1:]] or [(512,256),(512,32),(130,40),(256,256),(1000,100),(2048,64),(1500,33),(128,128),(3000,64)]
for (m,n) in shapes:
    A=rng.standard_normal((m,n))
    for cl in ([1] if len(sys.argv) > 1 else [0,1]):
        xb.set_option("qr_cluster",cl)
        xb.blasWrapper.qr(A)
        xb.profile_enable(True)
        for _ in range(5): Q,R=xb.blasWrapper.qr(A)
        sc,l,ms=xb.profile_get("qr")
        xb.profile_enable(False)
        k=min(m,n)
        err=np.linalg.norm(Q@R-A)/np.linalg.norm(A)
        print(m,n,'cluster',cl,'qr ms %.3f'%(ms/sc),'launches',l//sc,'recon %.1e'%err,'orth %.1e'%np.abs(Q.T@Q-np.eye(k)).max(),'tril %.1e'%np.abs(np.tril(R,-1)).max(),flush=True)
xb.set_option("qr_cluster",1)
