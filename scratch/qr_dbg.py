import sys; sys.path.insert(0,'.')
import numpy as np, xerus_b200 as xb
xb.init(0)
BW=xb.blasWrapper
rng=np.random.default_rng(1)
def rel(a,b): return np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-300)
Qo,_=np.linalg.qr(rng.standard_normal((64,64)))
for name,M in [('orthogonal',Qo),('uppertri',np.triu(rng.standard_normal((64,64)))),('diag',np.diag(rng.standard_normal(64))),('identity',np.eye(64)),('general',rng.standard_normal((64,64)))]:
    Q,R=BW.qr(M)
    print(name,'recon %.2e'%rel(Q@R,M),'orth %.2e'%np.linalg.norm(Q.T@Q-np.eye(64)),'|Q|==I?',np.allclose(np.abs(Q),np.eye(64)),'Rdiag',np.abs(np.diag(R))[:3])
    Q2,C,r=BW.qc(M); print('   qc rank',r,'recon %.2e'%rel(Q2@C,M),'orth %.2e'%np.linalg.norm(Q2.T@Q2-np.eye(r)))
    C3,Q3,r3=BW.cq(M); print('   cq rank',r3,'recon %.2e'%rel(C3@Q3,M),'orth %.2e'%np.linalg.norm(Q3@Q3.T-np.eye(r3)))
