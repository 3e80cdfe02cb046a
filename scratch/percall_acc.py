import sys; sys.path.insert(0,'.')
import numpy as np, xerus_b200 as xb
from oracle import tt_oracle as O
xb.init(0)
BW=xb.blasWrapper
rng=np.random.default_rng(5)
def rel(a,b): return np.linalg.norm(a-b)/((np.linalg.norm(a)+np.linalg.norm(b))/2)
A=rng.standard_normal((5,6,3,1,4,2,8,1))
# LAPACK oracle
o=O.tt_svd(A,1e-14); o.round(None,1e-14); print('lapack oracle', rel(o.to_dense(),A))
# patch oracle's L0 with the GPU per-call layer (keeps the reference's algorithm, swaps the boundary)
O.matrix_matrix_product=lambda alpha,A_,ta,B_,tb: BW.matrix_matrix_product(alpha,A_,ta,B_,tb)
O.qr=lambda M: BW.qr(M)
O.rq=lambda M: BW.rq(M)
O.qc=lambda M,signed_quirk=True: BW.qc(M)
O.cq=lambda M,signed_quirk=True: BW.cq(M)
O.svd=lambda M: BW.svd(M)
o=O.tt_svd(A,1e-14); print('gpu per-call ttsvd', rel(o.to_dense(),A), o.ranks())
for i,c in enumerate(o.cores): pass
o2=o.copy(); o2.round(None,1e-14); print('gpu per-call round(eps)', rel(o2.to_dense(),A), o2.ranks())
o3=o.copy(); o3.move_core(7); print('move_core(7)', rel(o3.to_dense(),A), o3.ranks())
# individual factorization accuracy on the shapes involved
for c in o.cores:
    M=c.reshape(-1,c.shape[-1])
    Q,C,r=BW.qc(M); print('qc',M.shape,r,'recon %.2e orth %.2e'%(rel(Q@C,M),np.linalg.norm(Q.T@Q-np.eye(r))))
    M2=c.reshape(c.shape[0],-1)
    C2,Q2,r2=BW.cq(M2); print('cq',M2.shape,r2,'recon %.2e orth %.2e'%(rel(C2@Q2,M2),np.linalg.norm(Q2@Q2.T-np.eye(r2))))
print('---- stepwise')
o2=o.copy(); o2.move_core(7)
_svd=O.svd
def svd_chk(M):
    U,S,Vt=_svd(M)
    Sl=np.linalg.svd(M,compute_uv=False)
    print('   svd',M.shape,'recon %.2e'%rel((U*S)@Vt,M),'S relerr %.2e'%(np.abs(S-Sl).max()/Sl[0]),'Uorth %.2e Vorth %.2e'%(np.linalg.norm(U.T@U-np.eye(len(S))),np.linalg.norm(Vt@Vt.T-np.eye(len(S)))),'cond %.1e'%(Sl[0]/Sl[-1]))
    return U,S,Vt
O.svd=svd_chk
d=8
for i in range(d-1):
    o2.round_edge(d-1-i,d-2-i,0,1e-14)
    print('edge',d-2-i,'err %.2e'%rel(o2.to_dense(),A), o2.ranks())
