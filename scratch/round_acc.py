import sys; sys.path.insert(0,'.')
import numpy as np, xerus_b200 as xb
from oracle import tt_oracle as O
xb.init(0)
rng=np.random.default_rng(5)
def rel(a,b): return np.linalg.norm(a-b)/((np.linalg.norm(a)+np.linalg.norm(b))/2)
for dims in [(5,6,3,1,4,2,8,1),(2,)*8,(8,8,8,8)]:
    A=rng.standard_normal(dims)
    t=xb.TTTensor.from_dense(A,1e-14)
    o=O.tt_svd(A,1e-14)
    print(dims,'ttsvd err gpu %.2e oracle %.2e'%(rel(t.to_dense(),A),rel(o.to_dense(),A)), t.ranks())
    t.round(1e-14); o.round(None,1e-14)
    print('   round(eps) err gpu %.2e oracle %.2e'%(rel(t.to_dense(),A),rel(o.to_dense(),A)), t.ranks(), o.ranks())
    t.round(576); o.round(576)
    print('   round(576) err gpu %.2e oracle %.2e'%(rel(t.to_dense(),A),rel(o.to_dense(),A)))
# QR / gemm accuracy
for (m,n) in [(720,8),(90,64),(512,256),(30,24)]:
    M=rng.standard_normal((m,n))
    Q,R=xb.blasWrapper.qr(M); Qo,Ro=np.linalg.qr(M)
    print('qr',m,n,'gpu %.2e lapack %.2e'%(rel(Q@R,M),rel(Qo@Ro,M)),'orth gpu %.2e lapack %.2e'%(np.linalg.norm(Q.T@Q-np.eye(min(m,n))),np.linalg.norm(Qo.T@Qo-np.eye(min(m,n)))))
