import sys; sys.path.insert(0,'.')
import numpy as np, xerus_b200 as xb
from oracle import tt_oracle as O
xb.init(0)
BW=xb.blasWrapper
rng=np.random.default_rng(5)
def rel(a,b): return np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-300)
A=rng.standard_normal((5,6,3,1,4,2,8,1))
o=O.tt_svd(A,1e-14); o.move_core(7)
for i in range(4): o.round_edge(7-i,6-i,0,1e-14)
F,T=o.cores[3],o.cores[2]
print('F',F.shape,'T',T.shape)
r=F.shape[0]
Fm=F.reshape(r,-1); Tm=T.reshape(-1,r)
print('Fm structure: upper-tri?',np.allclose(Fm,np.triu(Fm)),'lower?',np.allclose(Fm,np.tril(Fm)),'cond %.2e'%np.linalg.cond(Fm))
cA,Fq,ra=BW.cq(Fm); print('cq rank',ra,'recon %.2e orth %.2e'%(rel(cA@Fq,Fm),np.linalg.norm(Fq@Fq.T-np.eye(ra))))
Tq,cB,rb=BW.qc(Tm); print('qc rank',rb,'recon %.2e orth %.2e'%(rel(Tq@cB,Tm),np.linalg.norm(Tq.T@Tq-np.eye(rb))))
X=BW.matrix_matrix_product(1.0,cA,True,cB,True); print('X err %.2e'%rel(X,cA.T@cB.T), X.shape)
U,S,Vt=BW.svd(X); print('svd recon %.2e'%rel((U*S)@Vt,X),'S',S[:3],S[-3:])
k=O.truncation_rank(S,0,1e-14); print('k',k)
newF=BW.matrix_matrix_product(1.0,U[:,:k],True,Fq,False); newT=BW.matrix_matrix_product(1.0,Tq,False,(S[:k,None]*Vt[:k]),True)
print('pair recon %.2e'%rel(newT@newF, Tm@Fm))
# same with numpy
cA2,Fq2,_=O.cq(Fm,False); Tq2,cB2,_=O.qc(Tm,False); X2=cA2.T@cB2.T; U2,S2,Vt2=np.linalg.svd(X2)
print('numpy pair recon %.2e'%rel((Tq2@((S2[:,None]*Vt2).T))@(U2.T@Fq2), Tm@Fm))
