import numpy as np, sys
sys.path.insert(0,'/root/repo/profiles')
rng=np.random.default_rng(0)
d,n,r=32,2,256
ranks=[1]
for i in range(1,d): ranks.append(min(r, n**i, n**(d-i)))
ranks.append(1)
cores=[rng.standard_normal((ranks[i],n,ranks[i+1])) for i in range(d)]
for k in range(d-1):
    a,_,b=cores[k].shape
    Q,R=np.linalg.qr(cores[k].reshape(a*n,b))
    cores[k]=Q.reshape(a,n,-1)
    cores[k+1]=np.tensordot(R,cores[k+1],axes=(1,0))
def jacobi_sweeps(A, big=1e-7, maxs=60):
    A=A.copy(); m,n_=A.shape
    tol=np.sqrt(m)*2.2e-16
    for sweep in range(1,maxs+1):
        nbig=0
        perm=np.arange(n_)
        for rr in range(n_-1):
            a=perm[:n_//2]; b=perm[n_//2:][::-1]
            X=A[:,a]; Y=A[:,b]
            aa=(X*X).sum(0); bb=(Y*Y).sum(0); g=(X*Y).sum(0)
            ab=aa*bb
            rot = g*g > tol*tol*ab
            nbig += int((g*g > big*big*ab).sum())
            dd=bb-aa; h=np.sqrt(dd*dd+4*g*g); h[h==0]=1
            c2=0.5+0.5*np.abs(dd)/h; c=np.sqrt(c2); s=np.where(dd>=0,g,-g)/(h*c)
            c=np.where(rot,c,1.0); s=np.where(rot,s,0.0)
            A[:,a]=c*X-s*Y; A[:,b]=s*X+c*Y
            perm=np.concatenate(([perm[0]],[perm[-1]],perm[1:-1]))
        if nbig==0: return sweep
    return maxs
import scipy.linalg as sl
for k in range(d-1,0,-1):
    a,_,b=cores[k].shape
    M=cores[k].reshape(a,n*b)
    if k in (22, 18):
        Mw = M if M.shape[0]>=M.shape[1] else M.T
        o=np.argsort(-np.linalg.norm(Mw,axis=0)); Qs,Rs=np.linalg.qr(Mw[:,o])
        Qp,Rp,P=sl.qr(Mw,pivoting=True)
        Q2,R2=np.linalg.qr(Rs.T)          # second QR: Rs^T = Q2 R2 ; X = R2^T lower triangular
        Q3,R3=np.linalg.qr(Rp.T)
        res={'now: sorted+QR, rows of R':jacobi_sweeps(Rs.T.copy()),
             'DV: sorted+QR, then QR(R^T): cols of R2^T':jacobi_sweeps(R2.T.copy()), 'rows of R2^T':jacobi_sweeps(R2.copy()),
             'true QRCP rows of R':jacobi_sweeps(Rp.T.copy()), 'QRCP + second QR cols of R3^T':jacobi_sweeps(R3.T.copy())}
        print("edge",k,M.shape,res,flush=True)
    U,S,Vt=np.linalg.svd(M,full_matrices=False)
    kk=min(128,len(S))
    cores[k]=Vt[:kk].reshape(kk,n,b)
    cores[k-1]=np.tensordot(cores[k-1],U[:,:kk]*S[:kk],axes=(2,0))
    if k<18: break
