"""Numpy emulation of the one-sided Jacobi sweep counts on the matrices a C3 round really hands to the SVD (first truncation
edges and an interior edge), for the four orientation choices and for norm-sorted columns; second part: the entropy rule that
was considered and rejected.  CPU only (a few minutes); see DESIGN.md 3.3 "Sorted columns"."""
import numpy as np, sys

rng=np.random.default_rng(0)
d,n,r=32,2,256
ranks=[1]
for i in range(1,d): ranks.append(min(r, n**i, n**(d-i)))
ranks.append(1)
cores=[rng.standard_normal((ranks[i],n,ranks[i+1])) for i in range(d)]
# orth sweep left->right
for k in range(d-1):
    a,_,b=cores[k].shape
    Q,R=np.linalg.qr(cores[k].reshape(a*n,b))
    cores[k]=Q.reshape(a,n,-1)
    cores[k+1]=np.tensordot(R,cores[k+1],axes=(1,0))
# truncation sweep right->left, keep matrices of interest
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
for k in range(d-1,0,-1):
    a,_,b=cores[k].shape
    M=cores[k].reshape(a,n*b)
    if a in (64,128,256) and k>d//2-1 and k>=d-9:
        S=np.linalg.svd(M,compute_uv=False)
        print("edge",k,"M",M.shape,"sigma max/min %.2e"%(S[0]/S[-1]), "row-norm spread %.1e col-norm spread %.1e"%(np.ptp(np.log10(np.linalg.norm(M,axis=1))),np.ptp(np.log10(np.linalg.norm(M,axis=0)))),flush=True)
        Mw = M if M.shape[0]>=M.shape[1] else M.T   # tall or square working matrix
        Q,R=np.linalg.qr(Mw);  Q2,R2=np.linalg.qr(Mw.T) if Mw.shape[0]==Mw.shape[1] else (None,None)
        o=np.argsort(-np.linalg.norm(Mw,axis=0)); Qs,Rs=np.linalg.qr(Mw[:,o])
        res={'cols of Mw':jacobi_sweeps(Mw),'QR(Mw): rows of R (before)':jacobi_sweeps(R.T.copy()),'QR(Mw): cols of R':jacobi_sweeps(R.copy()),
             'sorted cols + QR: rows of R (now)':jacobi_sweeps(Rs.T.copy())}
        if R2 is not None:
            res['QR(Mw^T): rows of R']=jacobi_sweeps(R2.T.copy()); res['QR(Mw^T): cols of R']=jacobi_sweeps(R2.copy())
        print("   ",res,flush=True)
    U,S,Vt=np.linalg.svd(M,full_matrices=False)
    kk=min(128,len(S))
    cores[k]=Vt[:kk].reshape(kk,n,b)
    cores[k-1]=np.tensordot(cores[k-1],U[:,:kk]*S[:kk],axes=(2,0))
    if k<d-10: break

print("---- rule check")
def ent(v):
    p=v/v.sum(); p=p[p>0]; return float(-(p*np.log(p)).sum())
def check(name,M):
    Mw = M if M.shape[0]>=M.shape[1] else M.T
    Q,R=np.linalg.qr(Mw)
    hc=ent((R*R).sum(0)); hr=ent((R*R).sum(1))
    f1=jacobi_sweeps(R.T.copy()); f0=jacobi_sweeps(R.copy())
    print("%-28s H_col %.3f H_row %.3f (log n %.3f)  flip1 %d  flip0 %d  -> rule picks %s"%(name,hc,hr,np.log(R.shape[0]),f1,f0,'flip0' if hc+0.1<hr else 'flip1'),flush=True)
n_=128
G=rng.standard_normal((n_,n_))
D=np.logspace(0,-8,n_)
check("gaussian",G)
check("col graded G*D",G*D)
check("row graded D*G",(G.T*D).T)
check("both graded D*G*D",(G.T*D).T*D)
check("tall 256x128 gaussian",rng.standard_normal((256,128)))
check("tall col graded",rng.standard_normal((256,128))*D)
check("tall row graded",(rng.standard_normal((256,128)).T*np.logspace(0,-8,256)).T)
U1,_=np.linalg.qr(rng.standard_normal((n_,n_))); V1,_=np.linalg.qr(rng.standard_normal((n_,n_)))
check("U diag(graded) V^T (dense)",(U1*D)@V1.T)
check("rank deficient",(G[:, :40]@rng.standard_normal((40,n_))))
