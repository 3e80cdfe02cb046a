"""blasWrapper::solve on SPD systems (blocked Cholesky) and general ones (LU): time and residual per size."""
import sys, time; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, xerus_b200 as xb
xb.init(0)
rng=np.random.default_rng(0)
for n,nrhs in [(65,1),(200,3),(640,1),(1000,1),(2000,2),(4000,1)]:
    M=rng.standard_normal((n,n)); A=M@M.T+n*np.eye(n); b=rng.standard_normal((n,nrhs))
    x=xb.blasWrapper.solve(A,b)
    t0=time.perf_counter(); x=xb.blasWrapper.solve(A,b); dt=time.perf_counter()-t0
    t1=time.perf_counter(); xr=np.linalg.solve(A,b); dn=time.perf_counter()-t1
    print('spd n',n,'nrhs',nrhs,'solve %.2f ms (incl. H2D of A)'%(dt*1e3),'numpy %.2f ms'%(dn*1e3),'residual %.1e'%(np.linalg.norm(A@x-b)/np.linalg.norm(b)),'vs numpy %.1e'%(np.linalg.norm(x-xr)/np.linalg.norm(xr)),flush=True)
for n,nrhs in [(65,1),(130,2),(640,1),(1000,3),(2000,1),(4000,1)]:
    A=rng.standard_normal((n,n)); b=rng.standard_normal((n,nrhs))
    x=xb.blasWrapper.solve(A,b)
    t0=time.perf_counter(); x=xb.blasWrapper.solve(A,b); dt=time.perf_counter()-t0
    t1=time.perf_counter(); xr=np.linalg.solve(A,b); dn=time.perf_counter()-t1
    print('general n',n,'nrhs',nrhs,'solve %.2f ms (incl. H2D of A)'%(dt*1e3),'numpy %.2f ms'%(dn*1e3),'residual %.1e'%(np.linalg.norm(A@x-b)/np.linalg.norm(b)),'vs numpy %.1e'%(np.linalg.norm(x-xr)/np.linalg.norm(xr)),flush=True)
