import sys, time; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, xerus_b200 as xb
xb.init(0)
import os
if os.environ.get('XB_GEMM_SMALL'): xb.set_option('gemm_force_small', 1)
if os.environ.get('XB_ALS_PCG'): xb.set_option('als_persistent_cg', int(os.environ['XB_ALS_PCG']))
d,n=16,10
for r in [8,20,50]:
    rng=np.random.default_rng(16)
    A,b=xb.TTOperator.laplace(d,n),xb.TTTensor.ones([n]*d)
    x0=xb.TTTensor.random([n]*d,r,rng)
    for rep in range(2):
        x=x0.copy(); v=xb.ALSVariant(1,0,True)
        xb.synchronize(); t0=time.perf_counter(); l0=xb.kernel_launch_count()
        e=v(A,x,b,2)
        xb.synchronize(); dt=time.perf_counter()-t0
    res=A.apply(x).distance(b)/b.frob_norm()
    print('r',r,'ALS_SPD 1 full sweep: %.1f ms'%(dt*1e3),'energy %.10e'%e,'residual %.2e'%res,'cg its',v.last_local_iterations,'launches',xb.kernel_launch_count()-l0,flush=True)
