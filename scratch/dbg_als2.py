import sys; sys.path.insert(0,'.')
import numpy as np, xerus_b200 as xb
d,n,r=12,10,12
rng=np.random.default_rng(16)
A,b=xb.TTOperator.laplace(d,n),xb.TTTensor.ones([n]*d)
xb.set_option("als_direct_max",0); print("normb",b.frob_norm())
x=xb.TTTensor.random([n]*d,r,rng)
v=xb.ALSVariant(1,0,True); e=v(A,x,b,2)
print(e, A.apply(x).distance(b)/b.frob_norm(), v.last_local_iterations)
