import sys; sys.path.insert(0,'.')
import numpy as np, xerus_b200 as xb
xb.init(0)
rng=np.random.default_rng(0)
xb.set_option("svd_wpp",1)
for (m,n) in [(256,256),(64,64),(32,32)]:
    A=rng.standard_normal((m,n)); xb.blasWrapper.svd(A)
