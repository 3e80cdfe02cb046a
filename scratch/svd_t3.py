import sys; sys.path.insert(0,'.')
import numpy as np, xerus_b200 as xb
xb.init(0)
rng=np.random.default_rng(0)
A=rng.standard_normal((256,256))
for (bw,wpp) in [(8,1),(4,1),(2,1),(4,2)]:
    xb.set_option("svd_max_bw",bw); xb.set_option("svd_wpp",wpp)
    xb.blasWrapper.svd(A)
    xb.profile_enable(True)
    for _ in range(3): xb.blasWrapper.svd(A)
    sc,l,ms=xb.profile_get("svd_jacobi"); xb.profile_enable(False)
    print('bw',bw,'wpp',wpp,'jacobi ms/svd %.3f'%(ms/sc),flush=True)
