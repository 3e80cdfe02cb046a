import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import xerus_b200 as xb
from conftest import golden_tt
from oracle import tt_oracle as O
g = dict(np.load('tests/golden/xerus_ref_v1.npz'))
def fg(name):
    cores, core = golden_tt(g, name)
    cls = xb.TTOperator if cores[0].ndim == 4 else xb.TTTensor
    return cls.from_cores(cores, core_position=core)
def to_o(t): return O.TT(t.cores(), core_position=t.corePosition if t.canonicalized else None)
A, b = fg("als_mid.A"), fg("als_mid.b")
Ao, bo = O.TT(golden_tt(g, "als_mid.A")[0]), O.TT(golden_tt(g, "als_mid.b")[0])
print("== DMRG energies direct")
for hs in [1,2,3,4]:
    x = fg("als_mid.x0"); e = xb.DMRG_SPD(A, x, b, hs)
    xo = O.TT(golden_tt(g, "als_mid.x0")[0], core_position=0)
    eo = O.ALSVariant(2, True, fix_dmrg_turn=True)(Ao, xo, bo, hs)
    print(hs, e, eo, x.ranks(), xo.ranks(), O.residual(Ao, to_o(x), bo), O.residual(Ao, xo, bo))
print("== DMRG CG")
xb.set_option("als_direct_max", 0)
x = fg("als_mid.x0"); v = xb.ALSVariant(2,0,True); e = v(A, x, b, 1)
print(e, float(g["als_mid.dmrg_hs1.energy"]), x.ranks(), g["als_mid.dmrg_hs1.ranks"], v.last_local_iterations)
print("== config2 reduced")
for (d,n,r) in [(8,5,6),(10,6,8),(12,10,12),(16,10,20)]:
    rng = np.random.default_rng(16)
    A2, b2 = xb.TTOperator.laplace(d, n), xb.TTTensor.ones([n]*d)
    x0 = xb.TTTensor.random([n]*d, r, rng)
    for hs in [1,2]:
        x = x0.copy(); v = xb.ALSVariant(1,0,True); e = v(A2, x, b2, hs)
        res = A2.apply(x).distance(b2)/b2.frob_norm()
        print(d,n,r,'hs',hs,'energy',e,'res',res,'cg its',v.last_local_iterations, 'ranks', x.ranks()[:4])
