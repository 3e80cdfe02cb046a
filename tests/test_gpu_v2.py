"""Rows a11 / a12 / a19 of SURVEY.md 8 against golden vectors of the unmodified reference (tests/golden/xerus_ref_v2.npz,
oracle/drivers/ref_golden.cpp v2_section): TTNetwork::soft_threshold (ttNetwork.cpp:688-713), the TT-SVD constructor for
TTOperators and with per-bond rank caps (ttNetwork.cpp:112-160), ASD / ASD_SPD (als.cpp:73-103, :562-563)."""
import os

import numpy as np
import pytest

import xerus_b200 as xb
from conftest import ROOT, golden_tt
from oracle import tt_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden2():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "xerus_ref_v2.npz")))


def to_oracle(t):
    return O.TT(t.cores(), core_position=t.corePosition if t.canonicalized else None)


def from_golden(g, name, cls=xb.TTTensor):
    cores, core = golden_tt(g, name)
    return cls.from_cores(cores, core_position=core)


@pytest.mark.parametrize("name", ["scalar", "vector", "core5"])
def test_soft_threshold_golden(golden2, name):
    t = from_golden(golden2, "st.in")
    if name == "scalar":
        t.soft_threshold(2.0e4)
    elif name == "vector":
        t.soft_threshold(list(golden2["st.taus"]))
    else:
        t.move_core(5)
        t.soft_threshold(1.5e6)                  # every spectrum is thresholded to zero: ranks collapse edge by edge (eps = 0 cuts exact zeros)
    ref = O.TT(golden_tt(golden2, "st." + name)[0])
    assert t.ranks() == [int(v) for v in golden2["st.%s.ranks" % name]]
    assert t.canonicalized and t.corePosition == (5 if name == "core5" else 0)
    if name == "core5":
        assert np.linalg.norm(t.to_dense()) == 0.0
    else:
        assert abs(t.frob_norm() - float(golden2["st.%s.norm" % name])) < 1e-11 * t.frob_norm()
        assert O.tt_distance_rel(to_oracle(t), ref) < 1e-9


def test_soft_threshold_arguments():
    t = xb.TTTensor.random([3, 4, 3], 3, np.random.default_rng(1))
    with pytest.raises(xb.XerusError):
        t.soft_threshold([0.1])                  # d-1 taus (ttNetwork.cpp:690)
    before = t.to_dense()
    t.soft_threshold(0.0)                        # tau = 0: the tensor is unchanged
    assert np.linalg.norm(t.to_dense() - before) < 1e-12 * np.linalg.norm(before)
    one = xb.TTTensor.from_cores([np.arange(4.0).reshape(1, 4, 1)])
    one.soft_threshold([])                       # a single component has no edge
    assert np.array_equal(one.to_dense(), np.arange(4.0))


def test_tt_svd_operator_and_rank_caps_golden(golden2):
    full = golden2["opsvd.full"]
    t = xb.TTOperator.from_dense(full, 1e-14)
    assert t.ranks() == [int(v) for v in golden2["opsvd.ranks"]] and t.dimensions == list(full.shape)
    assert t.canonicalized and t.corePosition == 0
    assert np.linalg.norm(t.to_dense() - full) < 1e-12 * np.linalg.norm(full)
    t2 = xb.TTOperator.from_dense(full, 0.0, [4, 7, 3])
    assert t2.ranks() == [int(v) for v in golden2["opsvd.caps.ranks"]]
    assert np.linalg.norm(t2.to_dense() - golden2["opsvd.tt_caps.dense"]) < 1e-10 * np.linalg.norm(full)
    t3 = xb.TTTensor.from_dense(golden2["ttsvd2.full"], 0.0, [2, 5, 6, 3])
    assert t3.ranks() == [int(v) for v in golden2["ttsvd2.caps.ranks"]]
    assert np.linalg.norm(t3.to_dense() - golden2["ttsvd2.tt_caps.dense"]) < 1e-10 * np.linalg.norm(golden2["ttsvd2.full"])
    with pytest.raises(xb.XerusError):
        xb.TTOperator.from_dense(np.zeros((2, 3, 4)))          # odd number of modes (ttNetwork.cpp:113)
    with pytest.raises(xb.XerusError):
        xb.TTTensor.from_dense(golden2["ttsvd2.full"], 0.0, [2, 5])


@pytest.mark.parametrize("tag,d,n", [("asd_small", 6, 4), ("asd_mid", 8, 5)])
@pytest.mark.parametrize("hs", [1, 2, 6])
def test_asd_golden(golden2, tag, d, n, hs):
    A, b = xb.TTOperator.laplace(d, n), xb.TTTensor.ones([n] * d)
    x = from_golden(golden2, tag + ".x0")
    e = xb.ASD_SPD(A, x, b, hs)
    e_ref = float(golden2["%s.spd_hs%d.energy" % (tag, hs)])
    assert abs(e - e_ref) < 1e-9 * abs(e_ref)
    ref = O.TT(golden_tt(golden2, "%s.spd_hs%d.x" % (tag, hs))[0])
    assert x.ranks() == ref.ranks() and x.canonicalized and x.corePosition == 0
    assert O.tt_distance_rel(to_oracle(x), ref) < 1e-9
    if hs <= 2:          # the reference's non-SPD step (ratio of norms, als.cpp:97) diverges on this operator after that
        x = from_golden(golden2, tag + ".x0")
        e = xb.ASD(A, x, b, hs)
        e_ref = float(golden2["%s.gen_hs%d.energy" % (tag, hs)])
        assert abs(e - e_ref) < 1e-8 * abs(e_ref)
        assert O.tt_distance_rel(to_oracle(x), O.TT(golden_tt(golden2, "%s.gen_hs%d.x" % (tag, hs))[0])) < 1e-7


def test_asd_needs_one_site():
    A, b = xb.TTOperator.laplace(4, 3), xb.TTTensor.ones([3] * 4)
    x = xb.TTTensor.random([3] * 4, 2, np.random.default_rng(0))
    with pytest.raises(xb.XerusError):
        xb.ALSVariant(2, 0, True, localSolver="ASD")(A, x, b, 1)      # als.cpp:78
