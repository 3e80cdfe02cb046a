"""Parity of the sweep layer (device-resident TTTensor::round / move_core / TT arithmetic) against the CPU oracle on
seeded inputs and against golden vectors produced by the unmodified reference.  Tolerances (north_star): ranks equal,
reconstructed TT and kept singular values <= 1e-9 relative."""
import numpy as np
import pytest

import xerus_b200 as xb
from conftest import ROOT, golden_tt
from oracle import tt_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def to_oracle(t):
    return O.TT(t.cores(), core_position=t.corePosition if t.canonicalized else None)


def from_golden(g, name, cls=None):
    cores, core = golden_tt(g, name)
    cls = cls or (xb.TTOperator if cores[0].ndim == 4 else xb.TTTensor)
    return cls.from_cores(cores, core_position=core)


def check_gauge(t):
    """canonicalised: cores left of the core position are left-orthonormal, right of it right-orthonormal."""
    assert t.canonicalized
    pos = t.corePosition
    for i, c in enumerate(t.cores()):
        if i < pos:
            m = c.reshape(-1, c.shape[-1]); assert np.linalg.norm(m.T @ m - np.eye(m.shape[1])) < 1e-11
        if i > pos:
            m = c.reshape(c.shape[0], -1); assert np.linalg.norm(m @ m.T - np.eye(m.shape[0])) < 1e-11


# ---- plumbing ----------------------------------------------------------------------------------------------------
def test_component_roundtrip_and_state(golden):
    t = from_golden(golden, "c1.in")
    cores, core = golden_tt(golden, "c1.in")
    assert t.ranks() == [4, 16, 32, 32, 32, 16, 4] and t.dimensions == [4] * 8 and t.degree() == 8
    assert t.canonicalized and t.corePosition == core == 0
    for i, c in enumerate(cores):
        assert np.array_equal(t.get_component(i), c)
    t.set_component(3, cores[3] * 2.0)          # writing a non-core component drops the flag (ttNetwork.cpp:491)
    assert not t.canonicalized
    c = t.copy()
    assert np.array_equal(c.get_component(3), cores[3] * 2.0)
    with pytest.raises(xb.XerusError):
        t.move_core(8)
    with pytest.raises(xb.XerusError):
        t.set_component(2, np.zeros((16, 5, 32)))


def test_norm_inner_dense(golden):
    t = from_golden(golden, "c1.in")
    assert abs(t.frob_norm() - golden["c1.in.norm"]) < 1e-12 * golden["c1.in.norm"]
    r = from_golden(golden, "c1.round16")
    assert abs(t.inner(r) - golden["c1.round16.inner"]) < 1e-12 * golden["c1.round16.inner"]
    o = to_oracle(t)
    assert rel(t.to_dense(), o.to_dense()) < 1e-13
    nc = t.copy(); nc.set_component(0, nc.get_component(0))      # same tensor, flag dropped: norm through <t,t>
    nc.set_component(1, nc.get_component(1))
    assert not nc.canonicalized and abs(nc.frob_norm() - golden["c1.in.norm"]) < 1e-11 * golden["c1.in.norm"]


# ---- move_core ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,pos,keep", [("c1.core5", 5, False), ("c1.core3keep", 3, True)])
def test_move_core_golden(golden, name, pos, keep):
    t = from_golden(golden, "c1.in")
    t.move_core(pos, keep)
    ref, _ = golden_tt(golden, name)
    assert t.corePosition == pos and t.ranks() == [c.shape[-1] for c in ref[:-1]]
    check_gauge(t)
    assert O.tt_distance_rel(to_oracle(t), O.TT(ref)) < 1e-12
    s1 = np.linalg.svd(t.get_component(pos).reshape(t.get_component(pos).shape[0], -1), compute_uv=False)
    s2 = np.linalg.svd(ref[pos].reshape(ref[pos].shape[0], -1), compute_uv=False)
    assert rel(s1, s2) < 1e-11


def test_move_core_non_canonical_and_rank_repair():
    rng = np.random.default_rng(21)
    dims, rk = [3, 4, 2, 5, 3], [1, 3, 7, 6, 3, 1]
    cores = [rng.standard_normal((rk[i], dims[i], rk[i + 1])) for i in range(5)]
    t = xb.TTTensor.from_cores(cores)
    dense = O.TT(cores).to_dense()
    for pos in [0, 2, 4]:
        u = t.copy(); u.move_core(pos)
        check_gauge(u)
        assert rel(u.to_dense(), dense) < 1e-12
    # ranks above the maximal ones are repaired by the extra sweeps (ttNetwork.cpp:609-624)
    big = [rng.standard_normal(s) for s in [(1, 2, 5), (5, 2, 9), (9, 2, 6), (6, 2, 1)]]
    b = xb.TTTensor.from_cores(big)
    b.move_core(1, True)
    o = O.TT(big); o.move_core(1, True)
    assert b.ranks() == o.ranks() == [2, 4, 2]
    assert rel(b.to_dense(), O.TT(big).to_dense()) < 1e-12


# ---- round -------------------------------------------------------------------------------------------------------
def test_round_c1_golden(golden):
    """BASELINE config 1: TTTensor::random({4}x8, 32).round(16)."""
    t = from_golden(golden, "c1.in")
    sv = t.round(16, return_svals=True)
    assert t.ranks() == [int(v) for v in golden["c1.round16.ranks"]] == [4, 16, 16, 16, 16, 16, 4]
    assert t.corePosition == 0
    check_gauge(t)
    ref = O.TT(golden_tt(golden, "c1.round16")[0], core_position=0)
    assert O.tt_distance_rel(to_oracle(t), ref) < 1e-9
    assert abs(t.frob_norm() - golden["c1.round16.norm"]) < 1e-11 * golden["c1.round16.norm"]
    inp = from_golden(golden, "c1.in")
    assert abs(inp.inner(t) - golden["c1.round16.inner"]) < 1e-11 * golden["c1.round16.inner"]
    assert abs(inp.distance(t) / inp.frob_norm() - golden["c1.round16.relerr"]) < 1e-11
    # kept singular values per edge against the oracle's sweep
    o = O.TT(golden_tt(golden, "c1.in")[0], core_position=0)
    osv = o.round(16)                                      # osv[i] belongs to edge d-2-i
    for e in range(7):
        assert rel(sv[e], osv[6 - e]) < 1e-9


def test_round_eps_vector_and_core_restore(golden):
    t = from_golden(golden, "c1.in")
    t.round(0.35)
    assert t.ranks() == [int(v) for v in golden["c1.roundeps.ranks"]]
    assert O.tt_distance_rel(to_oracle(t), O.TT(golden_tt(golden, "c1.roundeps")[0])) < 1e-9
    t = from_golden(golden, "c1.in")
    t.round([3, 9, 20, 32, 11, 7, 2])
    assert t.ranks() == [int(v) for v in golden["c1.roundvec.ranks"]]
    assert O.tt_distance_rel(to_oracle(t), O.TT(golden_tt(golden, "c1.roundvec")[0])) < 1e-9
    t = from_golden(golden, "c1.core5")
    t.round(8)
    assert t.corePosition == 5 and t.ranks() == [int(v) for v in golden["c1.core5.round8.ranks"]]
    check_gauge(t)
    assert O.tt_distance_rel(to_oracle(t), O.TT(golden_tt(golden, "c1.core5.round8")[0])) < 1e-9
    with pytest.raises(xb.XerusError):
        t.round(1.5)
    with pytest.raises(xb.XerusError):
        t.round([1, 2])


def test_round_rank_deficient_sum_and_raw(golden):
    y = from_golden(golden, "sum.y")
    y.round(1e-12)
    assert y.ranks() == [int(v) for v in golden["sum.y.round.ranks"]]
    assert rel(y.to_dense(), golden["sum.y.dense"]) < 1e-11
    raw = from_golden(golden, "raw.in")
    assert not raw.canonicalized
    raw.round(4)
    assert raw.ranks() == [int(v) for v in golden["raw.round4.ranks"]]
    assert rel(raw.to_dense(), golden["raw.round4.dense"]) < 1e-10


def test_no_rounding_is_identity():
    # reference: src/unitTests/ttRounding.cxx:104-117 ("no_rounding")
    rng = np.random.default_rng(4)
    a = xb.TTTensor.random([2] * 7, [2] * 6, rng)
    before = a.to_dense()
    a.round(2)
    assert rel(a.to_dense(), before) < 1e-13
    c = xb.TTTensor.random([2] * 7, [2] * 6, rng)
    s = a + c * 0.0
    s.round(2)
    assert s.ranks() == [2] * 6 and rel(s.to_dense(), before) < 1e-12


@pytest.mark.parametrize("dims,r,target", [([4] * 6, 12, 5), ([2] * 12, 16, 7), ([3, 5, 2, 6, 4], 9, 4), ([10] * 4, 30, 11),
                                           ([2] * 16, 64, 32)])
def test_round_vs_oracle_random(dims, r, target):
    rng = np.random.default_rng(len(dims) * 100 + r)
    t = xb.TTTensor.random(dims, r, rng)
    o = to_oracle(t)
    sv = t.round(target, return_svals=True)
    osv = o.round(target)
    assert t.ranks() == o.ranks()
    d = len(dims)
    for e in range(d - 1):
        assert rel(sv[e], osv[d - 2 - e]) < 1e-9
    assert O.tt_distance_rel(to_oracle(t), o) < 1e-9


def test_round_idempotent_and_optimal_2d():
    rng = np.random.default_rng(8)
    t = xb.TTTensor.random([6] * 5, 20, rng)
    t.round(7)
    a = t.to_dense()
    t.round(7)
    assert rel(t.to_dense(), a) < 1e-12
    # order-2 TT rounding is the best rank-k approximation (Eckart-Young)
    M = rng.standard_normal((40, 30))
    m = xb.TTTensor.from_dense(M, 0.0)
    m.round(5)
    U, S, Vt = np.linalg.svd(M)
    assert rel(m.to_dense(), (U[:, :5] * S[:5]) @ Vt[:5]) < 1e-11


# ---- TT arithmetic that feeds a round (next rows of the scope table) -------------------------------------------------
def test_tt_svd_golden(golden):
    full = golden["ttsvd.full"]
    t = xb.TTTensor.from_dense(full, 1e-14)
    assert t.ranks() == [int(v) for v in golden["ttsvd.ranks"]] and t.corePosition == 0
    assert rel(t.to_dense(), full) < 1e-12
    t3 = xb.TTTensor.from_dense(full, 0.0, 3)
    assert rel(t3.to_dense(), golden["ttsvd.tt3.dense"]) < 1e-10


def test_sum_and_distance(golden):
    x = from_golden(golden, "sum.x")
    y = x + x
    # operator+= re-canonicalises with the rank-revealing move_core (ttNetwork.cpp:842-844).  The reference keeps
    # [3,10,12,3] here only because its rank test is sign dependent (blasLapackWrapper.cpp:269); this library's
    # |.|-rule reveals the true ranks of x + x, which are those of x.
    assert y.ranks() == x.ranks() == [3, 5, 6, 3]
    assert all(a <= b for a, b in zip(y.ranks(), [int(v) for v in golden["sum.y.ranks"]]))
    assert rel(y.to_dense(), golden["sum.y.dense"]) < 1e-13
    assert abs(y.distance(x) - x.frob_norm()) < 1e-11 * x.frob_norm()
    assert x.distance(x) < 1e-13 * x.frob_norm()


def test_operator_apply_then_round(golden):
    A, x = from_golden(golden, "mv.A"), from_golden(golden, "mv.x")
    y = A.apply(x)
    assert y.ranks() == [int(v) for v in golden["mv.y.ranks"]]
    assert rel(y.to_dense(), golden["mv.y.dense"]) < 1e-12
    before = y.copy()
    y.round(8)
    assert y.ranks() == [int(v) for v in golden["mv.y.round8.ranks"]]
    assert O.tt_distance_rel(to_oracle(y), O.TT(golden_tt(golden, "mv.y.round8")[0])) < 1e-9
    assert abs(before.distance(y) / before.frob_norm() - golden["mv.y.round8.relerr"]) < 1e-10
    L = xb.TTOperator.laplace(6, 4)
    for a, b in zip(L.cores(), golden_tt(golden, "mv.A")[0]):
        assert np.array_equal(a, b)
    assert rel(L.to_dense(), O.laplace_operator(6, 4).to_dense()) < 1e-14


def test_round_batched():
    rng = np.random.default_rng(12)
    tts = [xb.TTTensor.random([4] * 5, 10, rng) for _ in range(6)]
    refs = [to_oracle(t) for t in tts]
    xb.round_batched(tts, 4)
    for t, o in zip(tts, refs):
        o.round(4)
        assert t.ranks() == o.ranks() and O.tt_distance_rel(to_oracle(t), o) < 1e-9


def test_batch_sharding_single_rank():
    """BASELINE config 5 (reduced): y_b = A x_b ; y_b.round(r) per item through the sharding helpers, device path."""
    from xerus_b200 import parallel
    d, n, r, items = 6, 4, 6, 5
    A = xb.TTOperator.laplace(d, n)
    Ao = O.laplace_operator(d, n)
    make = lambda b: xb.TTTensor.random([n] * d, r, parallel.item_rng(99, b))
    res = parallel.gather_by_item(parallel.matvec_round_batch(A, make, items, r), items)
    for b, (ranks, nrm) in enumerate(res):
        xo = to_oracle(make(b))
        yo = O.tt_apply(Ao, xo); yo.round(r)
        assert list(ranks) == yo.ranks() and abs(nrm - yo.frob_norm()) < 1e-10 * yo.frob_norm()
    # two "ranks" processed one after the other cover every item exactly once and give the same answers
    parts = {}
    for rank in range(2):
        parts.update(parallel.matvec_round_batch(A, make, items, r, rank, 2))
    assert [parts[b] for b in range(items)] == res


def test_batch_on_multiple_workers_matches_sequential():
    """Several host threads, each on its own library worker (CUDA stream), give bit-identical per-item results."""
    from xerus_b200 import parallel
    d, n, r, items = 6, 4, 6, 12
    A = xb.TTOperator.laplace(d, n)
    make = lambda b: xb.TTTensor.random([n] * d, r, parallel.item_rng(7, b))
    seq = parallel.matvec_round_batch(A, make, items, r)
    par = parallel.matvec_round_batch(A, make, items, r, workers=4)
    assert seq == par


# ---- edge cases the reference tests exercise (ttRounding.cxx:27-102: orders 1 and 2, modes of size 1, ragged dims) -----------
@pytest.mark.parametrize("dims", [(2,), (2, 2), (2, 7), (5, 6, 3, 1, 4, 2, 8, 1), (1, 1, 1), (3, 1, 1, 4)])
def test_round_trip_small_and_degenerate(dims):
    rng = np.random.default_rng(sum(dims))
    A = rng.standard_normal(dims)
    t = xb.TTTensor.from_dense(A, 1e-14)
    assert rel(t.to_dense(), A) < 1e-14                      # ttRounding.cxx: approx_equal(B, A, 1e-14)
    t.round(1e-14)
    assert rel(t.to_dense(), A) < 1e-14
    t.round(int(np.prod(dims)))
    assert rel(t.to_dense(), A) < 1e-14
    for pos in range(len(dims)):
        t.move_core(pos)
        assert t.corePosition == pos and rel(t.to_dense(), A) < 1e-13


def test_rank_one_and_zero_tensors():
    ones = xb.TTTensor.ones([3, 4, 5])
    assert ones.ranks() == [1, 1] and abs(ones.frob_norm() - np.sqrt(60)) < 1e-13
    ones.round(5)
    assert ones.ranks() == [1, 1] and rel(ones.to_dense(), np.ones((3, 4, 5))) < 1e-14
    z = xb.TTTensor.from_cores([np.zeros((1, 3, 2)), np.zeros((2, 3, 1))])
    z.round(1)                                               # rank never drops to 0 (tensor.cpp:1464-1474)
    assert z.ranks() == [1] and np.all(z.to_dense() == 0)


def test_round_large_magnitudes():
    """Random TTs are not normalised (norm ~1e33 at config 3): everything is relative."""
    rng = np.random.default_rng(77)
    t = xb.TTTensor.random([2] * 10, 16, rng)
    t *= 1e120                       # (the numpy oracle forms squares of the norm: stay below 1e150)
    o = to_oracle(t)
    t.round(5); o.round(5)
    assert t.ranks() == o.ranks() and O.tt_distance_rel(to_oracle(t), o) < 1e-9


@pytest.mark.parametrize("option", ["svd_colsort", "svd_dsmem", "qr_defer"])
def test_round_sweep_variants_agree(option):
    """round() with the sweep-level optimisations switched off gives the same tensor: sorted-column pre-conditioning, DSMEM
    hand-over in the Jacobi kernel, Q formed on the side stream.  Degree 14, n = 2, rank 64 -> 32: interior SVDs of 64 columns
    (split kernel, clusters of 4) and first edges with column norms spread over decades."""
    rng = np.random.default_rng(8)
    base = xb.TTTensor.random([2] * 14, 64, rng)
    out = []
    try:
        for v in (0, 1):
            xb.set_option(option, v)
            t = base.copy()
            t.round(32)
            out.append(t)
    finally:
        xb.set_option(option, 1)
    assert out[0].ranks() == out[1].ranks()
    assert out[0].distance(out[1]) < 1e-12 * out[1].frob_norm()
    ref = O.TT(base.cores(), core_position=0)
    ref.round(32)
    assert O.tt_distance_rel(O.TT(out[1].cores(), core_position=out[1].corePosition), ref) < 1e-9


def test_components_in_one_call_roundtrip():
    """from_cores / cores() move all components with one synchronisation (xb_tt_set_components / xb_tt_get_components)."""
    rng = np.random.default_rng(2)
    cores = [rng.standard_normal(s) for s in [(1, 3, 5), (5, 2, 7), (7, 4, 2), (2, 3, 1)]]
    t = xb.TTTensor.from_cores(cores)
    assert t.ranks() == [5, 7, 2] and not t.canonicalized
    for a, b in zip(t.cores(), cores):
        assert np.array_equal(a, b)
    for i, c in enumerate(cores):
        assert np.array_equal(t.get_component(i), c)
    ops = [rng.standard_normal(s) for s in [(1, 2, 3, 4), (4, 3, 2, 1)]]
    A = xb.TTOperator.from_cores(ops)
    assert A.ranks() == [4] and all(np.array_equal(a, b) for a, b in zip(A.cores(), ops))
    with pytest.raises(xb.XerusError):
        xb.TTTensor.from_cores([np.zeros((2, 3, 1))])


# ---- round plans (captured CUDA graph of a whole sweep, speculative ranks) ---------------------------------------------
def test_round_plan_replay_is_bit_identical_and_falls_back():
    """The third round() of a shape replays the graph captured at the second: same kernels in the same order on the same data,
    so the result is bit for bit the one of the ordinary path; data that break the speculation (a rank-deficient TT of the same
    shape) raise the device flag and take the ordinary path."""
    rng = np.random.default_rng(77)
    dims, r = [3, 4, 5, 4, 3], 7
    inputs = [xb.TTTensor.random(dims, r, rng) for _ in range(4)]
    assert all(t.ranks() == inputs[0].ranks() for t in inputs)
    xb.set_option("round_plans", 0)
    plain = []
    for t in inputs:
        c = t.copy(); c.round(4); plain.append(c.cores())
    xb.set_option("round_plans", 1)
    l0 = xb.kernel_launch_count()
    for t, ref in zip(inputs, plain):
        c = t.copy(); c.round(4)
        assert c.ranks() == [3, 4, 4, 3] and c.canonicalized and c.corePosition == 0
        for a, b in zip(c.cores(), ref):
            assert np.array_equal(a, b)
    assert xb.kernel_launch_count() > l0
    # same shape, but numerically rank deficient at bond 2: the reference (and the ordinary path) cut that bond to 2
    low = xb.TTTensor.random(dims, [3, 2, 7, 3], rng)
    cores = low.cores()
    pad = [np.zeros(s.shape) for s in inputs[0].cores()]
    for p, c in zip(pad, cores):
        p[:c.shape[0], :, :c.shape[2]] = c
    deficient = xb.TTTensor.from_cores(pad, core_position=None)
    deficient.move_core(0, True)                       # keepRank: same ranks and canonical form as the planned shape
    assert deficient.ranks() == inputs[0].ranks()
    ref = to_oracle(deficient); ref.round(4)
    deficient.round(4)
    assert deficient.ranks() == ref.ranks()
    assert O.tt_distance_rel(to_oracle(deficient), ref) < 1e-9


def test_batched_items_match_the_single_calls():
    """xb_tt_apply_round_batched / xb_tt_round_batched (library threads + streams, round plans) against one call per item."""
    rng = np.random.default_rng(55)
    d, n, r = 6, 4, 12
    A = xb.TTOperator.laplace(d, n)
    xs = [xb.TTTensor.random([n] * d, r, rng) for _ in range(13)]
    single = []
    for x in xs:
        y = A.apply(x); y.round(r); single.append(y)
    ys = xb.apply_round_batched(A, xs, r)
    assert len(ys) == len(xs)
    for y, ref in zip(ys, single):
        assert y.ranks() == ref.ranks() and y.canonicalized and y.corePosition == 0
        assert y.distance(ref) < 1e-12 * ref.frob_norm()
    raw = [A.apply(x) for x in xs]
    xb.round_batched(raw, r)
    for y, ref in zip(raw, single):
        assert y.ranks() == ref.ranks() and y.distance(ref) < 1e-12 * ref.frob_norm()
    assert xb.apply_round_batched(A, [], r) == []
    with pytest.raises(xb.XerusError):
        xb.apply_round_batched(A, [xb.TTTensor.ones([n] * (d - 1))], r)


def test_shutdown_releases_every_worker_and_reinit_works():
    """xb_shutdown: streams, events, scratch, memory pools and round plans of every worker go; a second xb_init starts
    from a clean slate.  Runs in a process of its own (the handles of this process must not outlive a shutdown)."""
    import subprocess
    import sys
    code = r"""
import sys, threading, numpy as np
sys.path.insert(0, %r)
import xerus_b200 as xb
from xerus_b200._lib import call
rng = np.random.default_rng(3)
def work():
    ts = [xb.TTTensor.random([3, 4, 3, 4], 5, rng) for _ in range(6)]
    for t in ts[:3]:
        t.round(2)                                  # ordinary, capture, replay on worker 0
    xb.round_batched(ts[3:], 2)                     # batch workers
    def other():
        xb.worker_select(3)
        u = xb.TTTensor.random([3, 3, 3], 2, np.random.default_rng(1)); u.move_core(2); xb.synchronize()
    th = threading.Thread(target=other); th.start(); th.join()
    xb.synchronize_all()
    return [t.ranks() for t in ts]
xb.init(0)
a = work()
del rng
call("xb_shutdown")
rng = np.random.default_rng(3)
xb.init(0)
b = work()
assert a == b, (a, b)
print("reinit ok", a[0])
""" % ROOT
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "reinit ok" in p.stdout, p.stdout + p.stderr
