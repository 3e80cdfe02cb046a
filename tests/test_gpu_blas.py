"""Parity of the per-call C-ABI layer (the xerus::blasWrapper mirror) against the CPU oracle and against golden
vectors from the unmodified reference.  Tolerances: contractions rel-Frobenius <= 1e-10 (north_star); singular values
<= 1e-9 relative; Q/U/V factors are compared through invariants (sign/rotation ambiguity), as the reference's own
tests do (src/unitTests/fullTensor_factorisations.cxx:26-276)."""
import numpy as np
import pytest

import xerus_b200 as xb
from oracle import tt_oracle as O

pytestmark = pytest.mark.gpu
BW = xb.blasWrapper


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


# ---- contractions ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case,ta,tb", [("nn", False, False), ("tn", True, False), ("nt", False, True), ("tt", True, True)])
def test_gemm_golden(golden, case, ta, tb):
    A = golden["mm.At"] if ta else golden["mm.A"]
    B = golden["mm.Bt"] if tb else golden["mm.B"]
    assert rel(BW.matrix_matrix_product(1.5, A, ta, B, tb), golden["mm.C_" + case]) < 1e-13


def test_gemm_degenerate_golden(golden):
    g = golden
    assert rel(BW.matrix_matrix_product(2.0, g["mm.row1"], False, g["mm.B"], False), g["mm.C_left1"]) < 1e-13
    assert rel(BW.matrix_matrix_product(2.0, g["mm.A"], False, g["mm.col1"], False), g["mm.C_right1"]) < 1e-13
    assert rel(BW.matrix_matrix_product(2.0, g["mm.colm"], False, g["mm.rown"], False), g["mm.C_mid1"]) < 1e-13


def test_known_answer_products():
    # reference: src/unitTests/fullTensor_product.cxx:30-97 — integer known answers, bit exact
    A = np.array([[1.0, 2.0], [3.0, 4.0]])
    B = np.array([[5.0, 6.0], [7.0, 8.0]])
    assert np.array_equal(xb.contract(A, False, B, False, 1), [[19, 22], [43, 50]])
    assert np.array_equal(xb.contract(A, False, B, True, 1), [[17, 23], [39, 53]])
    assert np.array_equal(xb.contract(A, True, B, False, 1), [[26, 30], [38, 44]])
    assert np.array_equal(xb.contract(A, True, B, True, 1), [[23, 31], [34, 46]])


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (3, 5, 7), (64, 64, 16), (65, 63, 17), (128, 32, 32), (256, 256, 256),
                                   (512, 128, 256), (100, 1, 37), (1, 90, 41), (33, 47, 1), (200, 300, 2), (513, 257, 129)])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_vs_oracle(m, n, k, ta, tb):
    rng = np.random.default_rng(m * 1000003 + n * 1009 + k)
    A = rng.standard_normal((k, m) if ta else (m, k))
    B = rng.standard_normal((n, k) if tb else (k, n))
    assert rel(BW.matrix_matrix_product(-0.75, A, ta, B, tb), O.matrix_matrix_product(-0.75, A, ta, B, tb)) < 1e-10


def test_product_1000x1000_consistency():
    # reference: fullTensor_product.cxx:400-418 ("Product_1000x1000"): N/T variants of the same product agree
    rng = np.random.default_rng(7)
    A, B = rng.standard_normal((1000, 1000)), rng.standard_normal((1000, 1000))
    C = BW.matrix_matrix_product(1.0, A, False, B, False)
    assert rel(C, A @ B) < 1e-10
    At, Bt = np.ascontiguousarray(A.T), np.ascontiguousarray(B.T)
    assert np.array_equal(C, BW.matrix_matrix_product(1.0, At, True, B, False))
    assert np.array_equal(C, BW.matrix_matrix_product(1.0, A, False, Bt, True))
    assert np.array_equal(C, BW.matrix_matrix_product(1.0, At, True, Bt, True))


# ---- large-tile (128 x 128, cp.async) GEMM kernel: taken from about a wave of tiles on (csrc/gemm_f64.cu) -------------------
@pytest.mark.parametrize("m,n,k", [(1536, 1536, 256), (1500, 1302, 130), (1024, 2048, 64), (2050, 1026, 272)])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_large_tile_kernel(m, n, k, ta, tb):
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    A = rng.standard_normal((k, m) if ta else (m, k))
    B = rng.standard_normal((n, k) if tb else (k, n))
    try:
        xb.set_option("gemm_big", 0)
        C_small = BW.matrix_matrix_product(-0.75, A, ta, B, tb)
    finally:
        xb.set_option("gemm_big", 1)
    C_big = BW.matrix_matrix_product(-0.75, A, ta, B, tb)
    assert rel(C_big, O.matrix_matrix_product(-0.75, A, ta, B, tb)) < 1e-10
    # same summation order in both kernels: the tile size never changes the bits
    assert np.array_equal(C_big, C_small)


def test_gemm_large_tile_beta_and_odd_shapes():
    # device layer, beta != 0 (the trailing updates of the blocked factorizations), odd ldc (scalar epilogue), and an odd
    # extent that must fall back to the 64 x 64 kernel
    import torch
    from xerus_b200._lib import call
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5)
    for (m, n, k, ldc) in [(1536, 1408, 192, 1408), (1536, 1407, 192, 1407), (1537, 1408, 190, 1410)]:
        A = torch.randn(m, k, dtype=torch.float64, generator=g)
        B = torch.randn(k, n, dtype=torch.float64, generator=g)
        C0 = torch.randn(m, ldc, dtype=torch.float64, generator=g)
        dA, dB, dC = A.to(dev), B.to(dev), C0.to(dev)
        torch.cuda.synchronize()
        call("xb_dev_gemm", dC.data_ptr(), ldc, m, n, 0.5, dA.data_ptr(), k, 0, k, dB.data_ptr(), n, 0, -2.0)
        xb.synchronize()
        want = C0.clone()
        want[:, :n] = 0.5 * (A @ B) - 2.0 * C0[:, :n]
        assert rel(dC.cpu().numpy(), want.numpy()) < 1e-13


def test_prefetch_pins_host_operands_in_place():
    """xb_prefetch / xb_release (memory hooks): an existing host array is pinned where it lies; results do not change."""
    from xerus_b200._lib import call
    rng = np.random.default_rng(4)
    A, B = rng.standard_normal((300, 200)), rng.standard_normal((200, 150))
    want = BW.matrix_matrix_product(1.0, A, False, B, False)
    call("xb_prefetch", A.ctypes.data, A.nbytes)
    call("xb_prefetch", A.ctypes.data, A.nbytes)          # idempotent
    try:
        assert np.array_equal(BW.matrix_matrix_product(1.0, A, False, B, False), want)
    finally:
        call("xb_release", A.ctypes.data)
        call("xb_release", A.ctypes.data)                 # idempotent
    assert np.array_equal(BW.matrix_matrix_product(1.0, A, False, B, False), want)


def test_gemv_ger_level1(golden):
    rng = np.random.default_rng(3)
    A, y = rng.standard_normal((37, 23)), rng.standard_normal(23)
    assert rel(BW.matrix_vector_product(1.25, A, False, y), 1.25 * A @ y) < 1e-13
    At = np.ascontiguousarray(A.T)
    assert rel(BW.matrix_vector_product(1.25, At, True, y), 1.25 * A @ y) < 1e-13
    x = rng.standard_normal(37)
    assert rel(BW.dyadic_vector_product(-2.0, x, y), -2.0 * np.outer(x, y)) < 1e-14
    gx, gy = golden["l1.x"], golden["l1.y"]
    assert abs(BW.one_norm(gx) - golden["l1.one_norm"]) < 1e-12 * golden["l1.one_norm"]
    assert abs(BW.two_norm(gx) - golden["l1.two_norm"]) < 1e-13 * golden["l1.two_norm"]
    assert abs(BW.dot_product(gx, gy) - golden["l1.dot"]) < 1e-11
    assert BW.two_norm(np.zeros(0)) == 0.0


@pytest.mark.parametrize("case,lt,rt", [("nn", False, False), ("nt", False, True), ("tn", True, False), ("tt", True, True)])
def test_contract_golden(golden, case, lt, rt):
    A = golden["ct.At"] if lt else golden["ct.A"]
    B = golden["ct.Bt"] if rt else golden["ct.B"]
    C = xb.contract(A, lt, B, rt, 2)
    assert C.shape == golden["ct.C_" + case].shape and rel(C, golden["ct.C_" + case]) < 1e-13


@pytest.mark.parametrize("p", range(8))
def test_reshuffle_golden(golden, p):
    perm = [int(v) for v in golden["rs.perm%d" % p]]
    out = xb.reshuffle(golden["ct.A"], perm)
    assert out.shape == golden["rs.out%d" % p].shape and np.array_equal(out, golden["rs.out%d" % p])   # bit exact


@pytest.mark.parametrize("shape,perm", [((7, 1, 5), (2, 1, 0)), ((2, 3, 4, 5, 6), (4, 0, 3, 1, 2)), ((64, 50), (1, 0)),
                                        ((33, 2, 65), (0, 2, 1)), ((1, 1, 1), (2, 0, 1)), ((5,), (0,)), ((2, 2, 2, 2, 2, 2, 2, 2), (7, 6, 5, 4, 3, 2, 1, 0))])
def test_reshuffle_vs_oracle(shape, perm):
    rng = np.random.default_rng(11)
    t = rng.standard_normal(shape)
    assert np.array_equal(xb.reshuffle(t, perm), O.reshuffle(t, perm))


def test_index_notation_example(golden):
    # README.md:13  A(i,j) = B(i,k,l) * C(k,j,l): one reshuffle + one GEMM
    Cs = xb.reshuffle(golden["idx.C"], [0, 2, 1])
    assert rel(xb.contract(golden["idx.B"], False, Cs, False, 2), golden["idx.A"]) < 1e-13


# ---- factorizations --------------------------------------------------------------------------------------------
# (127|128|129, ...) and (2048|2049, ...) straddle the row range of the cluster panel kernel, 1030 / 1500 / 2048 its larger
# register tiles, (600, 33) a one-column last panel, (3000, 40) the one-CTA kernel working in global memory
QR_SHAPES = [(1, 1), (5, 1), (1, 5), (8, 8), (40, 12), (12, 40), (33, 32), (64, 64), (100, 37), (37, 100), (512, 256),
             (256, 256), (500, 50), (1000, 70), (130, 129), (127, 20), (128, 64), (129, 129), (600, 33), (1030, 40),
             (1500, 33), (2048, 48), (2049, 16), (3000, 40), (70, 300)]


@pytest.mark.parametrize("m,n", QR_SHAPES)
def test_qr_invariants(m, n):
    rng = np.random.default_rng(m * 131 + n)
    A = rng.standard_normal((m, n))
    Q, R = BW.qr(A)
    k = min(m, n)
    assert Q.shape == (m, k) and R.shape == (k, n)
    assert rel(Q @ R, A) < 1e-13
    assert np.linalg.norm(Q.T @ Q - np.eye(k)) < 1e-12
    assert np.array_equal(np.tril(R, -1), np.zeros_like(R))
    Qo, Ro = O.qr(A)                                   # unique up to row signs of R for full-rank A
    s = np.sign(np.diag(R[:, :k])) * np.sign(np.diag(Ro[:, :k]))
    assert rel(s[:, None] * R, Ro) < 1e-10


@pytest.mark.parametrize("m,n", QR_SHAPES)
def test_rq_invariants(m, n):
    rng = np.random.default_rng(m * 137 + n)
    A = rng.standard_normal((m, n))
    R, Q = BW.rq(A)
    k = min(m, n)
    assert R.shape == (m, k) and Q.shape == (k, n)
    assert rel(R @ Q, A) < 1e-13
    assert np.linalg.norm(Q @ Q.T - np.eye(k)) < 1e-12
    Ro, Qo = O.rq(A)                                   # LAPACK convention: upper trapezoid aligned bottom-right
    assert np.allclose(np.abs(R), np.abs(Ro), rtol=1e-9, atol=1e-11 * np.abs(Ro).max())


@pytest.mark.parametrize("tag", ["tall", "wide"])
def test_qr_rq_golden(golden, tag):
    A = golden["qr.%s.A" % tag]
    Q, R = BW.qr(A)
    s = np.sign(np.sum(Q * golden["qr.%s.Q" % tag], axis=0))
    assert rel(Q * s, golden["qr.%s.Q" % tag]) < 1e-11 and rel(s[:, None] * R, golden["qr.%s.R" % tag]) < 1e-11
    R2, Q2 = BW.rq(A)
    s = np.sign(np.sum(Q2 * golden["rq.%s.Q" % tag], axis=1))
    assert rel(s[:, None] * Q2, golden["rq.%s.Q" % tag]) < 1e-11 and rel(R2 * s, golden["rq.%s.R" % tag]) < 1e-11


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_qc_cq_golden(golden, idx):
    """Rank detection on the reference's own inputs.  The reference's answer for +D (idx 1) is 20 only because of its
    signed-threshold quirk (blasLapackWrapper.cpp:269); the true rank of both +D and -D is 7, which is what this
    library (and the oracle with signed_quirk=False) returns."""
    A = golden["qc%d.A" % idx]
    true_rank = 20 if idx == 0 else 7
    Q, Cm, r = BW.qc(A)
    assert r == true_rank == O.qc(A, signed_quirk=False)[2]
    assert rel(Q @ Cm, A) < 1e-12 and np.linalg.norm(Q.T @ Q - np.eye(r)) < 1e-12
    C2, Q2, r2 = BW.cq(A)
    assert r2 == true_rank
    assert rel(C2 @ Q2, A) < 1e-12 and np.linalg.norm(Q2 @ Q2.T - np.eye(r2)) < 1e-12
    if idx == 2:   # the case where the reference itself reduces the rank: same column space
        Qref = golden["qc2.Q"]
        assert np.linalg.norm(Qref - Q @ (Q.T @ Qref)) < 1e-10


# min(m, n) = 7 .. 512 walks through every instantiation of the specialised Jacobi kernel (64, 128, 192, 256, 384, 512
# elements per row part) and its block widths; (1100, 600) takes the generic kernel
SVD_SHAPES = [(1, 1), (4, 4), (33, 21), (21, 33), (32, 32), (64, 64), (100, 30), (30, 100), (128, 128), (256, 256),
              (512, 256), (256, 512), (300, 7), (7, 300), (257, 129), (190, 180), (400, 300), (330, 520), (512, 512),
              (1100, 600)]


@pytest.mark.parametrize("m,n", SVD_SHAPES)
def test_svd_vs_oracle(m, n):
    rng = np.random.default_rng(m * 139 + n)
    A = rng.standard_normal((m, n))
    U, S, Vt = BW.svd(A)
    k = min(m, n)
    assert U.shape == (m, k) and S.shape == (k,) and Vt.shape == (k, n)
    So = np.linalg.svd(A, compute_uv=False)
    assert np.max(np.abs(S - So)) < 1e-9 * So[0] * 1e-2          # truncated singular values <= 1e-9 relative (north_star)
    assert np.all(np.diff(S) <= 0)                                # descending
    assert rel((U * S) @ Vt, A) < 1e-12
    assert np.linalg.norm(U.T @ U - np.eye(k)) < 1e-11 and np.linalg.norm(Vt @ Vt.T - np.eye(k)) < 1e-11


VARIANT_DEFAULTS = {"svd_fast": 1, "svd_jacc": 1, "svd_recursive": 1, "svd_flip": 1, "qr_cluster": 1, "svd_gram": 0, "svd_split": 1,
                    "svd_dsmem": 1, "svd_colsort": 1, "qr_defer": 1}


def _variant_matrix(shape):
    rng = np.random.default_rng(11)
    if shape == "graded_150":
        return rng.standard_normal((200, 260)) @ np.diag(np.logspace(0, -6, 260)) @ rng.standard_normal((260, 150))
    # 150 columns = 19 blocks (generic kernel); 128 / 256 columns = power-of-two block counts: the split kernel with X and V
    # workers and the DSMEM hand-over (clusters of 8 / 16).  Columns scaled over eight decades in random order: the case
    # the norm-sorted pre-conditioning is there for.
    n = 128 if shape == "unsorted_cols_128" else 256
    return rng.standard_normal((300, n)) * 10.0 ** rng.uniform(-8, 0, n)


@pytest.mark.parametrize("shape", ["graded_150", "unsorted_cols_128", "unsorted_cols_256"])
@pytest.mark.parametrize("option", sorted(VARIANT_DEFAULTS))
def test_factorization_kernel_variants_agree(option, shape):
    """Every optimisation of the factorisation kernels can be switched off; both settings must give the same factors."""
    A = _variant_matrix(shape)
    res = []
    try:
        for v in (0, 1):
            xb.set_option(option, v)
            U, S, Vt = BW.svd(A)
            Q, R = BW.qr(A)
            k = A.shape[1]
            assert rel((U * S) @ Vt, A) < 1e-12 and rel(Q @ R, A) < 1e-13
            assert np.linalg.norm(U.T @ U - np.eye(k)) < 1e-11 and np.linalg.norm(Q.T @ Q - np.eye(k)) < 1e-12
            assert np.linalg.norm(Vt @ Vt.T - np.eye(k)) < 1e-11
            res.append((S, np.abs(R)))
    finally:
        xb.set_option(option, VARIANT_DEFAULTS[option])
    assert np.max(np.abs(res[0][0] - res[1][0])) < 1e-13 * res[0][0][0]
    assert np.allclose(res[0][1], res[1][1], rtol=1e-9, atol=1e-12 * res[0][1].max())


def test_svd_unsorted_column_scales_need_few_sweeps():
    """The matrices a TT sweep hands to the SVD have column norms spread over many decades in no particular order (products of
    cores at the first edges, [Q1 W, Q2 W] later).  With the columns sorted by norm before the QR pre-conditioning the Jacobi
    iteration needs a third of the sweeps, and the singular values keep their relative accuracy."""
    import ctypes as C
    import torch
    from xerus_b200._lib import call
    rng = np.random.default_rng(3)
    n = 256
    A = rng.standard_normal((n, n)) * 10.0 ** rng.uniform(-9, 0, n)
    dA = torch.from_numpy(A).cuda()
    U, Vt = (torch.empty(n, n, dtype=torch.float64, device="cuda") for _ in range(2))
    S = torch.empty(n, dtype=torch.float64, device="cuda")
    sweeps = {}
    try:
        for v in (0, 1):
            xb.set_option("svd_colsort", v)
            sw = C.c_int()
            call("xb_dev_svd", U.data_ptr(), S.data_ptr(), Vt.data_ptr(), dA.data_ptr(), n, n, n, 0, 0, C.byref(sw))
            xb.synchronize()
            sweeps[v] = sw.value
            s_ref = np.linalg.svd(A, compute_uv=False)
            assert np.max(np.abs(S.cpu().numpy() - s_ref) / s_ref[0]) < 1e-13
            assert rel((U.cpu().numpy() * S.cpu().numpy()) @ Vt.cpu().numpy(), A) < 1e-12
    finally:
        xb.set_option("svd_colsort", 1)
    assert sweeps[1] + 4 <= sweeps[0], sweeps


@pytest.mark.parametrize("tag", ["svd.tall", "svd.wide"])
def test_svd_golden(golden, tag):
    A = golden[tag + ".A"]
    U, S, Vt = BW.svd(A)
    assert rel(S, golden[tag + ".S"]) < 1e-12
    s = np.sign(np.sum(U * golden[tag + ".U"], axis=0))
    assert rel(U * s, golden[tag + ".U"]) < 1e-9 and rel(s[:, None] * Vt, golden[tag + ".Vt"]) < 1e-9


def test_svd_graded_and_rank_deficient():
    rng = np.random.default_rng(5)
    Qa, _ = np.linalg.qr(rng.standard_normal((96, 96)))
    Qb, _ = np.linalg.qr(rng.standard_normal((96, 96)))
    sig = np.logspace(0, -14, 96)
    A = (Qa * sig) @ Qb.T
    _, S, _ = BW.svd(A)
    assert np.max(np.abs(S - sig)) < 1e-12           # absolute accuracy eps * sigma_0; Jacobi does better on graded spectra
    B = rng.standard_normal((80, 9)) @ rng.standard_normal((9, 60))
    U, S, Vt = BW.svd(B)
    assert S[9] < 1e-12 * S[0] and rel((U * S) @ Vt, B) < 1e-12
    Z = np.zeros((12, 7))
    U, S, Vt = BW.svd(Z)
    assert np.all(S == 0)


def test_truncation_rule_golden(golden):
    A = golden["tsvd.A"]
    assert len(xb.calculate_svd(A, 1, 0, xb.EPSILON)[1]) == int(golden["tsvd.rank_eps"])
    U, S, Vt = xb.calculate_svd(A, 1, 3, xb.EPSILON)
    assert len(S) == 3 and rel(S, np.diag(golden["tsvd.S3"])) < 1e-12
    assert rel((U * S) @ Vt, (golden["tsvd.U3"] @ golden["tsvd.S3"]) @ golden["tsvd.Vt3"]) < 1e-11
    assert len(xb.calculate_svd(A, 1, 0, 0.5)[1]) == int(golden["tsvd.rank_eps05"])
    with pytest.raises(xb.XerusError):
        xb.calculate_svd(A, 1, 0, 1.5)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_solve_golden(golden, idx):
    x = BW.solve(golden["solve%d.A" % idx], golden["solve.rhs"])
    assert rel(x, golden["solve%d.x" % idx]) < 1e-9


def test_solve_multi_rhs_and_least_squares():
    rng = np.random.default_rng(9)
    G = rng.standard_normal((60, 60))
    spd = G.T @ G + np.eye(60)
    B = rng.standard_normal((60, 4))
    assert rel(BW.solve(spd, B), np.linalg.solve(spd, B)) < 1e-10       # (the reference is defective for nrhs > 1)
    assert rel(BW.solve(G, B), np.linalg.solve(G, B)) < 1e-9
    T = rng.standard_normal((50, 20))
    b = rng.standard_normal((50, 2))
    assert rel(BW.solve_least_squares(T, b), np.linalg.lstsq(T, b, rcond=None)[0]) < 1e-10
    W = rng.standard_normal((20, 50))
    b2 = rng.standard_normal((20, 1))
    assert rel(BW.solve(W, b2), np.linalg.lstsq(W, b2, rcond=None)[0]) < 1e-10   # m != n -> least squares (:553-559)


@pytest.mark.parametrize("n,nrhs", [(64, 1), (65, 2), (130, 1), (200, 5), (700, 1), (1000, 3)])
def test_solve_blocked_factorizations(n, nrhs):
    """Above 64 unknowns `solve` runs the blocked Cholesky (symmetric, definite diagonal: blasLapackWrapper.cpp:590-610) or
    the blocked LU with partial pivoting (:570); the last block is ragged for every size here but 64."""
    rng = np.random.default_rng(n + nrhs)
    G = rng.standard_normal((n, n))
    spd = G @ G.T + n * np.eye(n)
    B = rng.standard_normal((n, nrhs))
    X = BW.solve(spd, B)
    assert rel(X, np.linalg.solve(spd, B)) < 1e-11 and rel(spd @ X, B) < 1e-12
    X = BW.solve(G, B)
    assert rel(G @ X, B) < 1e-9 and rel(X, np.linalg.solve(G, B)) < 1e-7           # cond(G) ~ n .. 1e4
    P = np.eye(n)[rng.permutation(n)] * rng.choice([-1.0, 1.0], n)                   # pivoting is unavoidable here
    X = BW.solve(P, B)
    assert rel(X, P.T @ B) < 1e-14
    sym_indef = G + G.T                                                              # symmetric, indefinite: Cholesky fails, LU takes over
    X = BW.solve(sym_indef, B)
    assert rel(sym_indef @ X, B) < 1e-8
    with pytest.raises(xb.XerusError):
        BW.solve(np.ones((n, n)), B)                                                 # singular


def test_error_behaviour():
    with pytest.raises(xb.XerusError):
        BW.matrix_matrix_product(1.0, np.ones((3, 4)), False, np.ones((5, 2)), False)
    with pytest.raises(xb.XerusError):
        BW.qr(np.ones((0, 3)))


def test_unfoldings_taller_than_the_grid_y_limit():
    """TT unfoldings have millions of rows: to_dense / from_dense of a 2^23 tensor multiply and transpose matrices whose row tiles
    exceed the 65535 limit of grid.y (the m tiles go on grid.x then)."""
    rng = np.random.default_rng(8)
    d = 23
    t = xb.TTTensor.random([2] * d, 2, rng)
    full = t.to_dense()                                   # last product: (2^22 x 2) * (2 x 2): 131072 row tiles of 32
    assert full.shape == (2,) * d
    cores = t.cores()
    ref = cores[0].reshape(2, -1)
    for c in cores[1:]:
        ref = (ref @ c.reshape(c.shape[0], -1)).reshape(-1, c.shape[-1])
    assert np.linalg.norm(full.reshape(-1) - ref.reshape(-1)) < 1e-12 * np.linalg.norm(ref)
    back = xb.TTTensor.from_dense(full, 1e-12)            # first SVD sees a (2^22 x 2) matrix
    assert back.ranks() == t.ranks()
    assert back.distance(t) < 1e-10 * t.frob_norm()
    # transpose with more than 65535 row tiles, and a GEMM with that many m tiles, directly
    A = rng.standard_normal((2_200_000, 3))
    assert np.array_equal(xb.reshuffle(A, [1, 0]), A.T)
    B = rng.standard_normal((3, 5))
    assert rel(xb.blasWrapper.matrix_matrix_product(1.0, A, False, B, False), A @ B) < 1e-13


def test_svd_beyond_the_persistent_kernels():
    """More than 512 rows per working column after the reduction / more block pairs than SMs: the launch-per-round Jacobi kernel
    (jacobi_block_kernel), the fallback of Svd::factor that the cooperative kernels leave to large matrices."""
    rng = np.random.default_rng(12)
    A = rng.standard_normal((1100, 1100))
    U, S, Vt = BW.svd(A)
    s_ref = np.linalg.svd(A, compute_uv=False)
    assert np.max(np.abs(S - s_ref)) < 1e-12 * s_ref[0]
    assert rel((U * S) @ Vt, A) < 1e-12
    assert np.linalg.norm(U.T @ U - np.eye(1100)) < 1e-10 and np.linalg.norm(Vt @ Vt.T - np.eye(1100)) < 1e-10


def test_call_registry_uses_the_reference_shape_strings():
    """xb_perf_*: the (group, name, shape) registry of the reference's XERUS_PERFORMANCE_ANALYSIS (misc/performanceAnalysis.h:30-39)
    at the C ABI, with the strings of blasLapackWrapper.cpp:83-720."""
    rng = np.random.default_rng(5)
    xb.perf_reset()
    xb.perf_enable(True)
    try:
        A, B = rng.standard_normal((37, 23)), rng.standard_normal((23, 19))
        BW.matrix_matrix_product(1.0, A, False, B, False)
        BW.matrix_matrix_product(1.0, A, False, B, False)
        BW.svd(A)
        BW.qr(A)
        D = rng.standard_normal((20, 7)) @ rng.standard_normal((7, 30))
        BW.qc(D)
        t = xb.TTTensor.random([3, 4, 3, 4], 5, rng)
        t.round(2)
    finally:
        xb.perf_enable(False)
    ent = xb.perf_entries()
    assert ent[("Dense BLAS", "Matrix-Matrix-Multiplication", "37x23 * 23x19")][0] == 2
    assert ent[("Dense LAPACK", "Singular Value Decomposition", "37x23")][0] == 1
    assert ent[("Dense LAPACK", "QR Factorisation", "37x23")][0] == 1
    assert ("Dense LAPACK", "QRP Factorisation", "20x7 * 7x30") in ent
    assert any(k[0] == "TT sweep" and k[1] == "round" for k in ent)
    assert all(v[1] > 0 for v in ent.values())
    assert "Matrix-Matrix-Multiplication" in xb.perf_analysis()
    BW.qr(A)                                  # disabled: nothing is recorded
    assert xb.perf_entries()[("Dense LAPACK", "QR Factorisation", "37x23")][0] == 1


@pytest.mark.parametrize("m,n", [(1, 1), (1, 7), (7, 1), (33, 2), (2, 33), (5, 300), (300, 5), (31, 31), (32, 64), (512, 32), (32, 512)])
def test_single_cta_kernels_on_ragged_shapes(m, n):
    """qr_small_kernel / svd_small_kernel (min(m, n) <= 32, csrc/small_f64.cu): odd column counts (a zero column pads the
    tournament), one-column and one-row matrices, the largest rows the kernels take, both orientations — against the general
    kernels (small_kernels=0) and against numpy."""
    rng = np.random.default_rng(m * 1000 + n)
    A = rng.standard_normal((m, n)) * 10.0 ** rng.uniform(-3, 3)
    k = min(m, n)
    out = {}
    try:
        for v in (1, 0):
            xb.set_option("small_kernels", v)
            U, S, Vt = BW.svd(A)
            Q, R = BW.qr(A)
            Rq, Qr = BW.rq(A)
            out[v] = (S, np.abs(np.diag(R[:k, :k])))
            assert rel((U * S) @ Vt, A) < 1e-13 and rel(Q @ R, A) < 1e-13 and rel(Rq @ Qr, A) < 1e-13
            assert np.linalg.norm(U.T @ U - np.eye(k)) < 1e-12 and np.linalg.norm(Vt @ Vt.T - np.eye(k)) < 1e-12
            assert np.linalg.norm(Q.T @ Q - np.eye(k)) < 1e-13
            assert np.all(np.diff(S) <= 0)
    finally:
        xb.set_option("small_kernels", 1)
    s_ref = np.linalg.svd(A, compute_uv=False)
    assert np.max(np.abs(out[1][0] - s_ref)) < 1e-13 * s_ref[0]
    assert np.max(np.abs(out[1][0] - out[0][0])) < 1e-13 * s_ref[0]
    assert np.allclose(out[1][1], out[0][1], rtol=1e-10, atol=1e-13 * out[0][1].max())


def test_single_cta_svd_of_zero_and_rank_deficient_matrices():
    Z = np.zeros((6, 9))
    U, S, Vt = BW.svd(Z)
    assert np.all(S == 0) and np.all(np.isfinite(U)) and np.all(np.isfinite(Vt))
    rng = np.random.default_rng(2)
    D = rng.standard_normal((40, 3)) @ rng.standard_normal((3, 20))          # rank 3 of 20
    U, S, Vt = BW.svd(D)
    assert rel((U * S) @ Vt, D) < 1e-13 and np.all(S[3:] < 1e-13 * S[0])
    Q, C, r = BW.qc(D)
    assert r == 3 and rel(Q @ C, D) < 1e-12


@pytest.fixture
def chol_everywhere():
    """Cholesky-QR2 from the first row on (default: from 1024 rows, where it overtakes the cluster panel kernels)."""
    xb.set_option("qr_chol_min_rows", 0)
    yield
    xb.set_option("qr_chol_min_rows", 1024)


def _graded(m, n, cond, seed):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    return (U * np.logspace(0, -np.log10(cond), n)) @ V.T


def _qr_paths(A):
    """BW.qr(A) with the per-class profile on: (Q, R, taken, declined) where taken / declined count Cholesky-QR2 attempts."""
    xb.profile_enable(True)
    try:
        t0, d0 = xb.profile_get("qr_chol")[0], xb.profile_get("qr_chol_declined")[0]
        Q, R = BW.qr(A)
        t1, d1 = xb.profile_get("qr_chol")[0], xb.profile_get("qr_chol_declined")[0]
    finally:
        xb.profile_enable(False)
    return Q, R, t1 - t0, d1 - d0


@pytest.mark.parametrize("m,n", [(64, 33), (128, 128), (512, 128), (500, 50), (4096, 16), (3000, 120), (97, 97), (20000, 100), (300, 65), (2000, 7), (509, 60),
                                  (4096, 150), (3000, 200), (2500, 256), (2049, 129)])
@pytest.mark.parametrize("cond", [1.0, 1e2, 1e4])
def test_cholesky_qr2_on_well_conditioned_tall_matrices(m, n, cond, chol_everywhere):
    """Tall QRs of up to 128 columns take the Cholesky-QR2 path (csrc/qr_f64.cu: cholqr2), and so do 129..256 columns beyond 2048
    rows (two column halves, block Gram-Schmidt in between): orthogonality and residual at the level of the Householder path,
    R equal to LAPACK's up to row signs, and the same factors with the path switched off."""
    A = _graded(m, n, cond, m + n) * 3.0e7
    Q, R, taken, declined = _qr_paths(A)
    assert taken == 1 and declined == 0
    assert rel(Q @ R, A) < 1e-14
    assert np.linalg.norm(Q.T @ Q - np.eye(n)) < 5e-14
    assert np.array_equal(np.tril(R, -1), np.zeros_like(R)) and np.all(np.diag(R) > 0)
    Qo, Ro = O.qr(A)
    s = np.sign(np.diag(Ro))
    assert rel(s[:, None] * Ro, R) < 1e-12 * cond
    try:
        xb.set_option("qr_chol", 0)
        Qh, Rh, taken, _ = _qr_paths(A)
    finally:
        xb.set_option("qr_chol", 1)
    assert taken == 0
    sh = np.sign(np.diag(Rh))
    assert rel(sh[:, None] * Rh, R) < 1e-12 * cond and rel(Qh * sh, Q) < 1e-11 * cond


@pytest.mark.parametrize("case", ["cond1e7", "cond1e13", "rank_deficient", "zero_column", "huge", "tiny", "zero", "nan"])
def test_cholesky_qr2_declines_what_it_cannot_do(case, chol_everywhere):
    """Ill-conditioned, rank-deficient or badly scaled input is declined on the device and factored by Householder reflections:
    the caller sees the same contract either way."""
    m, n = 600, 80
    rng = np.random.default_rng(5)
    if case.startswith("cond"):
        A = _graded(m, n, float(case[4:]), 3)
    elif case == "rank_deficient":
        A = rng.standard_normal((m, 20)) @ rng.standard_normal((20, n))
    elif case == "zero_column":
        A = rng.standard_normal((m, n)); A[:, 17] = 0.0
    elif case == "huge":
        A = rng.standard_normal((m, n)) * 1e140                           # A^T A overflows unscaled
    elif case == "tiny":
        A = rng.standard_normal((m, n)) * 1e-140
    elif case == "zero":
        A = np.zeros((m, n))
    else:
        A = rng.standard_normal((m, n)); A[5, 7] = np.nan
    Q, R, taken, declined = _qr_paths(A)
    assert taken == 1 and (declined == 1 or case == "cond1e7")              # 1e7 is inside what the second pass can repair
    if case == "nan":
        return                                                              # garbage in, garbage out — but no hang and no exception
    assert rel(Q @ R, A) < 1e-13 if case != "zero" else np.all(R == 0)
    assert np.linalg.norm(Q.T @ Q - np.eye(n)) < 1e-12
    assert np.array_equal(np.tril(R, -1), np.zeros_like(R))


def test_cholesky_qr2_backs_off_inside_a_call_and_plans_replay_its_decisions(chol_everywhere):
    """A sweep over ill-conditioned cores stops trying after two declines in a row (and tries again 16 candidates later); a round
    plan records which candidates were accepted and replays exactly those, so plan and ordinary path stay bit-identical."""
    rng = np.random.default_rng(9)
    d, n, r = 8, 4, 40
    dims = [n] * d
    ranks = [1] + [min(r, n ** min(i, d - i)) for i in range(1, d)] + [1]
    cores = [rng.standard_normal((ranks[i], n, ranks[i + 1])) for i in range(d)]
    # a graded bond in the middle: the QRs that carry it are declined, the others are not
    cores[4] = np.einsum("a,anb->anb", np.logspace(0, -9, ranks[4]), cores[4])
    results = []
    for plans, repeats in ((0, 1), (1, 3)):                                  # ordinary; first sight, capture + replay, replay
        xb.set_option("round_plans", plans)
        for _ in range(repeats):
            t = xb.TTTensor.from_cores(cores)
            t.round(24)
            results.append(t.cores())
    for other in results[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(results[0], other))
    ref = O.TT([c.copy() for c in cores])
    ref.round(24)
    got = O.TT(results[0], core_position=0)
    assert got.ranks() == ref.ranks() and O.tt_distance_rel(got, ref) < 1e-10


def test_round_plan_drops_cholesky_qr2_after_repeated_declines(chol_everywhere):
    """A plan recorded on well-conditioned cores speculates that its Cholesky-QR2 candidates are accepted.  TTs of the same shape
    with a graded bond fail that check on replay (reason 4): the call falls back to the ordinary path, and after the second
    failure the plan is recorded again Householder-only.  Every result along the way must be right."""
    rng = np.random.default_rng(21)
    d, n, r = 6, 4, 40
    ranks = [1] + [min(r, n ** min(i, d - i)) for i in range(1, d)] + [1]

    def cores(graded):
        cs = [rng.standard_normal((ranks[i], n, ranks[i + 1])) for i in range(d)]
        if graded:
            cs[3] = np.einsum("a,anb->anb", np.logspace(0, -9, ranks[3]), cs[3])
        return cs

    xb.set_option("round_plans", 1)                       # new option epoch: no plan for this shape yet
    for graded in (False, False, True, True, True, True, False, True):
        cs = cores(graded)
        t = xb.TTTensor.from_cores(cs)
        t.round(24)
        ref = O.TT([c.copy() for c in cs])
        ref.round(24)
        got = O.TT(t.cores(), core_position=0)
        assert got.ranks() == ref.ranks()
        assert O.tt_distance_rel(got, ref) < 1e-10, graded
