"""TEST INFRASTRUCTURE ONLY — the CPU oracle.

A numpy/LAPACK restatement of the reference's (xerus v3.0.1) tensor-train hot path, written from the reference's
*behaviour*; every function cites the reference file:line it follows.  It exists so that the CUDA path can be
checked on arbitrary seeded inputs, at sizes where the compiled reference (`oracle/_ref`) is not available (the
reference tree does not travel to the GPU box).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import this
module, and only as the checker.  The product (`xerus_b200/`) never imports it.

Pinning: `tests/test_oracle_golden.py` checks every function below against golden vectors produced by the
unmodified reference itself (`oracle/_ref/ref_golden`, committed as `tests/golden/xerus_ref_v1.npz`), against
the reference's own known-answer tests (src/unitTests/fullTensor_product.cxx) and against SURVEY.md Appendix B.

Arithmetic below the reference's blasWrapper boundary is OpenBLAS + reference LAPACK (not vendored in the
reference, no pinned version: config.mk.default:118-126); here it is numpy/scipy's LAPACK (dgeqrf, dgeqp3,
dgesdd, dpotrf/dgesv), i.e. the same published algorithms.
"""
from __future__ import annotations

import copy
import numpy as np
import scipy.linalg as sla

EPSILON = 8 * np.finfo(np.float64).eps  # include/xerus/basic.h:50


# ----------------------------------------------------------------------------------------------------------------------
# L0: blasWrapper  (src/xerus/blasLapackWrapper.cpp)
# ----------------------------------------------------------------------------------------------------------------------

def matrix_matrix_product(alpha, A, transA, B, transB):
    """C = alpha * op(A) * op(B), row-major, beta = 0 (blasLapackWrapper.cpp:149-195)."""
    a = A.T if transA else A
    b = B.T if transB else B
    return alpha * (a @ b)


def qr(A):
    """Unpivoted thin QR, Q m x min, R min x n (blasLapackWrapper.cpp:388-438: dgeqrf + dorgqr)."""
    Q, R = np.linalg.qr(A, mode="reduced")
    return Q, R


def rq(A):
    """A = R * Q, R m x min (upper-trapezoidal), Q min x n with orthonormal rows (:455-498: dgerqf + dorgrq)."""
    R, Q = sla.rq(A, mode="economic")
    return R, Q


def qc(A, signed_quirk=True):
    """Column-pivoted QR with rank detection: A = Q (m x rank) * C (rank x n)  (:243-305).

    rank = first k >= 1 with |R[k,k]| < 16*eps*R[0,0]  (:266-272).  The reference compares against the *signed*
    R[0,0]; with `signed_quirk=True` this restatement does the same (rank reduction then only happens when
    LAPACK's R[0,0] is positive); with False it uses |R[0,0]|, which is the rule the CUDA path documents."""
    m, n = A.shape
    maxrank = min(m, n)
    Qf, Rf, P = sla.qr(A, mode="economic", pivoting=True)
    r00 = Rf[0, 0] if signed_quirk else abs(Rf[0, 0])
    rank = maxrank
    for k in range(1, maxrank):
        if abs(Rf[k, k]) < 16 * np.finfo(np.float64).eps * r00:
            rank = k
            break
    C = np.zeros((rank, n))
    C[:, P] = Rf[:rank, :]            # C = R * P^T, rows beyond rank dropped (:279-285)
    return Qf[:, :rank].copy(), C, rank


def cq(A, signed_quirk=True):
    """A = C (m x rank) * Q (rank x n): the pivoted QR of A^T (:317-371, dgeqp3 col-major on the row-major buffer)."""
    Qt, Ct, rank = qc(np.ascontiguousarray(A.T), signed_quirk)
    return np.ascontiguousarray(Ct.T), np.ascontiguousarray(Qt.T), rank


def svd(A):
    """Thin SVD, singular values descending (:201-232: dgesdd 'S')."""
    U, S, Vt = np.linalg.svd(A, full_matrices=False)
    return U, S, Vt


def _is_symmetric(A):
    """(:501-516) note: `max` is the maximum *signed* entry."""
    mx = max(0.0, float(A.max()))
    return bool(np.all(np.abs(A - A.T) < 4 * mx * np.finfo(np.float64).eps) or A.shape[0] == 1)


def _pos_neg_definite_diagonal(A):
    """(:519-537)"""
    d = np.diag(A)
    eps = np.finfo(np.float64).eps
    if d[0] > 0:
        return bool(np.all(d[1:] >= eps))
    return bool(np.all(d[1:] <= -eps))


def solve(A, b):
    """Solver dispatch of blasWrapper::solve (:542-651): LS if non-square, LU if non-symmetric, Cholesky if the
    diagonal is definite (falling back to LDL^T when the factorisation fails), else LDL^T."""
    m, n = A.shape
    if m != n:
        return np.linalg.lstsq(A, b, rcond=EPSILON)[0]
    if not _is_symmetric(A):
        return np.linalg.solve(A, b)
    if _pos_neg_definite_diagonal(A):
        try:
            c = sla.cho_factor(A, lower=False)
            return sla.cho_solve(c, b)
        except np.linalg.LinAlgError:
            pass
    return sla.solve(A, b, assume_a="sym")


# ----------------------------------------------------------------------------------------------------------------------
# L1: Tensor free functions  (src/xerus/tensor.cpp, indexedTensor_tensor_evaluate.cpp)
# ----------------------------------------------------------------------------------------------------------------------

def contract(lhs, lhs_trans, rhs, rhs_trans, num_modes):
    """xerus::contract (tensor.cpp:1252-1352): contract `num_modes` trailing modes of lhs (leading if lhs_trans)
    with `num_modes` leading modes of rhs (trailing if rhs_trans); one dgemm on the matricisations (:1310)."""
    ld, rd = lhs.shape, rhs.shape
    if lhs_trans:
        mid_l, keep_l = ld[:num_modes], ld[num_modes:]
    else:
        keep_l, mid_l = ld[:len(ld) - num_modes], ld[len(ld) - num_modes:]
    if rhs_trans:
        keep_r, mid_r = rd[:len(rd) - num_modes], rd[len(rd) - num_modes:]
    else:
        mid_r, keep_r = rd[:num_modes], rd[num_modes:]
    assert tuple(mid_l) == tuple(mid_r), "contracted dimensions do not coincide (tensor.cpp:1273-1279)"
    mid = int(np.prod(mid_l, dtype=np.int64))
    L = lhs.reshape(mid, -1) if lhs_trans else lhs.reshape(-1, mid)
    R = rhs.reshape(-1, mid) if rhs_trans else rhs.reshape(mid, -1)
    C = matrix_matrix_product(1.0, L, lhs_trans, R, rhs_trans)
    return C.reshape(tuple(keep_l) + tuple(keep_r))


def reshuffle(t, shuffle):
    """xerus::reshuffle (indexedTensor_tensor_evaluate.cpp:55-137): out mode shuffle[i] = in mode i."""
    inv = np.argsort(np.asarray(shuffle))
    return np.ascontiguousarray(np.transpose(t, inv))


def truncation_rank(S, max_rank, eps):
    """Rank rule of calculate_svd (tensor.cpp:1464-1474). max_rank == 0 means 'no cap'."""
    rank = len(S)
    if max_rank != 0:
        rank = min(rank, max_rank)
    for j in range(1, rank):
        if S[j] <= eps * S[0]:
            rank = j
            break
    return rank


def calculate_svd(t, split_pos, max_rank, eps):
    """calculate_svd (tensor.cpp:1424-1489): U (..., k), S (k,), Vt (k, ...)."""
    lhs = int(np.prod(t.shape[:split_pos], dtype=np.int64))
    U, S, Vt = svd(t.reshape(lhs, -1))
    k = truncation_rank(S, max_rank, eps)
    return (U[:, :k].reshape(t.shape[:split_pos] + (k,)), S[:k].copy(), Vt[:k].reshape((k,) + t.shape[split_pos:]))


# ----------------------------------------------------------------------------------------------------------------------
# L3/L4: TT network  (src/xerus/tensorNetwork.cpp, ttNetwork.cpp)
# ----------------------------------------------------------------------------------------------------------------------

class TT:
    """Tensor train (order-3 cores (r_l, n, r_r)) or TT operator (order-4 cores (r_l, m, n, r_r)).
    State mirrors TTNetwork: `canonicalized`, `core_position` (ttNetwork.h:52-58)."""

    def __init__(self, cores, core_position=None):
        self.cores = [np.array(c, dtype=np.float64, order="C") for c in cores]
        self.canonicalized = core_position is not None
        self.core_position = core_position if core_position is not None else 0
        for a, b in zip(self.cores[:-1], self.cores[1:]):
            assert a.shape[-1] == b.shape[0], "bond dimensions do not coincide"
        assert self.cores[0].shape[0] == 1 and self.cores[-1].shape[-1] == 1

    # -- structure -----------------------------------------------------------------------------------------------
    @property
    def d(self):
        return len(self.cores)

    def copy(self):
        return copy.deepcopy(self)

    def ranks(self):
        """ttNetwork.cpp:717-724"""
        return [c.shape[-1] for c in self.cores[:-1]]

    def dims(self):
        return [c.shape[1:-1] for c in self.cores]

    def exceeds_maximal_ranks(self):
        """ttNetwork.cpp:349-359: a bond may not exceed the product of the external dims on either side of it."""
        d = self.d
        for i in range(d):
            c = self.cores[i]
            ext = int(np.prod(c.shape[1:-1], dtype=np.int64))
            if c.shape[0] > ext * c.shape[-1] or c.shape[-1] > ext * c.shape[0]:
                return True
        return False

    # -- edge operations -----------------------------------------------------------------------------------------
    def transfer_core(self, frm, to, allow_rank_reduction=True, signed_quirk=True):
        """TensorNetwork::transfer_core (tensorNetwork.cpp:821-909) for TT chains: the shared bond is the last mode
        of the left core / first mode of the right core, so no reshuffle is needed (:833-847)."""
        F, T = self.cores[frm], self.cores[to]
        if to == frm + 1:      # move right: QC / QR of the left matricisation (:839-847)
            M = F.reshape(-1, F.shape[-1])
            if allow_rank_reduction:
                Q, R, _ = qc(M, signed_quirk)
            else:
                Q, R = qr(M)
            self.cores[frm] = Q.reshape(F.shape[:-1] + (Q.shape[1],))
            self.cores[to] = contract(R, False, T, False, 1)                 # (:877)
        elif to == frm - 1:    # move left: CQ / RQ of the right matricisation (:833-838)
            M = F.reshape(F.shape[0], -1)
            if allow_rank_reduction:
                R, Q, _ = cq(M, signed_quirk)
            else:
                R, Q = rq(M)
            self.cores[frm] = Q.reshape((Q.shape[0],) + F.shape[1:])
            self.cores[to] = contract(T, False, R, False, 1)                 # (:880), transR && !transR
        else:
            raise ValueError("not neighbours")

    def round_edge(self, frm, to, max_rank, eps, soft_threshold=0.0, signed_quirk=True):
        """TensorNetwork::round_edge (tensorNetwork.cpp:678-818) as TTNetwork::round calls it: `frm` is the right
        core (bond = its mode 0, transFrom), `to` the left core (bond = its last mode, transTo).  Returns the kept
        singular values."""
        assert to == frm - 1
        F, T = self.cores[frm], self.cores[to]
        r = F.shape[0]
        if 5 * F.size * T.size >= 6 * r ** 4:                                 # prior-QR branch (:745)
            coreA, Fq, _ = cq(F.reshape(r, -1), signed_quirk)                # from = coreA * Fq   (:749)
            Tq, coreB, _ = qc(T.reshape(-1, r), signed_quirk)                # to   = Tq * coreB   (:755)
            X = contract(coreA, True, coreB, True, 1)                        # X = coreA^T coreB^T (:761)
            U, S, Vt = calculate_svd(X, 1, max_rank, eps)                    # (:764)
            S = np.maximum(0.0, S - soft_threshold)                          # (:766)
            Vt = S[:, None] * Vt                                             # (:769)
            newF = contract(U, True, Fq, False, 1)                           # (:773)
            newT = contract(Tq, False, Vt, True, 1)                          # (:779)
            self.cores[frm] = newF.reshape((newF.shape[0],) + F.shape[1:])
            self.cores[to] = newT.reshape(T.shape[:-1] + (newT.shape[-1],))
        else:                                                                 # direct branch (:784-803)
            X = contract(F, True, T, True, 1)                                # (ext_F..., ext_T...)
            nF = F.ndim - 1
            U, S, Vt = calculate_svd(X, nF, max_rank, eps)
            S = np.maximum(0.0, S - soft_threshold)
            k = len(S)
            # toTensor = toTensor^T * S^T : (k, ext_T...) -> (ext_T..., k)            (:791)
            newT = (S[:, None] * Vt.reshape(k, -1)).T
            self.cores[to] = np.ascontiguousarray(newT).reshape(T.shape[:-1] + (k,))
            # fromTensor = U reshuffled so that the new bond is mode 0                (:796-802)
            self.cores[frm] = np.ascontiguousarray(np.moveaxis(U, -1, 0))
        return S

    # -- sweeps --------------------------------------------------------------------------------------------------
    def move_core(self, position, keep_rank=False, signed_quirk=True):
        """TTNetwork::move_core (ttNetwork.cpp:582-628)."""
        d = self.d
        assert position < d
        arr = not keep_rank
        if self.canonicalized:
            for n in range(self.core_position, position):
                self.transfer_core(n, n + 1, arr, signed_quirk)
            for n in range(self.core_position, position, -1):
                self.transfer_core(n, n - 1, arr, signed_quirk)
        else:
            for n in range(0, position):
                self.transfer_core(n, n + 1, arr, signed_quirk)
            for n in range(d - 1, position, -1):
                self.transfer_core(n, n - 1, arr, signed_quirk)
        while self.exceeds_maximal_ranks():                                   # (:609-624)
            for n in range(position, 0, -1):
                self.transfer_core(n, n - 1, arr, signed_quirk)
            for n in range(0, d - 1):
                self.transfer_core(n, n + 1, arr, signed_quirk)
            for n in range(d - 1, position, -1):
                self.transfer_core(n, n - 1, arr, signed_quirk)
        self.canonicalized = True
        self.core_position = position

    def round(self, max_ranks=None, eps=EPSILON, signed_quirk=True):
        """TTNetwork::round (ttNetwork.cpp:644-684).  `max_ranks`: int, list of d-1 ints, or None (no cap).
        Returns the list of kept singular values per edge (edge d-2 first), which the reference does not expose."""
        d = self.d
        assert eps < 1
        if max_ranks is None:
            max_ranks = [0] * (d - 1)          # 0 = unlimited (size_t max in the reference, :682-684)
        elif np.isscalar(max_ranks):
            max_ranks = [int(max_ranks)] * (d - 1)
        assert len(max_ranks) == d - 1
        initial_canon, initial_core = self.canonicalized, self.core_position
        self.move_core(d - 1, False, signed_quirk)                            # canonicalize_right (:654)
        svals = []
        for i in range(d - 1):                                                # (:656-658)
            svals.append(self.round_edge(d - 1 - i, d - 2 - i, max_ranks[d - 2 - i], eps, 0.0, signed_quirk))
        self.canonicalized, self.core_position = True, 0                      # assume_core_position(0) (:660)
        if initial_canon:
            self.move_core(initial_core, False, signed_quirk)                 # (:662-664)
        return svals

    def soft_threshold(self, taus, signed_quirk=True):
        """TTNetwork::soft_threshold (ttNetwork.cpp:688-713): the sweep of round() with no rank cap, eps = 0 and every singular
        value replaced by max(0, sigma - tau); taus[i] belongs to the i-th edge from the right (:700)."""
        d = self.d
        if np.isscalar(taus):
            taus = [float(taus)] * (d - 1)
        assert len(taus) == d - 1
        initial_canon, initial_core = self.canonicalized, self.core_position
        self.move_core(d - 1, False, signed_quirk)
        for i in range(d - 1):
            self.round_edge(d - 1 - i, d - 2 - i, 0, 0.0, taus[i], signed_quirk)
        self.canonicalized, self.core_position = True, 0
        if initial_canon:
            self.move_core(initial_core, False, signed_quirk)

    # -- values --------------------------------------------------------------------------------------------------
    def to_dense(self):
        """operator Tensor() (tensorNetwork.cpp:287-306) for a chain; operators come out as (m_1..m_d, n_1..n_d)."""
        res = self.cores[0]
        for c in self.cores[1:]:
            res = np.tensordot(res, c, axes=([res.ndim - 1], [0]))
        res = res.reshape(res.shape[1:-1])
        if self.cores[0].ndim == 4:
            d = self.d
            res = np.transpose(res, list(range(0, 2 * d, 2)) + list(range(1, 2 * d, 2)))
        return np.ascontiguousarray(res)

    def frob_norm(self):
        return float(np.sqrt(max(0.0, tt_inner(self, self))))


def reduce_to_maximal_ranks(ranks, dims):
    """TTNetwork::reduce_to_maximal_ranks (ttNetwork.cpp:370-402)."""
    ranks = list(ranks)
    d = len(dims)
    cur = 1
    for i in range(d - 1):
        cur *= dims[i]
        if cur < ranks[i]:
            ranks[i] = cur
        else:
            cur = ranks[i]
    cur = 1
    for i in range(d - 1, 0, -1):
        cur *= dims[i]
        if cur < ranks[i - 1]:
            ranks[i - 1] = cur
        else:
            cur = ranks[i - 1]
    return ranks


def tt_random(dims, ranks, rng, move_core=True):
    """TTTensor::random (ttNetwork.h:129-155) with a numpy Generator instead of libstdc++'s mt19937_64 stream
    (RNG parity is not a goal; inputs for parity tests are exchanged as cores)."""
    if np.isscalar(ranks):
        ranks = [int(ranks)] * (len(dims) - 1)
    rk = [1] + reduce_to_maximal_ranks(ranks, dims) + [1]
    t = TT([rng.standard_normal((rk[i], dims[i], rk[i + 1])) for i in range(len(dims))])
    if move_core:
        t.move_core(0)
    return t


def tt_ones(dims):
    return TT([np.ones((1, n, 1)) for n in dims], core_position=0)


def tt_inner(a, b):
    """<a, b> by the usual left-to-right transfer contraction (what `a(i&0)*b(i&0)` evaluates to)."""
    E = np.ones((1, 1))
    for ca, cb in zip(a.cores, b.cores):
        ra, rb = ca.shape[0], cb.shape[0]
        tmp = E.T @ ca.reshape(ra, -1)                               # (rb, n ra')
        tmp = tmp.reshape(rb * int(np.prod(ca.shape[1:-1])), -1)      # (rb n, ra')
        E = tmp.T @ cb.reshape(-1, cb.shape[-1])                      # (ra', rb')
    return float(E[0, 0])


def tt_distance_rel(a, b):
    """||a - b|| / ||b||, cancellation-free: stack the cores of a and -b block-diagonally, sweep QR left to right and
    take the Frobenius norm of the last core (the inner-product formula of SURVEY.md §8d loses half the digits)."""
    d = a.d
    if d == 1:
        return float(np.linalg.norm(a.cores[0] - b.cores[0]) / np.linalg.norm(b.cores[0]))
    carry = None
    for i, (ca, cb) in enumerate(zip(a.cores, b.cores)):
        ext = ca.shape[1:-1]
        la, ra, lb, rb = ca.shape[0], ca.shape[-1], cb.shape[0], cb.shape[-1]
        if i == 0:
            c = np.zeros((1,) + ext + (ra + rb,))
            c[..., :ra], c[..., ra:] = ca, -cb
        elif i == d - 1:
            c = np.zeros((la + lb,) + ext + (1,))
            c[:la], c[la:] = ca, cb
        else:
            c = np.zeros((la + lb,) + ext + (ra + rb,))
            c[:la, ..., :ra], c[la:, ..., ra:] = ca, cb
        if carry is not None:
            c = np.tensordot(carry, c, axes=([1], [0]))
        if i < d - 1:
            Q, carry = np.linalg.qr(c.reshape(-1, c.shape[-1]))
        else:
            return float(np.linalg.norm(c) / b.frob_norm())


def tt_add(a, b):
    """TTNetwork::operator+= (ttNetwork.cpp:797-847): block-diagonal stacking of the cores (first core: blocks side
    by side, last core: blocks on top of each other)."""
    d = a.d
    if d == 1:
        return TT([a.cores[0] + b.cores[0]], core_position=a.core_position if a.canonicalized else None)
    cores = []
    for i, (ca, cb) in enumerate(zip(a.cores, b.cores)):
        ext = ca.shape[1:-1]
        la, ra, lb, rb = ca.shape[0], ca.shape[-1], cb.shape[0], cb.shape[-1]
        if i == 0:
            c = np.zeros((1,) + ext + (ra + rb,))
            c[..., :ra], c[..., ra:] = ca, cb
        elif i == d - 1:
            c = np.zeros((la + lb,) + ext + (1,))
            c[:la], c[la:] = ca, cb
        else:
            c = np.zeros((la + lb,) + ext + (ra + rb,))
            c[:la, ..., :ra], c[la:, ..., ra:] = ca, cb
        cores.append(c)
    res = TT(cores)
    if a.canonicalized:                      # (:842-844) re-canonicalise with the rank-revealing move_core
        res.move_core(a.core_position)
    return res


def tt_apply(A, x):
    """y(i&0) = A(i/2,j/2) * x(j&0): TTStack collapse (ttStack.cpp:197-300): per site contract the operator core
    (a, m, n, b) with the tensor core (r, n, s) to ((a r), m, (b s)) — operator bond is the slow index."""
    cores = []
    for ca, cx in zip(A.cores, x.cores):
        t = np.einsum("amnb,rns->armbs", ca, cx)
        cores.append(t.reshape(ca.shape[0] * cx.shape[0], ca.shape[1], ca.shape[3] * cx.shape[2]))
    return TT(cores)


def tt_svd(full, eps=EPSILON, max_rank=0, is_operator=False):
    """TT-SVD constructor TTNetwork(Tensor, eps, maxRanks) (ttNetwork.cpp:112-160): successive SVDs from the right,
    Sigma pushed to the left remainder (:151-155).  max_rank: 0 = no cap, an int, or one cap per bond.  Operators: `full` has
    modes (m_1..m_d, n_1..n_d) and is reshuffled to (m_1,n_1,m_2,n_2,...) first (:129-135)."""
    N = 2 if is_operator else 1
    d = full.ndim // N
    if is_operator:
        shuffle = [0] * full.ndim
        for i in range(d):
            shuffle[i], shuffle[d + i] = 2 * i, 2 * i + 1
        full = reshuffle(full, shuffle)
    dims = full.shape
    caps = [int(max_rank)] * (d - 1) if np.isscalar(max_rank) else [int(r) for r in max_rank]
    cores = [None] * d
    remains = full.reshape(dims + (1,))
    for pos in range(d - 1, 0, -1):
        U, S, Vt = calculate_svd(remains, pos * N, caps[pos - 1], eps)
        cores[pos] = Vt
        remains = U * S
    cores[0] = remains.reshape((1,) + remains.shape)
    return TT(cores, core_position=0)


def laplace_operator(d, n):
    """Rank-2 Laplace-like TT operator used by the BASELINE configs."""
    L = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
    I = np.eye(n)
    cores = []
    for k in range(d):
        rl, rr = (1 if k == 0 else 2), (1 if k == d - 1 else 2)
        c = np.zeros((rl, n, n, rr))
        if d == 1:
            c[0, :, :, 0] = L
        elif k == 0:
            c[0, :, :, 0], c[0, :, :, 1] = L, I
        elif k == d - 1:
            c[0, :, :, 0], c[1, :, :, 0] = I, L
        else:
            c[0, :, :, 0], c[1, :, :, 0], c[1, :, :, 1] = I, L, I
        cores.append(c)
    return TT(cores)


# ----------------------------------------------------------------------------------------------------------------------
# L5: ALS / DMRG  (src/xerus/algorithms/als.cpp)
# ----------------------------------------------------------------------------------------------------------------------

class ALSVariant:
    """ALSVariant (als.h:37-223) with the lapack_solver local solver (als.cpp:43-71).

    sites = 1 (ALS) or 2 (DMRG); assume_spd selects x^T A x (SPD) or x^T A^T A x environments."""

    def __init__(self, sites=1, assume_spd=True, convergence_epsilon=1e-6, fix_dmrg_turn=False, solver="lapack"):
        self.solver = solver          # "lapack": ALSVariant::lapack_solver (als.cpp:43-71); "ASD": ALSVariant::ASD_solver (:73-103)
        self.sites = sites
        self.assume_spd = assume_spd
        self.convergence_epsilon = convergence_epsilon
        self.preserve_core_position = True
        # The reference pushes the wrong slice when a two-site sweep turns around (als.cpp:371,:376).
        # fix_dmrg_turn=True applies the obvious fix (push site currIndex+sites-1) so full sweeps can be checked.
        self.fix_dmrg_turn = fix_dmrg_turn

    # -- environment slices (als.cpp:184-215) ----------------------------------------------------------------------
    def _op_slice_left(self, env, xk, Ak):
        """env'(r1',r2',r3') = sum env(r1,r2,r3) x(r1,n1,r1') A(r2,n1,n2,r2') x(r3,n2,r3')   [SPD, :189-191]"""
        if self.assume_spd:
            t = np.einsum("abc,aid->bcid", env, xk)
            t = np.einsum("bcid,bije->cdje", t, Ak)
            return np.einsum("cdje,cjf->def", t, xk)
        # general: x(r1,n1,cr1) A(r2,n2,n1,cr2) A(r3,n2,n3,cr3) x(r4,n3,cr4)          [:193-197]
        t = np.einsum("abcd,aie->bcdie", env, xk)
        t = np.einsum("bcdie,bjif->cdejf", t, Ak)
        t = np.einsum("cdejf,cjkg->defkg", t, Ak)
        return np.einsum("defkg,dkh->efgh", t, xk)

    def _op_slice_right(self, env, xk, Ak):
        if self.assume_spd:
            t = np.einsum("def,aid->efai", env, xk)
            t = np.einsum("efai,bije->fabj", t, Ak)
            return np.einsum("fabj,cjf->abc", t, xk)
        t = np.einsum("efgh,aie->fghai", env, xk)
        t = np.einsum("fghai,bjif->ghabj", t, Ak)
        t = np.einsum("ghabj,cjkg->habck", t, Ak)
        return np.einsum("habck,dkh->abcd", t, xk)

    def _rhs_slice_left(self, env, bk, xk, Ak):
        if self.assume_spd or Ak is None:                                       # (:206-208)
            t = np.einsum("ab,aic->bic", env, bk)
            return np.einsum("bic,bid->cd", t, xk)
        t = np.einsum("abc,aid->bcid", env, bk)                                 # (:210-212)
        t = np.einsum("bcid,bije->cdje", t, Ak)
        return np.einsum("cdje,cjf->def", t, xk)

    def _rhs_slice_right(self, env, bk, xk, Ak):
        if self.assume_spd or Ak is None:
            t = np.einsum("cd,aic->dai", env, bk)
            return np.einsum("dai,bid->ab", t, xk)
        t = np.einsum("def,aid->efai", env, bk)
        t = np.einsum("efai,bije->fabj", t, Ak)
        return np.einsum("fabj,cjf->abc", t, xk)

    # -- the driver (als.cpp:483-553) ------------------------------------------------------------------------------
    def __call__(self, A, x, b, num_half_sweeps=0, convergence_epsilon=None):
        """Runs on `x` in place; returns the energy the reference returns (als.cpp:548)."""
        conv = self.convergence_epsilon if convergence_epsilon is None else convergence_epsilon
        sites = self.sites
        d = x.d
        target_rank = x.ranks()                                                 # (:324)
        canon_end, core_end = x.canonicalized, x.core_position
        # set_component on a non-core index clears `canonicalized` in the reference; emulate: the absorb loops
        # below touch components, so treat x as non-canonical when anything was absorbed.
        first, last = self._prepare_x_for_als(x, canon_end, core_end)
        Ac = A.cores if A is not None else [None] * d
        spd_like = self.assume_spd or A is None
        onesA = np.ones((1, 1, 1)) if spd_like else np.ones((1, 1, 1, 1))
        onesB = np.ones((1, 1)) if spd_like else np.ones((1, 1, 1))
        opL, opR, rhL, rhR = [onesA], [onesA], [onesB], [onesB]
        for i in range(d - 1, first + sites - 1, -1):                           # prepare_stacks (:238-244)
            if A is not None:
                opR.append(self._op_slice_right(opR[-1], x.cores[i], Ac[i]))
            rhR.append(self._rhs_slice_right(rhR[-1], b.cores[i], x.cores[i], Ac[i]))
        for i in range(0, first):                                               # (:245-251)
            if A is not None:
                opL.append(self._op_slice_left(opL[-1], x.cores[i], Ac[i]))
            rhL.append(self._rhs_slice_left(rhL[-1], b.cores[i], x.cores[i], Ac[i]))
        cur = first
        increasing = True
        norm_b = b.frob_norm()

        def energy_f():
            if A is not None and self.assume_spd:                               # (:265-278)
                xAx, bx = opL[-1], rhL[-1]
                for i in range(sites):
                    xAx = self._op_slice_left(xAx, x.cores[cur + i], Ac[cur + i])
                    bx = self._rhs_slice_left(bx, b.cores[cur + i], x.cores[cur + i], Ac[cur + i])
                return abs(0.5 * float(np.sum(xAx * opR[-1])) - float(np.sum(bx * rhR[-1])))
            if A is not None:                                                   # residual functional (:282-296)
                xAtAx, bAx = opL[-1], rhL[-1]
                for i in range(sites):
                    xAtAx = self._op_slice_left(xAtAx, x.cores[cur + i], Ac[cur + i])
                    bAx = self._rhs_slice_left(bAx, b.cores[cur + i], x.cores[cur + i], Ac[cur + i])
                v = float(np.sum(xAtAx * opR[-1])) - 2 * float(np.sum(bAx * rhR[-1]))
                return float(np.sqrt(v + norm_b ** 2) / norm_b)
            bx = rhL[-1]                                                        # (:305-316)
            for i in range(sites):
                bx = self._rhs_slice_left(bx, b.cores[cur + i], x.cores[cur + i], None)
            return 0.5 * float(np.sum(x.cores[cur] ** 2)) - float(np.sum(bx * rhR[-1]))

        last_e2, last_e, energy = 1e102, 1e101, 1e100
        energy = energy_f()
        half_sweeps = 0
        while True:
            if A is not None:
                xs = self._local_solve(opL[-1], opR[-1], rhL[-1], rhR[-1], Ac, b.cores, cur, increasing, target_rank, x.cores[cur])
                for p in range(sites):
                    x.cores[cur + p] = xs[p]
                if sites > 1:   # set_component on a non-core index clears the flag (ttNetwork.cpp:491)
                    x.canonicalized = False
            else:
                assert sites == 1, "approximation dmrg not implemented yet (als.cpp:543)"
                t = np.einsum("ab,aic->bic", rhL[-1], b.cores[cur])
                x.cores[cur] = np.einsum("bic,cd->bid", t, rhR[-1])
            # check_for_end_of_sweep (:426-475)
            if (not increasing and cur == first) or (increasing and cur == last - sites):
                half_sweeps += 1
                last_e2, last_e = last_e, energy
                energy = energy_f()
                if (half_sweeps == num_half_sweeps or abs(last_e - energy) < conv or abs(last_e2 - energy) < conv
                        or last - first <= sites):
                    if canon_end and self.preserve_core_position:
                        x.move_core(core_end, True)
                    return energy
                increasing = not increasing
            # move_to_next_index (:340-380)
            if increasing:
                if sites == 1:
                    x.move_core(cur + 1, True)
                if A is not None:
                    opR.pop()
                    opL.append(self._op_slice_left(opL[-1], x.cores[cur], Ac[cur]))
                rhR.pop()
                rhL.append(self._rhs_slice_left(rhL[-1], b.cores[cur], x.cores[cur], Ac[cur]))
                cur += 1
            else:
                if sites == 1:
                    x.move_core(cur - 1, True)
                pos = cur + sites - 1 if self.fix_dmrg_turn else cur
                if A is not None:
                    opL.pop()
                    opR.append(self._op_slice_right(opR[-1], x.cores[pos], Ac[pos]))
                rhL.pop()
                rhR.append(self._rhs_slice_right(rhR[-1], b.cores[pos], x.cores[pos], Ac[pos]))
                cur -= 1

    def _prepare_x_for_als(self, x, canon_end, core_end):
        """prepare_x_for_als (als.cpp:105-182): full-rank boundary cores are absorbed into their neighbour and
        replaced by reshaped identities; they are not optimised."""
        first, last = self._absorb_only(x)
        if first > 0 or last < x.d:
            # every absorb calls set_component on two indices, which clears `canonicalized` (ttNetwork.cpp:491)
            x.canonicalized = False
        if canon_end and core_end < first:
            x.canonicalized, x.core_position = True, first
        else:
            if canon_end and core_end >= last:
                x.canonicalized, x.core_position = True, last - 1
            x.move_core(first, True)
        return first, last

    def _absorb_only(self, x):
        d = x.d
        first, dim_prod = 0, 1
        while first + 1 < d:
            n_loc = x.cores[first].shape[1]
            new_prod = dim_prod * n_loc
            if x.cores[first].shape[-1] < new_prod:
                break
            cur = x.cores[first].reshape(-1, x.cores[first].shape[-1])
            x.cores[first + 1] = contract(cur, False, x.cores[first + 1], False, 1)
            x.cores[first] = np.eye(new_prod).reshape(dim_prod, n_loc, new_prod)
            first += 1
            dim_prod = new_prod
        last, dim_prod = d, 1
        while last > first + self.sites:
            n_loc = x.cores[last - 1].shape[1]
            new_prod = dim_prod * n_loc
            if x.cores[last - 2].shape[-1] < new_prod:
                break
            cur = x.cores[last - 1].reshape(x.cores[last - 1].shape[0], -1)
            x.cores[last - 2] = contract(x.cores[last - 2], False, cur, False, 1)
            x.cores[last - 1] = np.eye(new_prod).reshape(new_prod, n_loc, dim_prod)
            last -= 1
            dim_prod = new_prod
        return first, last

    # -- local problem (als.cpp:383-423, :43-71) -------------------------------------------------------------------
    def local_operator(self, envL, envR, Acores, cur):
        """Dense local operator with external order (l, n_1..n_s, r | l', n'_1..n'_s, r') (:391)."""
        sites = self.sites
        if self.assume_spd:
            T = envL                                                     # (l, a, l')
            T = np.moveaxis(T, 1, -1)                                     # (l, l', a)
            ext = []
            for p in range(sites):
                # T(..., a) * A(a, n, n', b) -> (..., n, n', b)
                T = np.tensordot(T, Acores[cur + p], axes=([T.ndim - 1], [0]))
            # T: (l, l', n1, n1', ..., ns, ns', b) ; envR (r, b, r')
            T = np.tensordot(T, envR, axes=([T.ndim - 1], [1]))          # (..., r, r')
            # reorder to (l, n1..ns, r, l', n1'..ns', r')
            idx_l, idx_lp = 0, 1
            n_idx = [2 + 2 * p for p in range(sites)]
            np_idx = [3 + 2 * p for p in range(sites)]
            r_idx, rp_idx = T.ndim - 2, T.ndim - 1
            T = np.transpose(T, [idx_l] + n_idx + [r_idx, idx_lp] + np_idx + [rp_idx])
        else:
            T = envL                                                     # (l, a1, a2, l')
            T = np.transpose(T, (0, 3, 1, 2))                             # (l, l', a1, a2)
            for p in range(sites):
                Ak = Acores[cur + p]
                # ATilde(.., n2, r3, r4, .., n4) = ATilde(.., r1, r2, ..) A(r1, x, n2, r3) A(r2, x, n4, r4)   (:393-397)
                T = np.tensordot(T, Ak, axes=([T.ndim - 2], [0]))        # (..., a2, x, n, b1)
                T = np.einsum("...axnb,axmc->...nmbc", T, Ak)            # (..., n, n', b1, b2)
            T = np.einsum("...bc,rbcs->...rs", T, envR)                   # envR (r, b1, b2, r')
            n_idx = [2 + 2 * p for p in range(sites)]
            np_idx = [3 + 2 * p for p in range(sites)]
            T = np.transpose(T, [0] + n_idx + [T.ndim - 2, 1] + np_idx + [T.ndim - 1])
        half = T.ndim // 2
        n_loc = int(np.prod(T.shape[:half]))
        return T.reshape(n_loc, n_loc), T.shape[:half]

    def local_rhs(self, envL, envR, Acores, bcores, cur):
        sites = self.sites
        if self.assume_spd or Acores[cur] is None:
            T = envL.T                                                   # (l, rb)
            for p in range(sites):
                T = np.tensordot(T, bcores[cur + p], axes=([T.ndim - 1], [0]))   # (l, n.., rb')
            T = np.tensordot(T, envR, axes=([T.ndim - 1], [0]))          # envR (rb, r) -> (l, n.., r)
        else:
            T = np.transpose(envL, (2, 0, 1))                            # (l, rb, a)
            for p in range(sites):
                # BTilde(.., n3, cr1, cr2) = BTilde(.., r1, r2) b(r1, n2, cr1) A(r2, n2, n3, cr2)        (:414-418)
                T = np.einsum("...ra,rnc,anmd->...mcd", T, bcores[cur + p], Acores[cur + p])
            T = np.einsum("...cd,cdr->...r", T, envR)                     # envR (rb, a, r)
        return T

    def _local_solve(self, envL, envR, rhsL, rhsR, Acores, bcores, cur, increasing, target_rank, xcur=None):
        Aloc, shape = self.local_operator(envL, envR, Acores, cur)
        bloc = self.local_rhs(rhsL, rhsR, Acores, bcores, cur)
        if self.solver == "ASD":                                                # (:73-103)
            assert self.sites == 1, "ASD only defined for single site alternation at the moment"
            xv = xcur.reshape(-1)
            grad = bloc.reshape(-1) - Aloc @ xv                                 # (:81)
            if self.assume_spd:
                alpha = float(grad @ grad) / float(grad @ (Aloc @ grad))        # (:85)
            else:
                grad = Aloc.T @ grad                                            # (:87)
                alpha = float(np.linalg.norm(grad)) / float(np.linalg.norm(Aloc @ grad))   # (:89): norms, not their squares
            return [(xv + alpha * grad).reshape(shape)]
        xloc = solve(Aloc, bloc.reshape(-1)).reshape(shape)
        sites = self.sites
        out = [None] * sites
        if increasing:                                                          # (:52-60)
            for p in range(sites - 1):
                U, S, Vt = calculate_svd(xloc, 2, target_rank[cur + p], EPSILON)
                out[p] = U
                xloc = S.reshape((-1,) + (1,) * (Vt.ndim - 1)) * Vt
            out[-1] = xloc
        else:                                                                   # (:61-70)
            for p in range(sites - 1, 0, -1):
                U, S, Vt = calculate_svd(xloc, xloc.ndim - 2, target_rank[cur + p - 1], EPSILON)
                out[p] = Vt
                xloc = U * S
            out[0] = xloc
        return out


ALS = ALSVariant(1, False)
ALS_SPD = ALSVariant(1, True)
DMRG = ALSVariant(2, False)
DMRG_SPD = ALSVariant(2, True)
ASD = ALSVariant(1, False, solver="ASD")
ASD_SPD = ALSVariant(1, True, solver="ASD")


def residual(A, x, b):
    """||A x - b|| / ||b|| (als.cpp:258-261) through TT arithmetic, cancellation-free."""
    return tt_distance_rel(tt_apply(A, x), b)
