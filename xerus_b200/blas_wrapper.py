"""Host mirror of the reference's `xerus::blasWrapper` namespace (include/xerus/blasLapackWrapper.h:37-146) and of the
`Tensor` free functions that sit directly on it (src/xerus/tensor.cpp): same names, argument meaning and error
behaviour, numpy arrays in place of raw row-major `double*`.  Every function runs on the GPU through the C ABI.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import XerusError, call

EPSILON = 8 * np.finfo(np.float64).eps   # include/xerus/basic.h:50


def _in(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_lib.dp)


def _out(shape):
    a = np.empty(shape, dtype=np.float64)
    return a, a.ctypes.data_as(_lib.dp)


def _sizes(v):
    arr = (C.c_size_t * max(1, len(v)))(*[int(x) for x in v])
    return arr


# ---- level 1 ----------------------------------------------------------------------------------------------------------
def one_norm(x):
    x, px = _in(x)
    r = C.c_double()
    call("xb_one_norm", px, x.size, C.byref(r))
    return r.value


def two_norm(x):
    x, px = _in(x)
    r = C.c_double()
    call("xb_two_norm", px, x.size, C.byref(r))
    return r.value


def dot_product(x, y):
    x, px = _in(x)
    y, py = _in(y)
    if x.size != y.size:
        raise XerusError(1, "dot_product: sizes differ")
    r = C.c_double()
    call("xb_dot_product", px, x.size, py, C.byref(r))
    return r.value


# ---- level 2 / 3 --------------------------------------------------------------------------------------------------------
def matrix_vector_product(alpha, A, transposed, y):
    """x = alpha * op(A) * y; A is stored (m x n), or (n x m) when transposed (blasLapackWrapper.cpp:114-131)."""
    A, pA = _in(A)
    y, py = _in(y)
    m, n = (A.shape[1], A.shape[0]) if transposed else A.shape
    if y.size != n:
        raise XerusError(1, "matrix_vector_product: sizes differ")
    x, px = _out((m,))
    call("xb_matrix_vector_product", px, m, float(alpha), pA, n, int(bool(transposed)), py)
    return x


def dyadic_vector_product(alpha, x, y):
    x, px = _in(x)
    y, py = _in(y)
    A, pA = _out((x.size, y.size))
    call("xb_dyadic_vector_product", pA, x.size, y.size, float(alpha), px, py)
    return A


def matrix_matrix_product(alpha, A, transposeA, B, transposeB):
    """C = alpha * op(A) * op(B) (blasLapackWrapper.cpp:149-195)."""
    A, pA = _in(A)
    B, pB = _in(B)
    left, mid = (A.shape[1], A.shape[0]) if transposeA else A.shape
    mid2, right = (B.shape[1], B.shape[0]) if transposeB else B.shape
    if mid != mid2:
        raise XerusError(1, "matrix_matrix_product: middle dimensions differ")
    Cm, pC = _out((left, right))
    call("xb_matrix_matrix_product", pC, left, right, float(alpha), pA, A.shape[1], int(bool(transposeA)), mid, pB,
         B.shape[1], int(bool(transposeB)))
    return Cm


# ---- LAPACK level -----------------------------------------------------------------------------------------------------
def svd(A):
    A, pA = _in(A)
    m, n = A.shape
    k = min(m, n)
    U, pU = _out((m, k))
    S, pS = _out((k,))
    Vt, pVt = _out((k, n))
    call("xb_svd", pU, pS, pVt, pA, m, n)
    return U, S, Vt


def qr(A):
    A, pA = _in(A)
    m, n = A.shape
    k = min(m, n)
    Q, pQ = _out((m, k))
    R, pR = _out((k, n))
    call("xb_qr", pQ, pR, pA, m, n)
    return Q, R


def rq(A):
    A, pA = _in(A)
    m, n = A.shape
    k = min(m, n)
    R, pR = _out((m, k))
    Q, pQ = _out((k, n))
    call("xb_rq", pR, pQ, pA, m, n)
    return R, Q


def qc(A):
    A, pA = _in(A)
    m, n = A.shape
    k = min(m, n)
    Q, pQ = _out((m * k,))
    Cm, pC = _out((k * n,))
    rank = C.c_size_t()
    call("xb_qc", pQ, pC, C.byref(rank), pA, m, n)
    r = rank.value
    return Q[:m * r].reshape(m, r).copy(), Cm[:r * n].reshape(r, n).copy(), r


def cq(A):
    A, pA = _in(A)
    m, n = A.shape
    k = min(m, n)
    Cm, pC = _out((m * k,))
    Q, pQ = _out((k * n,))
    rank = C.c_size_t()
    call("xb_cq", pC, pQ, C.byref(rank), pA, m, n)
    r = rank.value
    return Cm[:m * r].reshape(m, r).copy(), Q[:r * n].reshape(r, n).copy(), r


def solve(A, b):
    A, pA = _in(A)
    b, pb = _in(b)
    m, n = A.shape
    b2 = b.reshape(m, -1)
    x, px = _out((n, b2.shape[1]))
    call("xb_solve", px, pA, m, n, pb, b2.shape[1])
    return x.reshape((n,) + b.shape[1:])


def solve_least_squares(A, b):
    A, pA = _in(A)
    b, pb = _in(b)
    m, n = A.shape
    b2 = b.reshape(m, -1)
    x, px = _out((n, b2.shape[1]))
    call("xb_solve_least_squares", px, pA, m, n, pb, b2.shape[1])
    return x.reshape((n,) + b.shape[1:])


# ---- Tensor free functions --------------------------------------------------------------------------------------------
def contract(lhs, lhs_trans, rhs, rhs_trans, num_modes):
    """xerus::contract (src/xerus/tensor.cpp:1252-1352): matricise, one GEMM, reshape."""
    lhs = np.asarray(lhs, dtype=np.float64)
    rhs = np.asarray(rhs, dtype=np.float64)
    ld, rd = lhs.shape, rhs.shape
    if num_modes > len(ld) or num_modes > len(rd):
        raise XerusError(1, "contract: more modes to contract than the tensors have")
    keep_l, mid_l = (ld[num_modes:], ld[:num_modes]) if lhs_trans else (ld[:len(ld) - num_modes], ld[len(ld) - num_modes:])
    keep_r, mid_r = (rd[:len(rd) - num_modes], rd[len(rd) - num_modes:]) if rhs_trans else (rd[num_modes:], rd[:num_modes])
    if tuple(mid_l) != tuple(mid_r):
        raise XerusError(1, "contract: dimensions of the contracted modes do not coincide")   # tensor.cpp:1273-1279
    mid = int(np.prod(mid_l, dtype=np.int64))
    L = lhs.reshape(mid, -1) if lhs_trans else lhs.reshape(-1, mid)
    R = rhs.reshape(-1, mid) if rhs_trans else rhs.reshape(mid, -1)
    return matrix_matrix_product(1.0, L, lhs_trans, R, rhs_trans).reshape(tuple(keep_l) + tuple(keep_r))


def reshuffle(t, shuffle):
    """xerus::reshuffle (src/xerus/indexedTensor_tensor_evaluate.cpp:55-137): out mode shuffle[i] = in mode i."""
    t, pt = _in(t)
    shuffle = [int(s) for s in shuffle]
    if sorted(shuffle) != list(range(t.ndim)):
        raise XerusError(1, "reshuffle: shuffle is not a permutation")
    out_shape = [0] * t.ndim
    for i, s in enumerate(shuffle):
        out_shape[s] = t.shape[i]
    out, po = _out(tuple(out_shape))
    call("xb_reshuffle", po, pt, _sizes(t.shape), _sizes(shuffle), t.ndim)
    return out


def calculate_svd(t, split_pos, max_rank=0, eps=EPSILON):
    """calculate_svd (src/xerus/tensor.cpp:1424-1489): U (..., k), S (k,), Vt (k, ...); max_rank 0 = no cap."""
    t = np.asarray(t, dtype=np.float64)
    if not (0 <= eps < 1):
        raise XerusError(1, "Epsilon must be fullfill 0 <= _eps < 1.")
    lhs = int(np.prod(t.shape[:split_pos], dtype=np.int64))
    U, S, Vt = svd(t.reshape(lhs, -1))
    rank = len(S)
    if max_rank:
        rank = min(rank, int(max_rank))
    for j in range(1, rank):
        if S[j] <= eps * S[0]:
            rank = j
            break
    return (U[:, :rank].reshape(t.shape[:split_pos] + (rank,)).copy(), S[:rank].copy(),
            Vt[:rank].reshape((rank,) + t.shape[split_pos:]).copy())
