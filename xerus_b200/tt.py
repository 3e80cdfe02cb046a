"""Host mirror of the reference's TTTensor / TTOperator (include/xerus/ttNetwork.h:44-519) for the hot path: the
objects are thin handles on device-resident tensor trains (xb_tt in include/xb200.h); every method is one C-ABI call.
Names and argument meaning follow the reference (and its Python binding, src/xerus/python/ttnetwork.cpp:32-95).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import XerusError, call
from .blas_wrapper import EPSILON, _sizes


def reduce_to_maximal_ranks(ranks, dims):
    """TTNetwork::reduce_to_maximal_ranks (src/xerus/ttNetwork.cpp:370-402) — host-side index bookkeeping."""
    ranks = [int(r) for r in ranks]
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


class TTNetwork:
    is_operator = False

    def __init__(self, handle):
        self._h = handle

    # -- construction ----------------------------------------------------------------------------------------------
    @classmethod
    def _create(cls, dims, ranks):
        d = len(dims) // (2 if cls.is_operator else 1)
        h = C.c_void_p()
        call("xb_tt_create", C.byref(h), d, _sizes(dims), _sizes(ranks), int(cls.is_operator))
        return cls(h)

    @classmethod
    def from_cores(cls, cores, core_position=None):
        """Builds a TT from host cores (r_l, n, r_r) / (r_l, m, n, r_r) via set_component."""
        cores = [np.ascontiguousarray(c, dtype=np.float64) for c in cores]
        want = 4 if cls.is_operator else 3
        for c in cores:
            if c.ndim != want:
                raise XerusError(1, "Component must have degree %d. Given: %d" % (want, c.ndim))   # ttNetwork.cpp:478
        for a, b in zip(cores[:-1], cores[1:]):
            if a.shape[-1] != b.shape[0]:
                raise XerusError(1, "bond dimensions of neighbouring components do not coincide")
        if cls.is_operator:
            dims = [c.shape[1] for c in cores] + [c.shape[2] for c in cores]
        else:
            dims = [c.shape[1] for c in cores]
        ranks = [c.shape[-1] for c in cores[:-1]]
        t = cls._create(dims, ranks)
        if cores[0].shape[0] != 1 or cores[-1].shape[-1] != 1:
            raise XerusError(1, "the outer bonds of a tensor train have dimension one")
        ptrs = (_lib.dp * len(cores))(*[c.ctypes.data_as(_lib.dp) for c in cores])
        call("xb_tt_set_components", t._h, ptrs, _sizes(ranks))
        if core_position is not None:
            t.assume_core_position(core_position)
        return t

    @classmethod
    def random(cls, dimensions, ranks, rng=None):
        """TTNetwork::random (include/xerus/ttNetwork.h:129-155): i.i.d. N(0,1) components after
        reduce_to_maximal_ranks, then move_core(0).  `rng`: numpy Generator (RNG streams are not the reference's)."""
        rng = rng or np.random.default_rng()
        dimensions = [int(n) for n in dimensions]
        N = 2 if cls.is_operator else 1
        d = len(dimensions) // N
        if np.isscalar(ranks):
            ranks = [int(ranks)] * (d - 1)
        if len(ranks) + 1 != d:
            raise XerusError(1, "Non-matching amount of ranks given to TTNetwork::random.")
        if any(r == 0 for r in ranks) or any(n == 0 for n in dimensions):
            raise XerusError(1, "rank or dimension 0 is illegal")
        ext = [dimensions[i] * (dimensions[d + i] if cls.is_operator else 1) for i in range(d)]
        rk = [1] + reduce_to_maximal_ranks(ranks, ext) + [1]
        cores = []
        for i in range(d):
            shape = (rk[i], dimensions[i], dimensions[d + i], rk[i + 1]) if cls.is_operator else (rk[i], dimensions[i], rk[i + 1])
            cores.append(rng.standard_normal(shape))
        t = cls.from_cores(cores)
        t.move_core(0)
        return t

    @classmethod
    def ones(cls, dimensions):
        if cls.is_operator:
            d = len(dimensions) // 2
            cores = [np.ones((1, dimensions[i], dimensions[d + i], 1)) for i in range(d)]
        else:
            cores = [np.ones((1, n, 1)) for n in dimensions]
        t = cls.from_cores(cores)
        t.canonicalize_left()                      # ttNetwork.cpp:189
        return t

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().xb_tt_destroy(h)
            except Exception:
                pass
            self._h = None

    def copy(self):
        h = C.c_void_p()
        call("xb_tt_clone", C.byref(h), self._h)
        return type(self)(h)

    # -- structure -------------------------------------------------------------------------------------------------
    def degree(self):
        d = C.c_size_t()
        op = C.c_int()
        call("xb_tt_degree", self._h, C.byref(d), C.byref(op))
        return d.value * (2 if op.value else 1)

    @property
    def num_components(self):
        d = C.c_size_t()
        call("xb_tt_degree", self._h, C.byref(d), None)
        return d.value

    @property
    def dimensions(self):
        n = self.degree()
        buf = (C.c_size_t * n)()
        call("xb_tt_dims", self._h, buf)
        return list(buf)

    def ranks(self):
        d = self.num_components
        buf = (C.c_size_t * max(1, d - 1))()
        call("xb_tt_ranks", self._h, buf)
        return list(buf)[:d - 1]

    def rank(self, i):
        return self.ranks()[i]

    @property
    def canonicalized(self):
        c = C.c_int()
        call("xb_tt_core_position", self._h, C.byref(c), None)
        return bool(c.value)

    @property
    def corePosition(self):
        p = C.c_size_t()
        call("xb_tt_core_position", self._h, None, C.byref(p))
        return p.value

    def assume_core_position(self, pos):
        call("xb_tt_assume_core_position", self._h, int(pos))

    def get_component(self, idx):
        rl, ext, rr = C.c_size_t(), C.c_size_t(), C.c_size_t()
        call("xb_tt_component_size", self._h, int(idx), C.byref(rl), C.byref(ext), C.byref(rr))
        dims = self.dimensions
        d = self.num_components
        shape = (rl.value, dims[idx], dims[d + idx], rr.value) if self.is_operator else (rl.value, dims[idx], rr.value)
        out = np.empty(shape, dtype=np.float64)
        call("xb_tt_get_component", self._h, int(idx), out.ctypes.data_as(_lib.dp))
        return out

    component = get_component

    def set_component(self, idx, core):
        core = np.ascontiguousarray(core, dtype=np.float64)
        dims = self.dimensions
        d = self.num_components
        want = (dims[idx], dims[d + idx]) if self.is_operator else (dims[idx],)
        if core.ndim != len(want) + 2 or tuple(core.shape[1:-1]) != tuple(want):
            raise XerusError(1, "set_component: component has the wrong external dimensions")
        call("xb_tt_set_component", self._h, int(idx), core.ctypes.data_as(_lib.dp), core.shape[0], core.shape[-1])

    def cores(self):
        """All components as host arrays (one call, one synchronisation)."""
        d = self.num_components
        dims = self.dimensions
        rk = [1] + self.ranks() + [1]
        out = [np.empty((rk[i], dims[i], dims[d + i], rk[i + 1]) if self.is_operator else (rk[i], dims[i], rk[i + 1]), dtype=np.float64)
               for i in range(d)]
        ptrs = (_lib.dp * d)(*[c.ctypes.data_as(_lib.dp) for c in out])
        call("xb_tt_get_components", self._h, ptrs)
        return out

    # -- hot path ----------------------------------------------------------------------------------------------------
    def move_core(self, position, keepRank=False):
        call("xb_tt_move_core", self._h, int(position), int(bool(keepRank)))

    def canonicalize_left(self):
        self.move_core(0)

    def canonicalize_right(self):
        self.move_core(self.num_components - 1)

    def round(self, arg, eps=None, return_svals=False):
        """TTNetwork::round overloads (src/xerus/ttNetwork.cpp:644-684):
        round(int maxRank) -> eps = EPSILON; round(float eps) -> no rank cap; round(list maxRanks, eps=EPSILON)."""
        d = self.num_components
        if isinstance(arg, (float, np.floating)) and eps is None:
            max_ranks, e = [0] * (d - 1), float(arg)
        elif isinstance(arg, (int, np.integer)):
            if arg <= 0:
                raise XerusError(1, "MaxRank must be positive")                 # ttNetwork.cpp:676
            max_ranks, e = [int(arg)] * (d - 1), EPSILON if eps is None else float(eps)
        else:
            max_ranks, e = [int(a) for a in arg], EPSILON if eps is None else float(eps)
            if len(max_ranks) + 1 != d:
                raise XerusError(1, "There must be exactly degree/N-1 maxRanks.")   # ttNetwork.cpp:648
            if any(r == 0 for r in max_ranks):
                raise XerusError(1, "Trying to round a TTTensor to rank 0 is not possible.")
        if not e < 1:
            raise XerusError(1, "_eps must be smaller than one.")
        if return_svals:
            stride = max([1] + [min(a, b) for a, b in zip([1] + self.ranks(), self.ranks() + [1])] + self.ranks())
            sv = np.zeros((max(1, d - 1), stride))
            call("xb_tt_round_svals", self._h, _sizes(max_ranks), e, sv.ctypes.data_as(_lib.dp), stride)
            rk = self.ranks()
            return [sv[i, :rk[i]].copy() for i in range(d - 1)]
        call("xb_tt_round", self._h, _sizes(max_ranks), e)

    def soft_threshold(self, tau, preventZero=False):
        """TTNetwork::soft_threshold (src/xerus/ttNetwork.cpp:688-713): scalar tau, or one tau per edge — taus[i] belongs to the
        i-th edge from the right, as in the reference (:700)."""
        d = self.num_components
        taus = [float(tau)] * (d - 1) if np.isscalar(tau) else [float(t) for t in tau]
        if len(taus) + 1 != d:
            raise XerusError(1, "There must be exactly degree/N-1 taus.")          # ttNetwork.cpp:690
        arr = (C.c_double * max(1, d - 1))(*taus)
        call("xb_tt_soft_threshold", self._h, arr, int(bool(preventZero)))

    def frob_norm(self):
        r = C.c_double()
        call("xb_tt_frob_norm", self._h, C.byref(r))
        return r.value

    def inner(self, other):
        r = C.c_double()
        call("xb_tt_inner", self._h, other._h, C.byref(r))
        return r.value

    def distance(self, other):
        """||self - other||_F without cancellation."""
        r = C.c_double()
        call("xb_tt_distance", self._h, other._h, C.byref(r))
        return r.value

    def __imul__(self, factor):
        call("xb_tt_scale", self._h, float(factor))
        return self

    def __mul__(self, factor):
        t = self.copy()
        t *= factor
        return t

    __rmul__ = __mul__

    def __add__(self, other):
        h = C.c_void_p()
        call("xb_tt_add", C.byref(h), self._h, other._h)
        return type(self)(h)

    def __sub__(self, other):
        return self + (other * -1.0)

    def to_dense(self):
        """Tensor(tt) (src/xerus/tensorNetwork.cpp:287-306); operators come out as (m_1..m_d, n_1..n_d)."""
        out = np.empty(tuple(self.dimensions), dtype=np.float64)
        call("xb_tt_to_dense", self._h, out.ctypes.data_as(_lib.dp))
        return out


def _from_dense(cls, full, eps, max_ranks):
    """TT-SVD constructor TTNetwork(Tensor, eps, maxRanks) (src/xerus/ttNetwork.cpp:112-160); max_ranks: 0 / None = unlimited,
    an int for every bond, or one entry per bond."""
    full = np.ascontiguousarray(full, dtype=np.float64)
    N = 2 if cls.is_operator else 1
    if full.ndim % N:
        raise XerusError(1, "Number of indicis must be even for TTOperator")           # ttNetwork.cpp:113
    d = full.ndim // N
    if max_ranks is None or (np.isscalar(max_ranks) and int(max_ranks) == 0):
        mr = None
    elif np.isscalar(max_ranks):
        mr = _sizes([int(max_ranks)] * max(1, d - 1))
    else:
        if len(max_ranks) != d - 1:
            raise XerusError(1, "We need %d ranks but %d where given" % (d - 1, len(max_ranks)))   # :115
        mr = _sizes([int(r) for r in max_ranks] or [1])
    h = C.c_void_p()
    call("xb_tt_from_dense_ex", C.byref(h), full.ctypes.data_as(_lib.dp), d, _sizes(full.shape), int(cls.is_operator), float(eps), mr)
    return cls(h)


class TTTensor(TTNetwork):
    is_operator = False

    @classmethod
    def from_dense(cls, full, eps=EPSILON, max_rank=0):
        return _from_dense(cls, full, eps, max_rank)


class TTOperator(TTNetwork):
    is_operator = True

    @classmethod
    def from_dense(cls, full, eps=EPSILON, max_rank=0):
        """TTOperator(Tensor, eps, maxRanks): `full` has modes (m_1..m_d, n_1..n_d) (ttNetwork.cpp:129-135)."""
        return _from_dense(cls, full, eps, max_rank)

    def apply(self, x):
        """y(i&0) = A(i/2, j/2) * x(j&0)  (src/xerus/ttNetwork.cpp:889-967, src/xerus/ttStack.cpp:197-300)."""
        h = C.c_void_p()
        call("xb_tt_apply", C.byref(h), self._h, x._h)
        return TTTensor(h)

    def __matmul__(self, x):
        return self.apply(x)

    @classmethod
    def laplace(cls, d, n):
        """Rank-2 Laplace-like operator of the BASELINE configs."""
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
        return cls.from_cores(cores)


def round_batched(tts, max_rank, eps=EPSILON):
    """Batch of independent roundings (BASELINE config 5)."""
    arr = (C.c_void_p * len(tts))(*[t._h for t in tts])
    call("xb_tt_round_batched", arr, len(tts), int(max_rank), float(eps))


def apply_round_batched(A, xs, max_rank, eps=EPSILON):
    """The item of BASELINE config 5 for a whole batch in one call: [round(A x, max_rank) for x in xs]; the items run concurrently
    on library-owned threads / streams."""
    n = len(xs)
    out = (C.c_void_p * n)()
    arr = (C.c_void_p * n)(*[x._h for x in xs])
    call("xb_tt_apply_round_batched", out, A._h, arr, n, int(max_rank), float(eps))
    return [TTTensor(C.c_void_p(h)) for h in out]
