"""Mirror of the reference's misc::save_to_file / misc::load_from_file (include/xerus/misc/fileIO.h:102-164) on top of
the data-file layer of the C ABI (include/xb200.h section 4, csrc/fileio.cu): files written by xerus load here, files
written here load in xerus.  `read_file` / `write_tt_file` / `write_tensor_file` are host-side only (no device);
`save_to_file` / `load_from_file` take and return the device-resident TTTensor / TTOperator handles.
"""
import ctypes as C
import os

import numpy as np

from ._lib import XerusError, call, dp
from .blas_wrapper import _sizes


class FileFormat:
    """misc::FileFormat (fileIO.h:43)."""
    BINARY = 0
    TSV = 1


KINDS = ("Tensor", "TTTensor", "TTOperator")


def read_file(filename):
    """Parses a xerus data file on the host.  Returns a dict: kind ('Tensor' | 'TTTensor' | 'TTOperator'), dims,
    and for tensor trains ranks, canonicalized, core_position, components (list of numpy cores); for a Tensor `data`."""
    h = C.c_void_p()
    call("xb_file_open", C.byref(h), os.fsencode(filename))
    try:
        kind, nd, canon, core = C.c_int(), C.c_size_t(), C.c_int(), C.c_size_t()
        call("xb_file_info", h, C.byref(kind), C.byref(nd), C.byref(canon), C.byref(core))
        dims = (C.c_size_t * max(nd.value, 1))()
        call("xb_file_dims", h, dims)
        dims = [int(x) for x in dims[:nd.value]]
        out = {"kind": KINDS[kind.value], "dims": dims}
        if kind.value == 0:
            data = np.empty(dims, dtype=np.float64)
            if data.size:
                call("xb_file_read_component", h, 0, data.ctypes.data_as(dp))
            out["data"] = data
            return out
        N = 2 if kind.value == 2 else 1
        d = nd.value // N
        ranks = (C.c_size_t * max(d - 1, 1))()
        call("xb_file_ranks", h, ranks)
        ranks = [int(x) for x in ranks[:d - 1]]
        rk = [1] + ranks + [1]
        comps = []
        for i in range(d):
            shape = (rk[i], dims[i], dims[d + i], rk[i + 1]) if N == 2 else (rk[i], dims[i], rk[i + 1])
            c = np.empty(shape, dtype=np.float64)
            call("xb_file_read_component", h, i, c.ctypes.data_as(dp))
            comps.append(c)
        out.update(ranks=ranks, canonicalized=bool(canon.value), core_position=int(core.value), components=comps)
        return out
    finally:
        call("xb_file_close", h)


def write_tensor_file(filename, array, fmt=FileFormat.BINARY):
    a = np.ascontiguousarray(array, dtype=np.float64)
    call("xb_file_write_tensor", os.fsencode(filename), int(fmt), a.ctypes.data_as(dp), _sizes(a.shape), a.ndim)


def write_tt_file(filename, components, canonicalized=False, core_position=0, fmt=FileFormat.BINARY):
    """Writes host cores (r_l, n, r_r) (TTTensor) or (r_l, m, n, r_r) (TTOperator) as a xerus TTNetwork data file."""
    comps = [np.ascontiguousarray(c, dtype=np.float64) for c in components]
    if not comps:
        raise XerusError(1, "degree-0 tensor trains carry no cores")
    is_op = comps[0].ndim == 4
    for c in comps:
        if c.ndim != (4 if is_op else 3):
            raise XerusError(1, "Component must have degree %d. Given: %d" % (4 if is_op else 3, c.ndim))
    for a, b in zip(comps[:-1], comps[1:]):
        if a.shape[-1] != b.shape[0]:
            raise XerusError(1, "bond dimensions of neighbouring components do not coincide")
    if comps[0].shape[0] != 1 or comps[-1].shape[-1] != 1:
        raise XerusError(1, "the outer bonds of a tensor train have dimension one")
    dims = [c.shape[1] for c in comps] + ([c.shape[2] for c in comps] if is_op else [])
    ranks = [c.shape[-1] for c in comps[:-1]]
    ptrs = (dp * len(comps))(*[c.ctypes.data_as(dp) for c in comps])
    call("xb_file_write_tt", os.fsencode(filename), int(fmt), len(comps), _sizes(dims), _sizes(ranks or [1]), int(is_op),
         int(bool(canonicalized)), int(core_position), ptrs)


def save_to_file(obj, filename, fmt=FileFormat.BINARY):
    """misc::save_to_file(obj, filename, format) for TTTensor / TTOperator handles and dense arrays (as xerus::Tensor)."""
    from .tt import TTNetwork
    if isinstance(obj, TTNetwork):
        call("xb_tt_save", obj._h, os.fsencode(filename), int(fmt))
    else:
        write_tensor_file(filename, obj, fmt)


def load_from_file(filename):
    """misc::load_from_file<T>(filename): a device-resident TTTensor / TTOperator, or a numpy array for a Tensor file."""
    from .tt import TTOperator, TTTensor
    with open(filename, "rb") as f:
        first = f.readline().decode(errors="replace").strip()
    if "TTNetwork<" not in first:
        return read_file(filename)["data"]
    h = C.c_void_p()
    call("xb_tt_load", C.byref(h), os.fsencode(filename))
    return (TTOperator if "TTNetwork<true>" in first else TTTensor)(h)
