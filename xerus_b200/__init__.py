"""xerus_b200 — B200-native (sm_100a) implementation of xerus's tensor-train hot path behind a C ABI.

Host-side mirror of the reference interface for that path; all arithmetic runs in libxb200.so (hand-written CUDA).
"""
import os as _os

# hardware work queues of the CUDA context (read once, when the context is created): with the default of 8 the worker streams of
# the batched entry points share queues and independent TTs serialise behind each other (batches).  xb_init() does
# the same for C callers; a process that has already created its context keeps what it had.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from ._lib import XerusError, lib, declared_symbols, LIB_PATH      # noqa: F401,E402
from . import blas_wrapper as blasWrapper                          # noqa: F401
from .blas_wrapper import contract, reshuffle, calculate_svd, EPSILON   # noqa: F401
from .tt import TTTensor, TTOperator, TTNetwork, round_batched, apply_round_batched, reduce_to_maximal_ranks   # noqa: F401
from .als import ALSVariant, ALS, ALS_SPD, DMRG, DMRG_SPD, ASD, ASD_SPD   # noqa: F401
from .file_io import FileFormat, save_to_file, load_from_file, read_file, write_tt_file, write_tensor_file   # noqa: F401


def init(device=0):
    """Selects the CUDA device and creates the library stream (idempotent)."""
    from ._lib import call
    call("xb_init", int(device))


def set_option(key, value):
    """Tuning / diagnostic knobs of the library: svd_max_sweeps, svd_persistent, als_direct_max, ..."""
    from ._lib import call
    call("xb_set_option", key.encode(), float(value))


def synchronize():
    from ._lib import call
    call("xb_synchronize")


def kernel_launch_count():
    import ctypes as C
    from ._lib import call
    n = C.c_uint64()
    call("xb_kernel_launch_count", C.byref(n))
    return n.value


def stream_handle():
    """cudaStream_t (as int) on which all xb200 work is enqueued — for CUDA-event timing on that stream."""
    import ctypes as C
    from ._lib import call
    s = C.c_void_p()
    call("xb_get_stream", C.byref(s))
    return s.value


def profile_enable(on=True):
    """CUDA-event timing per kernel class inside the library (bench/roofline only)."""
    from ._lib import call
    call("xb_profile_enable", int(bool(on)))


def profile_get(kernel_class):
    """(scopes, launches, milliseconds) accumulated for a kernel class since profile_enable(True)."""
    import ctypes as C
    from ._lib import call
    s, l, ms = C.c_uint64(), C.c_uint64(), C.c_double()
    call("xb_profile_get", kernel_class.encode(), C.byref(s), C.byref(l), C.byref(ms))
    return s.value, l.value, ms.value


def perf_enable(on=True):
    """The reference's XERUS_PERFORMANCE_ANALYSIS registry at the C ABI (misc/performanceAnalysis.h:30-39): every call is recorded
    under (group, name, shape) with the reference's own strings."""
    from ._lib import call
    call("xb_perf_enable", int(bool(on)))


def perf_reset():
    from ._lib import call
    call("xb_perf_reset")


def perf_entries():
    """{(group, name, shape): (calls, microseconds)} — the analogue of misc::performanceAnalysis::calls."""
    import ctypes as C
    from ._lib import call
    n = C.c_size_t()
    call("xb_perf_count", C.byref(n))
    out = {}
    for i in range(n.value):
        g, nm, sh, calls, us = C.c_char_p(), C.c_char_p(), C.c_char_p(), C.c_uint64(), C.c_double()
        call("xb_perf_entry", i, C.byref(g), C.byref(nm), C.byref(sh), C.byref(calls), C.byref(us))
        out[(g.value.decode(), nm.value.decode(), sh.value.decode())] = (calls.value, us.value)
    return out


def perf_analysis():
    """Text report in the spirit of misc::performanceAnalysis::get_analysis() (performanceAnalysis.cpp:34-82)."""
    ent = perf_entries()
    lines = []
    groups = sorted({k[0] for k in ent})
    total = sum(v[1] for v in ent.values()) or 1.0
    for g in groups:
        gt = sum(v[1] for k, v in ent.items() if k[0] == g)
        lines.append("%s: %.3f ms (%.1f %%)" % (g, gt / 1e3, 100 * gt / total))
        for nm in sorted({k[1] for k in ent if k[0] == g}):
            nt = sum(v[1] for k, v in ent.items() if k[:2] == (g, nm))
            nc = sum(v[0] for k, v in ent.items() if k[:2] == (g, nm))
            lines.append("  %-32s %8d calls %12.3f ms" % (nm, nc, nt / 1e3))
            for k, v in sorted(((k, v) for k, v in ent.items() if k[:2] == (g, nm)), key=lambda kv: -kv[1][1]):
                lines.append("      %-40s %8d calls %12.3f ms" % (k[2], v[0], v[1] / 1e3))
    return "\n".join(lines)


def worker_select(worker):
    """Binds the calling host thread to a private worker (stream + scratch); see include/xb200.h."""
    from ._lib import call
    call("xb_worker_select", int(worker))


def synchronize_all():
    from ._lib import call
    call("xb_synchronize_all")
