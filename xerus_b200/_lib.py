"""ctypes binding of libxb200.so — exactly the symbols include/xb200.h declares.

There is no fallback: if the shared library is missing this raises, and every entry point fails with
XB_ERR_NO_DEVICE when no CUDA device is present (the library itself enforces that).
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libxb200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "xb200.h")


class XerusError(RuntimeError):
    """Mirror of xerus::misc::generic_error (reference: include/xerus/misc/exceptions.h:37-73)."""

    def __init__(self, code, msg):
        super().__init__("xb200 error %d: %s" % (code, msg))
        self.code = code


class ALSOptions(C.Structure):
    _fields_ = [("sites", C.c_uint32), ("assume_spd", C.c_int), ("num_half_sweeps", C.c_size_t),
                ("convergence_epsilon", C.c_double), ("preserve_core_position", C.c_int),
                ("local_tolerance", C.c_double), ("local_max_iterations", C.c_size_t), ("local_solver", C.c_int)]


def declared_symbols():
    """All function names declared in include/xb200.h (used by the CPU-side symbol test)."""
    with open(HEADER) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xb_[a-z0-9_]+)\s*\(", text)))


_lib = None
P = C.POINTER
dp, sz, szp, vp = P(C.c_double), C.c_size_t, P(C.c_size_t), C.c_void_p

_SIGS = {
    "xb_init": [C.c_int], "xb_shutdown": [], "xb_synchronize": [], "xb_get_stream": [P(vp)], "xb_worker_select": [C.c_int], "xb_synchronize_all": [],
    "xb_kernel_launch_count": [P(C.c_uint64)], "xb_set_option": [C.c_char_p, C.c_double],
    "xb_profile_enable": [C.c_int], "xb_profile_get": [C.c_char_p, P(C.c_uint64), P(C.c_uint64), dp],
    "xb_alloc": [P(vp), sz], "xb_free": [vp], "xb_alloc_host": [P(vp), sz], "xb_free_host": [vp],
    "xb_upload": [vp, vp, sz], "xb_download": [vp, vp, sz], "xb_prefetch": [vp, sz], "xb_release": [vp],
    "xb_one_norm": [dp, sz, dp], "xb_two_norm": [dp, sz, dp], "xb_dot_product": [dp, sz, dp, dp],
    "xb_matrix_vector_product": [dp, sz, C.c_double, dp, sz, C.c_int, dp],
    "xb_dyadic_vector_product": [dp, sz, sz, C.c_double, dp, dp],
    "xb_matrix_matrix_product": [dp, sz, sz, C.c_double, dp, sz, C.c_int, sz, dp, sz, C.c_int],
    "xb_svd": [dp, dp, dp, dp, sz, sz],
    "xb_qc": [dp, dp, szp, dp, sz, sz], "xb_cq": [dp, dp, szp, dp, sz, sz],
    "xb_qr": [dp, dp, dp, sz, sz], "xb_rq": [dp, dp, dp, sz, sz],
    "xb_solve": [dp, dp, sz, sz, dp, sz], "xb_solve_least_squares": [dp, dp, sz, sz, dp, sz],
    "xb_reshuffle": [dp, dp, szp, szp, sz],
    "xb_dev_gemm": [vp, sz, sz, sz, C.c_double, vp, sz, C.c_int, sz, vp, sz, C.c_int, C.c_double],
    "xb_dev_qr": [vp, vp, vp, sz, sz], "xb_dev_lq": [vp, vp, vp, sz, sz],
    "xb_dev_svd": [vp, vp, vp, vp, sz, sz, sz, C.c_int, C.c_int, P(C.c_int)],
    "xb_dev_reshuffle": [vp, vp, szp, szp, sz], "xb_dev_two_norm": [vp, sz, dp],
    "xb_tt_create": [P(vp), sz, szp, szp, C.c_int], "xb_tt_destroy": [vp], "xb_tt_clone": [P(vp), vp],
    "xb_tt_degree": [vp, szp, P(C.c_int)], "xb_tt_ranks": [vp, szp], "xb_tt_dims": [vp, szp],
    "xb_tt_core_position": [vp, P(C.c_int), szp], "xb_tt_assume_core_position": [vp, sz],
    "xb_tt_set_component": [vp, sz, dp, sz, sz], "xb_tt_get_component": [vp, sz, dp],
    "xb_tt_component_size": [vp, sz, szp, szp, szp],
    "xb_tt_set_components": [vp, P(dp), szp], "xb_tt_get_components": [vp, P(dp)],
    "xb_tt_move_core": [vp, sz, C.c_int], "xb_tt_round": [vp, szp, C.c_double],
    "xb_tt_round_svals": [vp, szp, C.c_double, dp, sz], "xb_tt_round_batched": [P(vp), sz, sz, C.c_double],
    "xb_tt_apply_round_batched": [P(vp), vp, P(vp), sz, sz, C.c_double],
    "xb_tt_frob_norm": [vp, dp], "xb_tt_inner": [vp, vp, dp], "xb_tt_distance": [vp, vp, dp],
    "xb_tt_scale": [vp, C.c_double], "xb_tt_add": [P(vp), vp, vp], "xb_tt_apply": [P(vp), vp, vp],
    "xb_tt_from_dense": [P(vp), dp, sz, szp, C.c_double, sz], "xb_tt_to_dense": [vp, dp],
    "xb_tt_from_dense_ex": [P(vp), dp, sz, szp, C.c_int, C.c_double, szp], "xb_tt_soft_threshold": [vp, dp, C.c_int],
    "xb_als_default_options": [P(ALSOptions), C.c_uint32, C.c_int],
    "xb_als_solve": [vp, vp, vp, P(ALSOptions), dp, szp],
    "xb_env_apply": [vp, vp, sz, sz, P(vp), szp, sz, vp, sz, sz, vp, sz, sz],
    "xb_env_apply_rows": [vp, vp, sz, sz, P(vp), szp, sz, vp, sz, sz, vp, sz, sz],
    "xb_env_apply_rows_fused": [vp, sz, sz, P(vp), szp, sz, vp, sz, sz, vp, sz, sz, C.c_int, C.c_int, P(vp), C.c_uint, P(vp)],
    "xb_perf_enable": [C.c_int], "xb_perf_reset": [], "xb_perf_count": [szp],
    "xb_perf_entry": [sz, P(C.c_char_p), P(C.c_char_p), P(C.c_char_p), P(C.c_uint64), dp],
    "xb_peer_buffer_bytes": [sz, sz, C.c_int, szp], "xb_peer_buffer_create": [sz, P(vp), C.c_char_p],
    "xb_peer_buffer_open": [C.c_char_p, P(vp)], "xb_peer_buffer_close": [vp], "xb_peer_buffer_check": [vp], "xb_peer_buffer_destroy": [vp],
    "xb_env_apply_fused": [vp, sz, sz, P(vp), szp, sz, vp, sz, sz, vp, sz, sz, C.c_int, C.c_int, P(vp), C.c_uint, P(vp)],
    "xb_file_open": [P(vp), C.c_char_p], "xb_file_close": [vp],
    "xb_file_info": [vp, P(C.c_int), szp, P(C.c_int), szp], "xb_file_dims": [vp, szp], "xb_file_ranks": [vp, szp],
    "xb_file_read_component": [vp, sz, dp],
    "xb_file_write_tensor": [C.c_char_p, C.c_int, dp, szp, sz],
    "xb_file_write_tt": [C.c_char_p, C.c_int, sz, szp, szp, C.c_int, C.c_int, sz, P(dp)],
    "xb_tt_load": [P(vp), C.c_char_p], "xb_tt_save": [vp, C.c_char_p, C.c_int],
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libxb200.so is not built (run `python -m xerus_b200.build`); "
                              "xerus_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.xb_last_error.restype = C.c_char_p
        L.xb_last_error.argtypes = []
        L.xb_version.restype = C.c_int
        L.xb_version.argtypes = []
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise XerusError(status, lib().xb_last_error().decode(errors="replace"))


def call(name, *args):
    check(getattr(lib(), name)(*args))
