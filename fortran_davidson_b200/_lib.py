"""ctypes binding of libdavidson_b200.so (the C ABI in include/davidson_b200.h).

The library is the product: if it is missing or no sm_100 GPU is usable, calls FAIL LOUDLY --
there is no CPU / numpy / LAPACK fallback anywhere in this package.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DAV_B200_LIB: load another build of the same C ABI (diagnostics: A/B of two builds in scripts/)
LIB_PATH = os.environ.get("DAV_B200_LIB") or os.path.join(HERE, "libdavidson_b200.so")

GEMV_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64, C.c_int64, C.c_void_p)
# device functor: (d_x, ldx, d_y, ldy, n, b, row_begin, nrows, cuda_stream, ctx), raw device addresses
DEVICE_GEMV_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                             C.c_int64, C.c_void_p, C.c_void_p)

# every symbol include/davidson_b200.h declares
SYMBOLS = [
    "dav_last_error", "dav_version", "dav_device_count", "dav_set_default_device", "dav_generalized_eigensolver_dense", "dav_release_cache", "dav_upload_bytes",
    "dav_generalized_eigensolver_free", "dav_generalized_eigensolver_free_builtin", "dav_get_unique_id",
    "dav_create", "dav_create_distributed", "dav_destroy", "dav_alloc_pinned", "dav_free_pinned", "dav_partition_rows",
    "dav_matrix_generate_diagonal_dominant", "dav_matrix_upload", "dav_matrix_upload_rows",
    "dav_matrix_set_operator",
    "dav_matrix_set_callback", "dav_matrix_set_device_callback", "dav_matrix_clear", "dav_matrix_download", "dav_solve", "dav_solve_local", "dav_get_stats", "dav_set_profiling",
    "dav_set_matvec_impl", "dav_block_matvec", "dav_bench_block_matvec", "dav_generate_diagonal_dominant",
    "dav_generate_preconditioner", "dav_norm", "dav_norm_value", "dav_lapack_generalized_eigensolver",
    "dav_lapack_generalized_eigensolver_lowest", "dav_sym_eigh_info", "dav_lapack_qr", "dav_lapack_solver", "dav_lapack_matmul",
    "dav_lapack_matrix_vector", "dav_lapack_sort", "dav_free_matmul", "dav_compute_on_the_fly",
    "dav_debug_matvec_schedule", "dav_bench_fp64_pipe", "dav_debug_matvec_rect", "dav_debug_collective",
    "dav_comm_info", "dav_debug_chol_inv", "dav_debug_gemm_bench", "dav_debug_pip_small",
]


class Stats(C.Structure):
    _fields_ = [
        ("solve_ms", C.c_double), ("matvec_ms", C.c_double), ("matvec_bytes", C.c_double),
        ("matvec_flops", C.c_double), ("matvec_launches", C.c_int), ("kernel_launches", C.c_int),
        ("iterations", C.c_int), ("trace_len", C.c_int), ("trace_k", C.c_int * 64), ("trace_err", C.c_double * 64),
        ("last_matvec_b", C.c_int), ("last_matvec_ms", C.c_double), ("rr_ms", C.c_double), ("orth_ms", C.c_double),
        ("resid_ms", C.c_double), ("proj_ms", C.c_double), ("init_ms", C.c_double),
        ("gjd_inner_iterations", C.c_int), ("gather_ms", C.c_double), ("output_ms", C.c_double),
        ("comm_ms", C.c_double), ("collectives", C.c_int), ("spans_dropped", C.c_int), ("peer_transport", C.c_int),
        ("pip_fallbacks", C.c_int),
    ]


class DavidsonError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("davidson_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Loads the shared library (building nothing: run `python -m fortran_davidson_b200.build` or
    __graft_entry__.build() first)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -m fortran_davidson_b200.build` "
                              "(there is no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.dav_last_error.restype = C.c_char_p
        _lib.dav_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    return _lib


def check(code):
    if code != 0:
        raise DavidsonError(code, lib().dav_last_error().decode(errors="replace"))


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None
