"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Mirrors the reference's Fortran call forms (davidson.f90:51-83, :277-312; array_utils.f90;
lapack_wrapper.f90) on numpy arrays in Fortran order.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg may import this module.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

METHODS = {"DPR": 0, "GJD": 1}
OP_BENCHMARK_MTX, OP_IDENTITY, OP_TEST_MTX, OP_TEST_STX, OP_DENSE = 0, 1, 2, 3, 4

_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        _lib = C.CDLL(path)
        _lib.orc_norm.restype = C.c_double
        _lib.orc_uniform01.restype = C.c_double
        _lib.orc_uniform01.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        _lib.orc_get_num_threads.restype = C.c_int
    return _lib


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(n))


def get_num_threads():
    return lib().orc_get_num_threads()


# ---- lapack_wrapper.f90 -------------------------------------------------------------------------
def lapack_generalized_eigensolver(mtx, stx=None):
    mtx = _f(mtx)
    n = mtx.shape[0]
    stx_f = _f(stx) if stx is not None else None
    w = np.zeros(n)
    v = np.zeros((n, n), order="F")
    info = lib().orc_lapack_generalized_eigensolver(C.c_int(n), _p(mtx), _p(stx_f), _p(w), _p(v))
    if info:
        raise RuntimeError("lapack_generalized_eigensolver failed, info=%d" % info)
    return w, v


def lapack_generalized_eigensolver_lowest(mtx, stx, lowest):
    mtx, stx = _f(mtx), _f(stx)
    n = mtx.shape[0]
    w = np.zeros(lowest)
    v = np.zeros((n, lowest), order="F")
    info = lib().orc_lapack_generalized_eigensolver_lowest(C.c_int(n), _p(mtx), _p(stx), C.c_int(lowest), _p(w),
                                                           _p(v))
    if info:
        raise RuntimeError("lapack_generalized_eigensolver_lowest failed, info=%d" % info)
    return w, v


def lapack_qr(basis):
    q = np.array(basis, dtype=np.float64, order="F", copy=True)
    info = lib().orc_lapack_qr(C.c_int(q.shape[0]), C.c_int(q.shape[1]), _p(q))
    if info:
        raise RuntimeError("lapack_qr failed, info=%d" % info)
    return q


def lapack_solver(arr, brr):
    a = np.array(arr, dtype=np.float64, order="F", copy=True)
    b = np.array(brr, dtype=np.float64, order="F", copy=True).reshape(-1)
    info = lib().orc_lapack_solver(C.c_int(a.shape[0]), _p(a), _p(b))
    if info:
        raise RuntimeError("lapack_solver failed, info=%d" % info)
    return b


def lapack_matmul(transA, transB, arr, brr, alpha=1.0):
    arr, brr = _f(arr), _f(brr)
    m = arr.shape[1] if transA == "T" else arr.shape[0]
    n = brr.shape[0] if transB == "T" else brr.shape[1]
    out = np.zeros((m, n), order="F")
    lib().orc_lapack_matmul(C.c_char(transA.encode()), C.c_char(transB.encode()), C.c_int(arr.shape[0]),
                            C.c_int(arr.shape[1]), _p(arr), C.c_int(brr.shape[0]), C.c_int(brr.shape[1]), _p(brr),
                            C.c_double(alpha), _p(out))
    return out


def lapack_matrix_vector(transA, mtx, vector, alpha=1.0):
    mtx = _f(mtx)
    vector = np.ascontiguousarray(vector, dtype=np.float64)
    out = np.zeros(mtx.shape[0])
    lib().orc_lapack_matrix_vector(C.c_char(transA.encode()), C.c_int(mtx.shape[0]), C.c_int(mtx.shape[1]), _p(mtx),
                                   _p(vector), C.c_double(alpha), _p(out))
    return out


def lapack_sort(id_, vector):
    """Returns (sorted vector, 1-based keys); the reference sorts its argument in place."""
    v = np.array(vector, dtype=np.float64, copy=True)
    keys = np.zeros(v.size, dtype=np.int32)
    info = lib().orc_lapack_sort(C.c_char(id_.encode()), C.c_int(v.size), _p(v), keys.ctypes.data_as(_ip))
    if info:
        raise RuntimeError("lapack_sort failed, info=%d" % info)
    return v, keys


# ---- array_utils.f90 ----------------------------------------------------------------------------
def norm(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().orc_norm(C.c_int(v.size), _p(v))


def uniform01(seed, lo, hi):
    return lib().orc_uniform01(seed, lo, hi)


def generate_diagonal_dominant(m, sparsity, diag_val=None, seed=0):
    out = np.zeros((m, m), order="F")
    dv = C.byref(C.c_double(diag_val)) if diag_val is not None else None
    lib().orc_generate_diagonal_dominant(C.c_int(m), C.c_double(sparsity), dv, C.c_uint64(seed), _p(out))
    return out


def diagonal(matrix):
    matrix = _f(matrix)
    out = np.zeros(matrix.shape[0])
    lib().orc_diagonal(C.c_int(matrix.shape[0]), _p(matrix), _p(out))
    return out


def generate_preconditioner(diag, dim_sub):
    d = np.array(diag, dtype=np.float64, copy=True)
    out = np.zeros((d.size, dim_sub), order="F")
    lib().orc_generate_preconditioner(C.c_int(d.size), _p(d), C.c_int(dim_sub), _p(out))
    return out


# ---- matrix-free operators ----------------------------------------------------------------------
def compute_on_the_fly(op, i, dim):
    """Column i (1-based) of the on-the-fly operator `op`."""
    out = np.zeros(dim)
    lib().orc_compute_on_the_fly(C.c_int(op), C.c_int(i), C.c_int(dim), _p(out))
    return out


def operator_matrix(op, dim):
    return np.asfortranarray(np.stack([compute_on_the_fly(op, j, dim) for j in range(1, dim + 1)], axis=1))


def free_matmul(op, array, dense=None):
    array = _f(array)
    out = np.zeros_like(array, order="F")
    dense_f = _f(dense) if dense is not None else None
    lib().orc_free_matmul(C.c_int(op), C.c_int(array.shape[0]), C.c_int(array.shape[1]), _p(array), _p(out),
                          _p(dense_f))
    return out


# ---- davidson.f90 -------------------------------------------------------------------------------
class Result:
    def __init__(self, eigenvalues, eigenvectors, iters, trace_k, trace_err):
        self.eigenvalues = eigenvalues
        self.eigenvectors = eigenvectors
        self.iters = iters
        self.trace_k = trace_k
        self.trace_err = trace_err


def generalized_eigensolver(matrix, lowest, method, max_iterations, tolerance, max_dim_sub=None,
                            second_matrix=None):
    """generalized_eigensolver_dense (davidson.f90:51-246)."""
    matrix = _f(matrix)
    n = matrix.shape[0]
    second = _f(second_matrix) if second_matrix is not None else None
    ev = np.zeros(lowest)
    vec = np.zeros((n, lowest), order="F")
    iters = C.c_int(0)
    cap = max_iterations + 1
    tk = np.zeros(cap, dtype=np.int32)
    te = np.zeros(cap)
    info = lib().orc_generalized_eigensolver_dense(
        C.c_int(n), _p(matrix), _p(second), C.c_int(lowest), C.c_int(METHODS[method]), C.c_int(max_iterations),
        C.c_double(tolerance), C.c_int(max_dim_sub or 0), _p(ev), _p(vec), C.byref(iters), tk.ctypes.data_as(_ip),
        _p(te), C.c_int(cap))
    if info:
        raise RuntimeError("generalized_eigensolver_dense failed, info=%d" % info)
    nit = min(iters.value, max_iterations)
    return Result(ev, vec, iters.value, tk[:nit].copy(), te[:nit].copy())


def generalized_eigensolver_free(dim, op_a, op_b, lowest, method, max_iterations, tolerance, max_dim_sub=None,
                                 dense_a=None, dense_b=None):
    """generalized_eigensolver_free (davidson.f90:277-460)."""
    da = _f(dense_a) if dense_a is not None else None
    db = _f(dense_b) if dense_b is not None else None
    ev = np.zeros(lowest)
    vec = np.zeros((dim, lowest), order="F")
    iters = C.c_int(0)
    cap = max_iterations + 1
    tk = np.zeros(cap, dtype=np.int32)
    te = np.zeros(cap)
    info = lib().orc_generalized_eigensolver_free(
        C.c_int(dim), C.c_int(op_a), _p(da), C.c_int(op_b), _p(db), C.c_int(lowest), C.c_int(METHODS[method]),
        C.c_int(max_iterations), C.c_double(tolerance), C.c_int(max_dim_sub or 0), _p(ev), _p(vec), C.byref(iters),
        tk.ctypes.data_as(_ip), _p(te), C.c_int(cap))
    if info:
        raise RuntimeError("generalized_eigensolver_free failed, info=%d" % info)
    nit = iters.value if iters.value > 0 else max_iterations
    return Result(ev, vec, iters.value, tk[:nit].copy(), te[:nit].copy())
