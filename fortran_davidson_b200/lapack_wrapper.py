"""Mirror of the reference's `lapack_wrapper` module (src/lapack_wrapper.f90): same call forms,
computed by the device kernels of libdavidson_b200.so (Jacobi eigensolver, CholeskyQR2, SIMT DGEMM)."""
import ctypes as C

import numpy as np

from ._lib import check, dp, lib


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def lapack_generalized_eigensolver(mtx, stx=None):
    """lapack_generalized_eigensolver(mtx, eigenvalues, eigenvectors [, stx]) (lapack_wrapper.f90:14-91)."""
    mtx = _f(mtx)
    n = mtx.shape[0]
    stx_f = _f(stx) if stx is not None else None
    w = np.zeros(n)
    v = np.zeros((n, n), order="F")
    check(lib().dav_lapack_generalized_eigensolver(C.c_int(n), dp(mtx), dp(stx_f), dp(w), dp(v)))
    return w, v


def sym_eigh_info(mtx, reps=0):
    """Diagnostic of the Rayleigh-Ritz eigensolver (csrc/trideig.cu): returns (w, v, info, ms) with
    info = {accepted, smax, orth_defect, residual}; accepted = -1 when the size goes to Jacobi directly."""
    mtx = _f(mtx)
    n = mtx.shape[0]
    w = np.zeros(n)
    v = np.zeros((n, n), order="F")
    info = np.zeros(8)
    ms = (C.c_float * max(reps, 1))()
    check(lib().dav_sym_eigh_info(C.c_int(n), dp(mtx), dp(w), dp(v), dp(info), C.c_int(reps), ms))
    return w, v, {"accepted": int(info[0]), "smax": info[1], "orth_defect": info[2], "residual": info[3], "prof": list(info[4:8])}, list(ms)[:reps]


def lapack_generalized_eigensolver_lowest(mtx, stx, lowest):
    """lapack_generalized_eigensolver_lowest (lapack_wrapper.f90:93-174)."""
    mtx, stx = _f(mtx), _f(stx)
    n = mtx.shape[0]
    w = np.zeros(lowest)
    v = np.zeros((n, lowest), order="F")
    check(lib().dav_lapack_generalized_eigensolver_lowest(C.c_int(n), dp(mtx), dp(stx), C.c_int(lowest), dp(w),
                                                          dp(v)))
    return w, v


def lapack_qr(basis):
    """lapack_qr(basis) (lapack_wrapper.f90:176-236); returns the orthonormalised copy."""
    q = np.array(basis, dtype=np.float64, order="F", copy=True)
    check(lib().dav_lapack_qr(C.c_int64(q.shape[0]), C.c_int(q.shape[1]), dp(q), C.c_int64(q.shape[0])))
    return q


def lapack_solver(arr, brr):
    """lapack_solver(arr, brr) (lapack_wrapper.f90:238-277); returns x."""
    a = _f(arr)
    b = np.array(brr, dtype=np.float64, copy=True).reshape(-1)
    check(lib().dav_lapack_solver(C.c_int(a.shape[0]), dp(a), dp(b)))
    return b


def lapack_matmul(transA, transB, arr, brr, alpha=1.0):
    """lapack_matmul(transA, transB, arr, brr [, alpha]) (lapack_wrapper.f90:279-328)."""
    arr, brr = _f(arr), _f(brr)
    m = arr.shape[1] if transA == "T" else arr.shape[0]
    n = brr.shape[0] if transB == "T" else brr.shape[1]
    out = np.zeros((m, n), order="F")
    check(lib().dav_lapack_matmul(C.c_char(transA.encode()), C.c_char(transB.encode()), C.c_int64(arr.shape[0]),
                                  C.c_int64(arr.shape[1]), dp(arr), C.c_int64(brr.shape[0]),
                                  C.c_int64(brr.shape[1]), dp(brr), C.c_double(alpha), dp(out)))
    return out


def lapack_matrix_vector(transA, mtx, vector, alpha=1.0):
    """lapack_matrix_vector(transA, mtx, vector [, alpha]) (lapack_wrapper.f90:330-364)."""
    mtx = _f(mtx)
    v = np.ascontiguousarray(vector, dtype=np.float64)
    out = np.zeros(mtx.shape[1] if transA == "T" else mtx.shape[0])
    check(lib().dav_lapack_matrix_vector(C.c_char(transA.encode()), C.c_int64(mtx.shape[0]),
                                         C.c_int64(mtx.shape[1]), dp(mtx), dp(v), C.c_double(alpha), dp(out)))
    return out


def lapack_sort(id_, vector):
    """lapack_sort(id, vector) (lapack_wrapper.f90:367-392): returns (sorted vector, 1-based keys)."""
    v = np.array(vector, dtype=np.float64, copy=True)
    keys = np.zeros(v.size, dtype=np.int32)
    check(lib().dav_lapack_sort(C.c_char(id_.encode()), C.c_int64(v.size), dp(v),
                                keys.ctypes.data_as(C.POINTER(C.c_int32))))
    return v, keys
