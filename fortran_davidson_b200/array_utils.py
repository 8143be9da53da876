"""Mirror of the reference's `array_utils` module (src/array_utils.f90) on the device library."""
import ctypes as C

import numpy as np

from ._lib import check, dp, lib


def generate_diagonal_dominant(m, sparsity, diag_val=None, seed=0):
    """generate_diagonal_dominant(m, sparsity [, diag_val]) (array_utils.f90:86-113), generated on the
    GPU with the counter-based stream keyed on (seed, min(i,j), max(i,j))."""
    arr = np.zeros((m, m), order="F")
    dv = C.byref(C.c_double(diag_val)) if diag_val is not None else None
    check(lib().dav_generate_diagonal_dominant(C.c_int64(m), C.c_double(sparsity), dv, C.c_uint64(seed), dp(arr),
                                               C.c_int64(m)))
    return arr


def generate_preconditioner(diag, dim_sub):
    """generate_preconditioner(diag, dim_sub) (array_utils.f90:136-160)."""
    d = np.ascontiguousarray(diag, dtype=np.float64)
    out = np.zeros((d.size, dim_sub), order="F")
    check(lib().dav_generate_preconditioner(C.c_int64(d.size), dp(d), C.c_int(dim_sub), dp(out), C.c_int64(d.size)))
    return out


def norm(vector):
    """norm(vector) (array_utils.f90:46-53)."""
    v = np.ascontiguousarray(vector, dtype=np.float64)
    out = C.c_double(0.0)
    check(lib().dav_norm(C.c_int64(v.size), dp(v), C.byref(out)))
    return out.value


def diagonal(matrix):
    """diagonal(matrix) (array_utils.f90:115-134) -- pure indexing, no arithmetic."""
    return np.array(np.diagonal(matrix), dtype=np.float64)


def eye(m, n, alpha=1.0):
    """eye(m, n [, alpha]) (array_utils.f90:16-44)."""
    out = np.zeros((m, n), order="F")
    np.fill_diagonal(out, alpha)
    return out


def concatenate(arr, brr):
    """concatenate(arr, brr) (array_utils.f90:55-84): column-wise append (returns the new array)."""
    return np.asfortranarray(np.hstack([arr, brr]))
