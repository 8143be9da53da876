"""SECOND, INDEPENDENT CPU restatement of the reference's Davidson loop (TEST INFRASTRUCTURE ONLY).

Written from the Fortran sources -- not from oracle/davidson_oracle.cpp -- with the LAPACK / BLAS routines the
reference calls taken from scipy.linalg.lapack / scipy.linalg.blas (the very routines: DSYEV, DSYGV, DGEQRF, DORGQR,
DSYSV, DGEMM, DGEMV).  Its only purpose: tests/test_oracle_pinning.py runs BOTH restatements on every golden case and
requires identical iteration counts, basis schedules, residual traces and eigenvectors.  Two restatements by two
routes (C++ against the raw LAPACK symbols, numpy against scipy's wrappers) that agree statement for statement are
the strongest pin the image allows: there is no Fortran compiler here, so the reference itself cannot run
("parity unpinned" until a gfortran build exists -- DESIGN.md section 2).

Each function cites the reference lines it restates (paths under /root/reference/src).
"""
import numpy as np
from scipy.linalg import blas, lapack


# ---- lapack_wrapper.f90 -----------------------------------------------------------------------------------------
def lapack_matmul(transA, transB, arr, brr, alpha=1.0):
    """lapack_wrapper.f90:279-328 -- DGEMM(transA, transB, m, n, k, alpha, arr, lda, brr, ldb, 0, mtx, m)."""
    return blas.dgemm(alpha, np.asfortranarray(arr), np.asfortranarray(brr), trans_a=(transA == "T"),
                      trans_b=(transB == "T"))


def lapack_matrix_vector(transA, mtx, vector, alpha=1.0):
    """lapack_wrapper.f90:330-364 -- DGEMV(transA, m, n, alpha, mtx, m, vector, 1, 0, rs, 1)."""
    return blas.dgemv(alpha, np.asfortranarray(mtx), np.ascontiguousarray(vector), trans=(1 if transA == "T" else 0))


def lapack_generalized_eigensolver(mtx, stx=None):
    """lapack_wrapper.f90:14-91 -- DSYEV('V','U') or DSYGV(itype=1,'V','U') on copies; ALL eigenpairs, ascending."""
    a = np.array(mtx, dtype=np.float64, order="F", copy=True)
    if stx is None:
        w, v, info = lapack.dsyev(a, compute_v=1, lower=0, overwrite_a=1)
        name = "DSYEV"
    else:
        b = np.array(stx, dtype=np.float64, order="F", copy=True)
        w, v, info = lapack.dsygv(a, b, itype=1, jobz="V", uplo="U", overwrite_a=1, overwrite_b=1)
        name = "DSYGV"
    if info != 0:  # check_lapack_call, lapack_wrapper.f90:395-408: print + error stop
        raise RuntimeError("call to subroutine: %s has failed! info: %d" % (name, info))
    return w, v


def lapack_qr(basis):
    """lapack_wrapper.f90:176-236 -- DGEQRF then DORGQR(m, n, min(m, n)): the first n columns of Q."""
    a = np.array(basis, dtype=np.float64, order="F", copy=True)
    qr, tau, _work, info = lapack.dgeqrf(a, overwrite_a=1)
    if info != 0:
        raise RuntimeError("call to subroutine: DGEQRF has failed! info: %d" % info)
    q, _work, info = lapack.dorgqr(qr, tau, overwrite_a=1)
    if info != 0:
        raise RuntimeError("call to subroutine: DORGQR has failed! info: %d" % info)
    return q


def lapack_solver(arr, brr):
    """lapack_wrapper.f90:238-277 -- DSYSV('U'); on info > 0 the zero pivot is replaced by tiny() and DSYSV runs
    again on the (already factorised, as in the reference) array."""
    a = np.array(arr, dtype=np.float64, order="F", copy=True)
    b = np.array(brr, dtype=np.float64, order="F", copy=True).reshape(-1, 1)
    udut, ipiv, x, info = lapack.dsysv(a, b, lower=0, overwrite_a=0, overwrite_b=0)
    if info > 0:
        udut[info - 1, info - 1] = np.finfo(np.float64).tiny
        udut, ipiv, x, info = lapack.dsysv(udut, b, lower=0)
        if info != 0:
            raise RuntimeError("call to subroutine: DSYSV has failed! info: %d" % info)
    elif info < 0:
        raise RuntimeError("call to subroutine: DSYSV has failed! info: %d" % info)
    return x[:, 0]


def lapack_sort(id_, vector):
    """lapack_wrapper.f90:367-392 -- DLASRT sorts `vector` in place; keys(i) = the LAST j with
    |sorted(j) - original(i)| < 1e-16 (the inner loop does not exit on the first match).  1-based keys."""
    xs = np.array(vector, dtype=np.float64, copy=True)
    srt = np.sort(xs) if id_ == "I" else np.sort(xs)[::-1]
    vector[:] = srt
    keys = np.zeros(xs.size, dtype=np.int64)
    # the comparison constant 1e-16 is a default-real literal in the reference (single precision, 1.0e-16)
    thr = float(np.float32(1e-16))
    for i in range(xs.size):
        hit = np.nonzero(np.abs(srt - xs[i]) < thr)[0]
        if hit.size:
            keys[i] = hit[-1] + 1
    return keys


# ---- array_utils.f90 ---------------------------------------------------------------------------------------------
def norm(vector):
    """array_utils.f90:46-53 -- sqrt(sum(vector ** 2))."""
    return float(np.sqrt(np.sum(np.asarray(vector) ** 2.0)))


def diagonal(matrix):
    """array_utils.f90:115-134."""
    return np.array(np.diagonal(matrix), dtype=np.float64)


def search_key(keys, i):
    """array_utils.f90:162-179 -- first j with keys(j) == i (undefined when absent; None here)."""
    hit = np.nonzero(keys == i)[0]
    return int(hit[0]) if hit.size else None


def generate_preconditioner(diag, dim_sub):
    """array_utils.f90:136-160 -- one-hot columns at the dim_sub smallest diagonal entries, ascending; `diag` is
    sorted in place like in the reference."""
    keys = lapack_sort("I", diag)
    precond = np.zeros((diag.size, dim_sub), order="F")
    for i in range(1, dim_sub + 1):
        k = search_key(keys, i)
        if k is None:
            raise RuntimeError("search_key: rank %d absent (duplicate diagonal entries; undefined in the reference)" % i)
        precond[k, i - 1] = 1.0
    return precond


def eye(m, n, alpha=1.0):
    """array_utils.f90:16-44."""
    out = np.zeros((m, n), order="F")
    np.fill_diagonal(out, alpha)
    return out


# ---- davidson.f90, dense ----------------------------------------------------------------------------------------
def compute_DPR_generalized_dense(matrix, eigenvalues, residues, second_matrix=None):
    """davidson.f90:673-698 -- r(ii,j) / (theta_j * B(ii,ii) - A(ii,ii)); no zero-denominator guard."""
    dA = np.diagonal(matrix)
    if second_matrix is not None:
        den = eigenvalues[None, :] * np.diagonal(second_matrix)[:, None] - dA[:, None]
    else:
        den = eigenvalues[None, :] - dA[:, None]
    return np.asfortranarray(residues / den)


def compute_GJD_generalized_dense(matrix, eigenvalues, ritz_vectors, residues, second_matrix=None):
    """davidson.f90:700-734 -- per Ritz pair: xs = I - u u^T, ys = A - theta B, arr = xs ys xs, DSYSV arr t = -r."""
    m = matrix.shape[0]
    correction = np.zeros((m, ritz_vectors.shape[1]), order="F")
    for k in range(ritz_vectors.shape[1]):
        rs = np.asfortranarray(ritz_vectors[:, k:k + 1])
        xs = eye(m, m) - lapack_matmul("N", "T", rs, rs)
        if second_matrix is not None:
            ys = matrix - eigenvalues[k] * second_matrix
        else:  # substract_from_diagonal, davidson.f90:736-750
            ys = np.array(matrix, order="F", copy=True)
            ys[np.arange(m), np.arange(m)] -= eigenvalues[k]
        arr = lapack_matmul("N", "N", xs, lapack_matmul("N", "N", ys, xs))
        correction[:, k] = lapack_solver(arr, -residues[:, k])
    return correction


class Result:
    def __init__(self, eigenvalues, eigenvectors, iters, trace_k, trace_err):
        self.eigenvalues, self.eigenvectors, self.iters = eigenvalues, eigenvectors, iters
        self.trace_k, self.trace_err = trace_k, trace_err


def generalized_eigensolver_dense(matrix, lowest, method, max_iterations, tolerance, max_dim_sub=None,
                                  second_matrix=None):
    """davidson.f90:51-246, statement by statement."""
    matrix = np.asfortranarray(matrix, dtype=np.float64)
    gev = second_matrix is not None                                          # :122
    if gev:
        second_matrix = np.asfortranarray(second_matrix, dtype=np.float64)
    initial_dimension = lowest * 2                                           # :108
    has_converged = np.zeros(lowest, dtype=bool)                             # :112
    max_dim = max_dim_sub if max_dim_sub else lowest * 10                    # :115-119
    d = diagonal(matrix)                                                     # :127
    V = generate_preconditioner(d, initial_dimension)                        # :128
    matrix_proj = lapack_matmul("T", "N", V, lapack_matmul("N", "N", matrix, V))        # :131
    if gev:
        second_matrix_proj = lapack_matmul("T", "N", V, lapack_matmul("N", "N", second_matrix, V))  # :134
    trace_k, trace_err = [], []
    eigenvalues = eigenvectors = None
    iters = None
    for i in range(1, max_iterations + 1):                                   # :138
        if gev:                                                              # :152-156
            eigenvalues_sub, eigenvectors_sub = lapack_generalized_eigensolver(matrix_proj, second_matrix_proj)
        else:
            eigenvalues_sub, eigenvectors_sub = lapack_generalized_eigensolver(matrix_proj)
        ritz_vectors = lapack_matmul("N", "N", V, eigenvectors_sub)          # :159
        residues = np.zeros((matrix.shape[0], V.shape[1]), order="F")
        for j in range(V.shape[1]):                                          # :163-170
            if gev:
                guess = eigenvalues_sub[j] * lapack_matrix_vector("N", second_matrix, ritz_vectors[:, j])
            else:
                guess = eigenvalues_sub[j] * ritz_vectors[:, j]
            residues[:, j] = lapack_matrix_vector("N", matrix, ritz_vectors[:, j]) - guess
        errors = np.array([norm(residues[:, j]) for j in range(lowest)])     # :173-178
        has_converged |= errors < tolerance
        trace_k.append(V.shape[1])
        trace_err.append(float(errors.max()))
        eigenvalues = eigenvalues_sub[:lowest].copy()                        # :186
        eigenvectors = np.asfortranarray(ritz_vectors[:, :lowest])           # :187
        if has_converged.all():                                              # :189-192
            iters = i
            break
        if V.shape[1] <= max_dim:                                            # :195
            if method == "DPR":
                correction = compute_DPR_generalized_dense(matrix, eigenvalues_sub, residues, second_matrix)
            elif method == "GJD":
                correction = compute_GJD_generalized_dense(matrix, eigenvalues_sub, ritz_vectors, residues,
                                                           second_matrix)
            else:
                raise ValueError("unknown method (the reference leaves the correction undefined)")
            V = lapack_qr(np.concatenate([V, correction], axis=1))           # :210-213
        else:
            V = lapack_matmul("N", "N", V, eigenvectors_sub[:, :initial_dimension])    # :218
        matrix_proj = lapack_matmul("T", "N", V, lapack_matmul("N", "N", matrix, V))   # :223
        if gev:
            second_matrix_proj = lapack_matmul("T", "N", V, lapack_matmul("N", "N", second_matrix, V))  # :226
    if iters is None:                                                        # :232-235
        iters = max_iterations + 1
    return Result(eigenvalues, eigenvectors, iters, np.array(trace_k), np.array(trace_err))


# ---- davidson.f90, matrix free ----------------------------------------------------------------------------------
def extract_diagonal_free(fun_gemv, dim):
    """davidson.f90:490-523 -- one application per unit vector."""
    out = np.zeros(dim)
    for ii in range(dim):
        tmp = np.zeros((dim, 1), order="F")
        tmp[ii, 0] = 1.0
        out[ii] = fun_gemv(tmp)[ii, 0]
    return out


def free_matmul(fun, array):
    """davidson.f90:526-569 -- matrix(i, j) = dot_product(fun(i, dim), array(:, j))  (row i taken as column i)."""
    array = np.asarray(array)
    dim1 = array.shape[0]
    out = np.zeros_like(array, order="F")
    for i in range(dim1):
        out[i, :] = fun(i + 1, dim1) @ array
    return out


def compute_DPR_free(eigenvalues, residues, diag_matrix, diag_second_matrix):
    """davidson.f90:463-488."""
    den = eigenvalues[None, :] * diag_second_matrix[:, None] - diag_matrix[:, None]
    return np.asfortranarray(residues / den)


def generalized_eigensolver_free(fun_matrix_gemv, fun_second_matrix_gemv, dim, lowest, max_iterations, tolerance,
                                 max_dim_sub=None, diag_matrix=None, diag_second_matrix=None):
    """davidson.f90:277-460 (`method` is ignored there: always DPR).  diag_* may be passed to skip the dim operator
    applications of extract_diagonal_free (same values)."""
    initial_dimension = lowest * 2                                           # :352
    max_dim = max_dim_sub if max_dim_sub else lowest * 10                    # :355-359
    if diag_matrix is None:
        diag_matrix = extract_diagonal_free(fun_matrix_gemv, dim)            # :365
    if diag_second_matrix is None:
        diag_second_matrix = extract_diagonal_free(fun_second_matrix_gemv, dim)  # :366
    copy_d = np.array(diag_matrix, copy=True)                                # :371
    V = generate_preconditioner(copy_d, initial_dimension)                   # :372
    trace_k, trace_err = [], []
    iters = 0  # intent(out), assigned only on convergence (:417)
    eigenvalues_sub = ritz_vectors = None
    for i in range(1, max_iterations + 1):                                   # :375
        matrixV = np.asfortranarray(fun_matrix_gemv(V))                      # :378-381
        second_matrixV = np.asfortranarray(fun_second_matrix_gemv(V))
        matrix_proj = lapack_matmul("T", "N", V, matrixV)
        second_matrix_proj = lapack_matmul("T", "N", V, second_matrixV)
        eigenvalues_sub, eigenvectors_sub = lapack_generalized_eigensolver(matrix_proj, second_matrix_proj)  # :394
        ritz_vectors = lapack_matmul("N", "N", V, eigenvectors_sub[:, :lowest])                              # :397
        lam = eye(V.shape[1], V.shape[1])                                    # :401-404
        lam[np.arange(V.shape[1]), np.arange(V.shape[1])] = eigenvalues_sub
        residues = lapack_matmul("N", "N", second_matrixV, eigenvectors_sub)  # :407-410
        guess = lapack_matmul("N", "N", residues, lam)
        residues = lapack_matmul("N", "N", matrixV, eigenvectors_sub) - guess
        errors = np.array([norm(residues[:, j]) for j in range(lowest)])     # :412-414
        trace_k.append(V.shape[1])
        trace_err.append(float(errors.max()))
        if (errors < tolerance).all():                                       # :416-419 (not sticky)
            iters = i
            break
        if V.shape[1] <= max_dim:                                            # :422
            correction = compute_DPR_free(eigenvalues_sub, residues, diag_matrix, diag_second_matrix)  # :428
            V = lapack_qr(np.concatenate([V, correction], axis=1))           # :431-434
        else:
            V = lapack_matmul("N", "N", V, eigenvectors_sub[:, :initial_dimension])  # :438
    return Result(eigenvalues_sub[:lowest].copy(), np.asfortranarray(ritz_vectors), iters, np.array(trace_k),
                  np.array(trace_err))


# ---- on-the-fly operators (benchmark_free.f90:38-76, tests/test_utils.f90:37-116) ----------------------------------
_expf = None


def _expf32(x32):
    """exp in SINGLE precision through the C library's expf (what `exp(real(i)/real(dim))` compiles to); numpy's own
    float32 exp is a different implementation and may differ in the last bit."""
    global _expf
    if _expf is None:
        import ctypes
        import ctypes.util
        libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        libm.expf.restype = ctypes.c_float
        libm.expf.argtypes = [ctypes.c_float]
        _expf = libm.expf
    return np.float64(_expf(float(x32)))


def _expensive_column(i, dim, use_sin):
    """expensive_function_1 / _2 (test_utils.f90:76-116; the same loop as benchmark_free.f90:50-59): entry j of
    column i; `1e-4` is a default-real (single precision) literal promoted to double."""
    x = _expf32(np.float32(i) / np.float32(dim))
    j = np.arange(1, dim + 1)
    y = np.array([_expf32(np.float32(jj) / np.float32(dim)) for jj in j])
    arg = np.where(j >= i, np.arctan2(x, y), np.arctan2(y, x))
    f = np.sin if use_sin else np.cos
    return f(np.log(np.sqrt(arg))) * np.float64(np.float32(1e-4))


def benchmark_matrix_column(i, dim):
    """compute_matrix_on_the_fly (benchmark_free.f90:38-63; identical in test_utils.f90:37-52)."""
    v = _expensive_column(i, dim, False)
    v[i - 1] += float(np.float32(i))
    return v


def identity_column(i, dim):
    """compute_stx_on_the_fly of the benchmark (benchmark_free.f90:65-76)."""
    v = np.zeros(dim)
    v[i - 1] = 1.0
    return v


def test_stx_column(i, dim):
    """compute_stx_on_the_fly of the tests (test_utils.f90:55-72): sin variant, diagonal := 1."""
    v = _expensive_column(i, dim, True)
    v[i - 1] = 1.0
    return v
