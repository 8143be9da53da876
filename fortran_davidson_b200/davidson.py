"""Mirror of the reference's `davidson` module (src/davidson.f90:586-627): the generic
`generalized_eigensolver` over a dense matrix or a pair of operator callbacks, `eigensolver`
(README.md:20), `free_matmul` (davidson.f90:526-569), plus `DavidsonSolver`, the device-resident
handle the benchmarks use (matrices stay in HBM, row-block sharded over the ranks)."""
import ctypes as C

import numpy as np

from ._lib import DEVICE_GEMV_FN, GEMV_FN, Stats, check, dp, lib

METHODS = {"DPR": 0, "GJD": 1}
OP_BENCHMARK_MTX, OP_IDENTITY, OP_TEST_MTX, OP_TEST_STX = 0, 1, 2, 3
MATVEC_AUTO, MATVEC_SIMT, MATVEC_TMA_DMMA = 0, 1, 2


def _wrap_callback(fun):
    """Adapts fun(input_vect(n, b)) -> output_vect(n, b) (davidson.f90:317-325) to the C callback."""

    def cb(xp, yp, n, b, _ctx):
        x = np.ctypeslib.as_array(xp, shape=(b, n)).T  # column-major n x b view
        y = np.ctypeslib.as_array(yp, shape=(b, n)).T
        y[:, :] = np.asarray(fun(x), dtype=np.float64)

    return GEMV_FN(cb)


def generalized_eigensolver(matrix, lowest, method, max_iterations, tolerance, max_dim_sub=None, second_matrix=None,
                            fun_second_matrix_gemv=None, dim=None):
    """call generalized_eigensolver(matrix, eigenvalues, eigenvectors, lowest, method, max_iterations,
    tolerance, iters [, max_dim_sub] [, second_matrix])            (dense,  davidson.f90:51-83)
    call generalized_eigensolver(fun_matrix_gemv, eigenvalues, ritz_vectors, lowest, method, max_iterations,
    tolerance, iters, max_dim_sub, fun_second_matrix_gemv)          (free,   davidson.f90:277-312)

    Dispatches on the type of the first argument like the Fortran generic interface
    (davidson.f90:601-625).  Returns (eigenvalues, eigenvectors, iters); in the matrix-free form
    `iters` is None when the loop did not converge (the reference leaves it unassigned)."""
    L = lib()
    if callable(matrix):
        if fun_second_matrix_gemv is None or dim is None:
            raise TypeError("matrix-free form needs fun_second_matrix_gemv and dim")
        cb_a, cb_b = _wrap_callback(matrix), _wrap_callback(fun_second_matrix_gemv)
        ev = np.zeros(lowest)
        vec = np.zeros((dim, lowest), order="F")
        iters = C.c_int(-1)
        check(L.dav_generalized_eigensolver_free(C.c_int64(dim), cb_a, None, cb_b, None, None, None, C.c_int(lowest),
                                                 method.encode(), C.c_int(max_iterations), C.c_double(tolerance),
                                                 C.c_int(max_dim_sub or 0), dp(ev), dp(vec), C.c_int64(dim),
                                                 C.byref(iters)))
        return ev, vec, (iters.value if iters.value >= 0 else None)
    a = np.asfortranarray(matrix, dtype=np.float64)
    n = a.shape[0]
    b = np.asfortranarray(second_matrix, dtype=np.float64) if second_matrix is not None else None
    ev = np.zeros(lowest)
    vec = np.zeros((n, lowest), order="F")
    iters = C.c_int(0)
    check(L.dav_generalized_eigensolver_dense(C.c_int64(n), dp(a), C.c_int64(n), dp(b), C.c_int64(n), C.c_int(lowest),
                                              method.encode(), C.c_int(max_iterations), C.c_double(tolerance),
                                              C.c_int(max_dim_sub or 0), dp(ev), dp(vec), C.c_int64(n),
                                              C.byref(iters)))
    return ev, vec, iters.value


def eigensolver(matrix, lowest, method, max_iterations, tolerance, max_dim_sub=None):
    """`eigensolver` of README.md:20: the standard problem (no second_matrix)."""
    return generalized_eigensolver(matrix, lowest, method, max_iterations, tolerance, max_dim_sub)


def generalized_eigensolver_builtin(dim, op_matrix, op_second_matrix, lowest, method, max_iterations, tolerance,
                                    max_dim_sub=None):
    """Matrix-free solve with the built-in device generators (benchmark_free.f90 / test_utils.f90 operators)."""
    ev = np.zeros(lowest)
    vec = np.zeros((dim, lowest), order="F")
    iters = C.c_int(-1)
    check(lib().dav_generalized_eigensolver_free_builtin(C.c_int64(dim), C.c_int(op_matrix), C.c_int(op_second_matrix),
                                                         C.c_int(lowest), method.encode(), C.c_int(max_iterations),
                                                         C.c_double(tolerance), C.c_int(max_dim_sub or 0), dp(ev),
                                                         dp(vec), C.c_int64(dim), C.byref(iters)))
    return ev, vec, (iters.value if iters.value >= 0 else None)


def free_matmul(op, array):
    """free_matmul(fun, array) (davidson.f90:526-569) for a built-in generator `op`."""
    x = np.asfortranarray(array, dtype=np.float64)
    out = np.zeros_like(x, order="F")
    check(lib().dav_free_matmul(C.c_int(op), C.c_int64(x.shape[0]), C.c_int64(x.shape[1]), dp(x), dp(out)))
    return out


def compute_matrix_on_the_fly(op, i, dim):
    """compute_matrix_on_the_fly(i, dim) / compute_stx_on_the_fly (benchmark_free.f90:38-76): column i (1-based)."""
    out = np.zeros(dim)
    check(lib().dav_compute_on_the_fly(C.c_int(op), C.c_int64(i), C.c_int64(dim), dp(out)))
    return out


class DavidsonSolver:
    """Device-resident solver handle (dav_solver_t)."""

    def __init__(self, device=0, rank=0, world_size=1, nccl_id=None):
        self._h = C.c_void_p()
        self._cbs = []
        L = lib()
        if world_size > 1:
            buf = C.create_string_buffer(bytes(nccl_id), 128)
            check(L.dav_create_distributed(C.byref(self._h), C.c_int(device), C.c_int(rank), C.c_int(world_size), buf))
        else:
            check(L.dav_create(C.byref(self._h), C.c_int(device)))
        self.rank, self.world_size = rank, world_size
        self.n = 0

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        check(lib().dav_get_unique_id(buf))
        return buf.raw

    def close(self):
        self._release_pinned()
        if self._h:
            lib().dav_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rows(self):
        b, e = C.c_int64(), C.c_int64()
        check(lib().dav_partition_rows(C.c_int64(self.n), C.c_int(self.world_size), C.c_int(self.rank), C.byref(b),
                                       C.byref(e)))
        return b.value, e.value

    def generate_diagonal_dominant(self, which, n, sparsity, diag_val=None, seed=0):
        check(lib().dav_matrix_generate_diagonal_dominant(self._h, C.c_int(which), C.c_int64(n), C.c_double(sparsity),
                                                          C.c_int(0 if diag_val is None else 1),
                                                          C.c_double(diag_val or 0.0), C.c_uint64(seed)))
        self.n = n

    def upload(self, which, matrix):
        a = np.asfortranarray(matrix, dtype=np.float64)
        check(lib().dav_matrix_upload(self._h, C.c_int(which), C.c_int64(a.shape[0]), dp(a), C.c_int64(a.shape[0])))
        self.n = a.shape[0]

    def upload_ptr(self, which, n, ptr, ld):
        check(lib().dav_matrix_upload(self._h, C.c_int(which), C.c_int64(n), C.c_void_p(ptr), C.c_int64(ld)))
        self.n = n

    def upload_rows_ptr(self, which, n, ptr, ld):
        """This rank's row block only: (rows() x n) column-major host memory at `ptr`, leading dimension ld."""
        check(lib().dav_matrix_upload_rows(self._h, C.c_int(which), C.c_int64(n), C.c_void_p(ptr), C.c_int64(ld)))
        self.n = n

    def set_operator(self, which, n, op):
        check(lib().dav_matrix_set_operator(self._h, C.c_int(which), C.c_int64(n), C.c_int(op)))
        self.n = n

    def set_callback(self, which, n, fun, diag=None):
        cb = _wrap_callback(fun)
        self._cbs.append(cb)
        d = np.ascontiguousarray(diag, dtype=np.float64) if diag is not None else None
        check(lib().dav_matrix_set_callback(self._h, C.c_int(which), C.c_int64(n), cb, None, dp(d)))
        self.n = n

    def set_device_callback(self, which, n, fun, diag=None):
        """Matrix-free operator as a DEVICE functor: fun(d_x, ldx, d_y, ldy, n, b, row_begin, nrows, stream) gets raw
        device addresses (ints) and must enqueue on the CUDA stream `stream` the computation of rows
        [row_begin, row_begin + nrows) of Op * X (see dav_matrix_set_device_callback)."""

        def cb(xp, ldx, yp, ldy, n_, b, r0, nr, stream, _ctx):
            fun(int(xp or 0), int(ldx), int(yp or 0), int(ldy), int(n_), int(b), int(r0), int(nr), int(stream or 0))

        cfn = DEVICE_GEMV_FN(cb)
        self._cbs.append(cfn)
        d = np.ascontiguousarray(diag, dtype=np.float64) if diag is not None else None
        check(lib().dav_matrix_set_device_callback(self._h, C.c_int(which), C.c_int64(n), cfn, None, dp(d)))
        self.n = n

    def clear(self, which):
        check(lib().dav_matrix_clear(self._h, C.c_int(which)))

    def download(self, which):
        r0, r1 = self.rows()
        out = np.zeros((r1 - r0, self.n), order="F")
        check(lib().dav_matrix_download(self._h, C.c_int(which), dp(out), C.c_int64(max(r1 - r0, 1))))
        return out

    def set_matvec_impl(self, impl):
        check(lib().dav_set_matvec_impl(self._h, C.c_int(impl)))

    def _pinned_vectors(self, n, lowest):
        """(n x lowest) column-major result array in page-locked memory (reused between solves of one shape):
        the library then writes the eigenvectors by DMA instead of staging + host copy."""
        key = (n, lowest)
        if getattr(self, "_pin_key", None) != key:
            self._release_pinned()
            ptr = C.c_void_p()
            check(lib().dav_alloc_pinned(C.c_size_t(8 * n * lowest), C.byref(ptr)))
            self._pin_ptr, self._pin_key = ptr, key
            buf = (C.c_double * (n * lowest)).from_address(ptr.value)
            self._pin_arr = np.frombuffer(buf, dtype=np.float64).reshape((n, lowest), order="F")
        return self._pin_arr

    def _release_pinned(self):
        if getattr(self, "_pin_ptr", None):
            self._pin_arr = None
            lib().dav_free_pinned(self._pin_ptr)
            self._pin_ptr, self._pin_key = None, None

    def solve(self, lowest, method, max_iterations, tolerance, max_dim_sub=None, want_vectors=True, pinned=False,
              local=False):
        """pinned=True returns the eigenvectors in a page-locked array owned by this handle (valid until the next
        solve of another shape or close()).  local=True (dav_solve_local): every rank gets only its rows() of the
        Ritz vectors instead of the all-gathered block."""
        ev = np.zeros(lowest)
        rows = self.n
        if local:
            r0, r1 = self.rows()
            rows = max(r1 - r0, 1)
        if want_vectors:
            vec = self._pinned_vectors(rows, lowest) if pinned else np.zeros((rows, lowest), order="F")
        else:
            vec = None
        if local:
            iters = C.c_int(-1)
            check(lib().dav_solve_local(self._h, C.c_int(lowest), C.c_int(METHODS[method]), C.c_int(max_iterations),
                                        C.c_double(tolerance), C.c_int(max_dim_sub or 0), dp(ev), dp(vec),
                                        C.c_int64(rows), C.byref(iters)))
            return ev, vec, (iters.value if iters.value >= 0 else None)
        iters = C.c_int(-1)
        check(lib().dav_solve(self._h, C.c_int(lowest), C.c_int(METHODS[method]), C.c_int(max_iterations),
                              C.c_double(tolerance), C.c_int(max_dim_sub or 0), dp(ev), dp(vec), C.c_int64(self.n),
                              C.byref(iters)))
        return ev, vec, (iters.value if iters.value >= 0 else None)

    def set_profiling(self, per_phase_spans):
        """Per-phase event spans of the following solves (Stats.matvec_ms, rr_ms, ...); off by default."""
        check(lib().dav_set_profiling(self._h, C.c_int(1 if per_phase_spans else 0)))

    def stats(self):
        s = Stats()
        check(lib().dav_get_stats(self._h, C.byref(s)))
        return s

    def block_matvec(self, which, x):
        x = np.asfortranarray(x, dtype=np.float64)
        r0, r1 = self.rows()
        w = np.zeros((r1 - r0, x.shape[1]), order="F")
        check(lib().dav_block_matvec(self._h, C.c_int(which), C.c_int64(x.shape[1]), dp(x), C.c_int64(x.shape[0]),
                                     dp(w), C.c_int64(max(r1 - r0, 1))))
        return w

    def bench_fp64_pipe(self, reps=3):
        """Measured DMMA (FP64 tensor pipe) peak of this device in TFLOP/s (register-only kernel)."""
        out = C.c_double(0.0)
        check(lib().dav_bench_fp64_pipe(self._h, C.c_int(reps), C.byref(out)))
        return out.value

    def debug_collective(self, kind, count, reps=20):
        """(microseconds per call, max abs error) of one inter-GPU exchange; see dav_debug_collective."""
        out = (C.c_double * 2)()
        check(lib().dav_debug_collective(self._h, C.c_int(kind), C.c_int64(count), C.c_int(reps), out))
        return out[0], out[1]

    def comm_info(self):
        """{'peer': bool, 'peer_calls': int, 'nccl_calls': int} of a distributed handle."""
        p, a, b = C.c_int(0), C.c_longlong(0), C.c_longlong(0)
        check(lib().dav_comm_info(self._h, C.byref(p), C.byref(a), C.byref(b)))
        return {"peer": bool(p.value), "peer_calls": a.value, "nccl_calls": b.value}

    def bench_block_matvec(self, which, b, reps):
        ms = (C.c_float * reps)()
        check(lib().dav_bench_block_matvec(self._h, C.c_int(which), C.c_int64(b), C.c_int(reps), ms))
        return list(ms)
