"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the committed golden
fixtures on the same seeded inputs.  Tolerances are BASELINE.json's: eigenvalues 1e-10 relative,
eigenvectors equal up to sign (1e-8), residual <= tolerance, iteration count +-1."""
import numpy as np
import pytest
import scipy.linalg as sl

import fortran_davidson_b200 as fd
from conftest import case_inputs
from fortran_davidson_b200 import davidson as dv
from fortran_davidson_b200 import lapack_wrapper as lw
from fortran_davidson_b200._lib import DavidsonError
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

EV_RTOL = 1e-10
VEC_ATOL = 1e-8


def _check_pairs(A, B, ev, vec, ref_ev, ref_vec, tol):
    assert np.abs(ev - ref_ev).max() / np.abs(ref_ev).max() < EV_RTOL
    Bm = B if B is not None else np.eye(A.shape[0])
    for j in range(len(ev)):
        s = np.sign(vec[:, j] @ (Bm @ ref_vec[:, j]))
        assert np.abs(s * vec[:, j] - ref_vec[:, j]).max() < VEC_ATOL
        res = A @ vec[:, j] - ev[j] * (Bm @ vec[:, j])
        assert np.linalg.norm(res) < max(tol, 1e-8)


# ---------------------------------------------------------------- generators (bit exact)
@pytest.mark.parametrize("m,sparsity,diag,seed", [(1, 1e-4, None, 0), (50, 1e-4, None, 0), (257, 1e-3, 1.0, 7),
                                                  (1000, 5e-2, None, 123456789)])
def test_generate_diagonal_dominant_bit_exact(m, sparsity, diag, seed):
    a = fd.generate_diagonal_dominant(m, sparsity, diag, seed)
    b = orc.generate_diagonal_dominant(m, sparsity, diag, seed)
    assert np.array_equal(a, b)


def test_on_the_fly_columns_match_oracle():
    for op in (dv.OP_BENCHMARK_MTX, dv.OP_IDENTITY, dv.OP_TEST_MTX, dv.OP_TEST_STX):
        for (i, dim) in ((1, 50), (17, 50), (50, 50), (999, 1000)):
            a = dv.compute_matrix_on_the_fly(op, i, dim)
            b = orc.compute_on_the_fly(op, i, dim)
            assert np.allclose(a, b, rtol=1e-14, atol=1e-19), (op, i, dim)


@pytest.mark.parametrize("op", [dv.OP_BENCHMARK_MTX, dv.OP_IDENTITY, dv.OP_TEST_STX])
@pytest.mark.parametrize("n,b", [(50, 3), (333, 20), (1000, 70)])
def test_free_matmul_matches_oracle(op, n, b):
    x = np.random.default_rng(n + b).standard_normal((n, b))
    w = fd.free_matmul(op, x)
    ref = orc.free_matmul(op, x)
    assert np.abs(w - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("op", [dv.OP_BENCHMARK_MTX, dv.OP_TEST_MTX, dv.OP_TEST_STX])
@pytest.mark.parametrize("n,b", [(257, 8), (3000, 16), (5000, 40), (4097, 64), (3001, 100), (2500, 128), (2000, 200)])
def test_free_matmul_tensor_pipe_matches_libm_kernel(op, n, b):
    """The tensor-pipe generator (piecewise polynomial entries + DMMA, csrc/freeops_dmma.cu) against the SIMT
    kernel that evaluates atan2/log/sqrt/cos with libm, on the same operator (benchmark_free.f90:38-76)."""
    x = np.random.default_rng(n * 7 + b).standard_normal((n, b))
    s = fd.DavidsonSolver()
    s.set_operator(0, n, op)
    s.set_matvec_impl(dv.MATVEC_SIMT)
    ref = s.block_matvec(0, x)
    s.set_matvec_impl(dv.MATVEC_TMA_DMMA)
    w = s.block_matvec(0, x)
    s.close()
    assert np.abs(w - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
    if n <= 3000:
        assert np.abs(w - orc.free_matmul(op, x)).max() <= 1e-13 * max(1.0, np.abs(ref).max())


# ---------------------------------------------------------------- block matvec
@pytest.mark.parametrize("impl", [dv.MATVEC_SIMT, dv.MATVEC_TMA_DMMA])
@pytest.mark.parametrize("n,b", [(50, 6), (300, 1), (1000, 5), (1024, 8), (2000, 20), (4097, 33), (3000, 64),
                                 (2500, 80), (2048, 128), (1500, 130)])
def test_block_matvec_parity(impl, n, b):
    rng = np.random.default_rng(n * 1000 + b)
    A = orc.generate_diagonal_dominant(n, 1e-2, seed=n)
    A = A + 0.01 * rng.standard_normal((n, n))  # deliberately NOT symmetric: the kernel computes A*X, not A^T*X
    X = rng.standard_normal((n, b))
    s = fd.DavidsonSolver()
    s.upload(0, A)
    s.set_matvec_impl(impl)
    W = s.block_matvec(0, X)
    ref = A @ X
    assert np.abs(W - ref).max() <= 1e-12 * np.abs(ref).max()
    W2 = s.block_matvec(0, X)
    assert np.array_equal(W, W2)  # bit reproducible (no atomics)
    s.close()


@pytest.mark.parametrize("n,b,uploaded", [(1000, 5, True), (4097, 33, True), (3000, 64, True), (9700, 128, True),
                                          (9473, 100, True), (19100, 64, False), (19000, 40, False),
                                          (38100, 16, False), (37888, 32, False), (76000, 8, False)])
def test_block_matvec_schedules(n, b, uploaded, monkeypatch):
    """The work schedules of the block matvec -- full waves + stream-K remainder (DAV_MATVEC_SCHEDULE=1), full waves +
    aligned split-K remainder (=2) -- and both pipeline-stage depths (DAV_MATVEC_BK=16/32 columns of A per stage)
    against pure stream-K (=0) and the SIMT kernel, at sizes with and without a full wave of row tiles on 148 SMs
    (tile rows: 64 for b > 64, 128 for b > 32, else 256)."""
    rng = np.random.default_rng(n + b)
    X = rng.standard_normal((n, b))
    s = fd.DavidsonSolver()
    if uploaded:
        A = orc.generate_diagonal_dominant(n, 1e-3, seed=n) + 0.01 * rng.standard_normal((n, n))
        s.upload(0, A)
    else:
        s.generate_diagonal_dominant(0, n, 1e-4, None, 3)
    s.set_matvec_impl(dv.MATVEC_TMA_DMMA)
    W = {}
    variants = [(sched, bk) for bk in (16, 32) for sched in (0, 1, 2)]
    for sched, bk in variants:
        monkeypatch.setenv("DAV_MATVEC_SCHEDULE", str(sched))
        monkeypatch.setenv("DAV_MATVEC_BK", str(bk))
        W[sched, bk] = s.block_matvec(0, X)
        assert np.array_equal(W[sched, bk], s.block_matvec(0, X))  # bit reproducible
    s.set_matvec_impl(dv.MATVEC_SIMT)
    Ws = s.block_matvec(0, X)
    s.close()
    scale = np.abs(Ws).max()
    for v in variants:
        assert np.abs(W[v] - W[0, 16]).max() <= 1e-13 * scale, v
        assert np.abs(W[v] - Ws).max() <= 1e-12 * scale, v
    if uploaded:
        ref = A @ X
        for v in variants:
            assert np.abs(W[v] - ref).max() <= 1e-12 * np.abs(ref).max(), v


@pytest.mark.parametrize("m,k,b", [(50000, 100000, 64), (25000, 100000, 128), (12544, 100000, 64), (12192, 100000, 16),
                                   (19000, 30000, 100)])
def test_block_matvec_rectangular_row_blocks(m, k, b, monkeypatch):
    """The row blocks of the sharded solve are rectangular (m = rows of a rank, k = n): TMA/DMMA matvec under every
    schedule against the library's tall-skinny GEMM on device-generated data (one GPU validates the multi-GPU tiles)."""
    import ctypes as C
    L = fd.lib()
    L.dav_debug_matvec_rect.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_double),
                                        C.POINTER(C.c_double)]
    for sched in (None, "0", "1", "2"):
        if sched is None:
            monkeypatch.delenv("DAV_MATVEC_SCHEDULE", raising=False)
        else:
            monkeypatch.setenv("DAV_MATVEC_SCHEDULE", sched)
        d, s = C.c_double(), C.c_double()
        assert L.dav_debug_matvec_rect(0, m, k, b, C.byref(d), C.byref(s)) == 0, L.dav_last_error()
        assert s.value > 1.0 and d.value <= 1e-11 * s.value, (sched, d.value, s.value)


def test_block_matvec_linearity_large():
    """Size-independent property at a size the oracle would not finish quickly: A(x+2y) == Ax + 2Ay and
    device-generated A equals the oracle's stream on a sampled sub-block."""
    n = 20000
    s = fd.DavidsonSolver()
    s.generate_diagonal_dominant(0, n, 1e-4, None, 0)
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal((n, 20)), rng.standard_normal((n, 20))
    wx, wy, wxy = s.block_matvec(0, x), s.block_matvec(0, y), s.block_matvec(0, x + 2 * y)
    assert np.abs(wxy - (wx + 2 * wy)).max() < 1e-9 * np.abs(wxy).max()
    # rows 0..3 of A*x against entries regenerated on the host
    for i in (0, 1, 7777, n - 1):
        row = np.array([orc.uniform01(0, min(i, j), max(i, j)) * 1e-4 if i != j else i + 1.0 for j in range(n)])
        assert np.abs(row @ x - wx[i]).max() < 1e-10 * max(1.0, np.abs(wx[i]).max())
    s.close()


@pytest.mark.parametrize("ta,m,k,n", [("N", 1000, 37, 5), ("N", 4099, 128, 64), ("N", 100, 100, 100), ("N", 70001, 64, 33),
                                      ("N", 20000, 256, 128), ("N", 1, 1, 1), ("N", 65, 3, 129),
                                      ("T", 70, 5000, 33), ("T", 128, 100000, 64), ("T", 1, 3, 1), ("T", 256, 20011, 128),
                                      ("T", 64, 12500, 32), ("T", 33, 1000, 130)])
def test_gemm_tensor_pipe_variant(ta, m, k, n, monkeypatch):
    """The tall-skinny products around the matvec (lapack_matmul call sites of davidson.f90:131,159,218,223) on the
    SIMT kernel (DAV_GEMM_IMPL=0) and on the DMMA kernel (=1) against numpy; m x n = op(A) (m x k) * B (k x n)."""
    from fortran_davidson_b200 import lapack_wrapper as lw
    rng = np.random.default_rng(m * 7 + k * 3 + n)
    A = rng.standard_normal((k, m) if ta == "T" else (m, k))
    B = rng.standard_normal((k, n))
    ref = (A.T if ta == "T" else A) @ B
    tol = 1e-13 * max(1.0, np.abs(ref).max()) * max(1.0, np.sqrt(k) / 4)
    out = {}
    for impl in ("0", "1"):
        monkeypatch.setenv("DAV_GEMM_IMPL", impl)
        out[impl] = lw.lapack_matmul(ta, "N", A, B, 0.5)
        assert np.abs(out[impl] - 0.5 * ref).max() <= tol, impl
        assert np.array_equal(out[impl], lw.lapack_matmul(ta, "N", A, B, 0.5))  # deterministic split-K reduction


# ---------------------------------------------------------------- lapack_wrapper / array_utils mirrors
def test_lapack_wrapper_mirrors():
    lw, au = fd.lapack_wrapper, fd.array_utils
    mtx = orc.generate_diagonal_dominant(50, 1e-3, seed=11)
    stx = orc.generate_diagonal_dominant(50, 1e-3, seed=12)
    # test_call_lapack.f90:22-30 + test_lapack.py:47-51
    w, v = lw.lapack_generalized_eigensolver(mtx)
    es, vs = sl.eigh(mtx)
    assert np.allclose(w, es, rtol=1e-12) and np.allclose(np.abs(v), np.abs(vs), atol=1e-9)
    w, v = lw.lapack_generalized_eigensolver(mtx, stx)
    es, vs = sl.eigh(mtx, b=stx)
    assert np.allclose(w, es, rtol=1e-11) and np.allclose(np.abs(v), np.abs(vs), atol=1e-8)
    wo, vo = orc.lapack_generalized_eigensolver(mtx, stx)
    assert np.allclose(w, wo, rtol=1e-11)
    wl, vl = lw.lapack_generalized_eigensolver_lowest(mtx, stx, 4)
    assert np.allclose(wl, es[:4], rtol=1e-11) and np.allclose(np.abs(vl), np.abs(vs[:, :4]), atol=1e-8)
    # random dense symmetric (not diagonally dominant), odd size, k > smem-resident threshold
    rng = np.random.default_rng(5)
    for k in (1, 2, 7, 33, 130, 200):
        S = rng.standard_normal((k, k)); S = S + S.T
        w, v = lw.lapack_generalized_eigensolver(S)
        assert np.allclose(w, np.linalg.eigvalsh(S), rtol=0, atol=1e-12 * k * max(1.0, np.abs(S).max()))
        assert np.abs(v.T @ v - np.eye(k)).max() < 1e-12 * k
        assert np.abs(S @ v - v * w).max() < 1e-11 * k
    # only the upper triangle is read (DSYEV 'U')
    junk = mtx.copy(); junk[np.tril_indices(50, -1)] = 99.0
    assert np.allclose(lw.lapack_generalized_eigensolver(junk)[0], sl.eigh(mtx)[0], rtol=1e-12)
    # QR (test_call_lapack.f90:33): equal to LAPACK's Q up to column signs
    q = lw.lapack_qr(mtx)
    qo = orc.lapack_qr(mtx)
    assert np.allclose(np.abs(q), np.abs(qo), atol=1e-10) and np.abs(q.T @ q - np.eye(50)).max() < 1e-13
    tall = rng.standard_normal((3001, 17))
    qt = lw.lapack_qr(tall)
    assert np.abs(qt.T @ qt - np.eye(17)).max() < 1e-13 and np.allclose(np.abs(qt), np.abs(orc.lapack_qr(tall)), atol=1e-10)
    # solver / matmul / matrix_vector / sort / norm / preconditioner
    a = rng.standard_normal((30, 30)); a = a + a.T
    b = rng.standard_normal(30)
    assert np.allclose(lw.lapack_solver(a, b), orc.lapack_solver(a, b), rtol=1e-9, atol=1e-11)
    p, qq = rng.standard_normal((700, 5)), rng.standard_normal((700, 4))
    assert np.allclose(lw.lapack_matmul("T", "N", p, qq), orc.lapack_matmul("T", "N", p, qq), rtol=1e-13, atol=1e-13)
    q3 = rng.standard_normal((3, 5))
    assert np.allclose(lw.lapack_matmul("N", "T", p, q3), p @ q3.T, rtol=1e-13, atol=1e-13)
    assert np.allclose(lw.lapack_matmul("T", "T", p, rng.standard_normal((6, 700))).shape, (5, 6))
    assert np.allclose(lw.lapack_matmul("N", "N", p.T, qq, 2.0), 2.0 * p.T @ qq, rtol=1e-13, atol=1e-12)
    assert np.allclose(lw.lapack_matrix_vector("N", a, b), a @ b, rtol=1e-13, atol=1e-13)
    assert np.allclose(lw.lapack_matrix_vector("T", p, qq[:, 0]), p.T @ qq[:, 0], rtol=1e-13, atol=1e-12)
    v = np.array([3.0, 1.0, 2.0, 0.5, 2.0])
    sv_, keys = lw.lapack_sort("I", v)
    assert list(sv_) == [0.5, 1.0, 2.0, 2.0, 3.0] and list(keys) == [5, 2, 3, 1, 4]
    sd, kd = lw.lapack_sort("D", v)
    assert list(sd) == [3.0, 2.0, 2.0, 1.0, 0.5] and list(kd) == [1, 4, 2, 5, 3]
    x = rng.standard_normal(100001)
    assert abs(au.norm(x) - np.linalg.norm(x)) < 1e-12 * np.linalg.norm(x)
    d = rng.standard_normal(5000)
    assert np.array_equal(au.generate_preconditioner(d, 12), orc.generate_preconditioner(d, 12))
    pre = au.generate_preconditioner(np.array([2.0, 1.0, 1.0, 3.0]), 3)  # ties: stable by index
    assert pre[1, 0] == 1.0 and pre[2, 1] == 1.0 and pre[0, 2] == 1.0 and pre.sum() == 3.0


# ---------------------------------------------------------------- dense solver vs oracle / golden
DENSE_DPR = ["matrix_txt_DPR", "readme_std_DPR", "readme_gev_DPR", "test_dense_numpy_std_DPR",
             "test_dense_numpy_gen_DPR", "main_f90_DPR", "collapse_n1000_DPR", "collapse_n1000_gev_DPR",
             "collapse_n2000_DPR"]


@pytest.mark.parametrize("name", DENSE_DPR)
def test_dense_dropin_parity(name, golden_cases):
    g = golden_cases[name]
    A, B = case_inputs(name)
    ev, vec, iters = fd.generalized_eigensolver(A, g["lowest"], g["method"], g["max_iterations"], g["tolerance"],
                                                g["max_dim_sub"], B)
    assert abs(iters - g["iters"]) <= 1
    r = orc.generalized_eigensolver(A, g["lowest"], g["method"], g["max_iterations"], g["tolerance"],
                                    g["max_dim_sub"], B)
    _check_pairs(A, B, ev, vec, r.eigenvalues, r.eigenvectors, g["tolerance"])
    assert np.allclose(ev, g["eigh"])  # the reference's own acceptance check (test_davidson.py:39-40)


DENSE_GJD = ["matrix_txt_GJD", "readme_std_GJD", "readme_gev_GJD", "test_dense_numpy_std_GJD",
             "test_dense_numpy_gen_GJD", "main_f90_GJD"]


@pytest.mark.parametrize("name", DENSE_GJD)
def test_dense_gjd_parity(name, golden_cases):
    """GJD on device (block MINRES on the projected correction equation) against the reference's O(n^3) DSYSV solve."""
    g = golden_cases[name]
    A, B = case_inputs(name)
    ev, vec, iters = fd.generalized_eigensolver(A, g["lowest"], "GJD", g["max_iterations"], g["tolerance"],
                                                g["max_dim_sub"], B)
    assert abs(iters - g["iters"]) <= 1
    r = orc.generalized_eigensolver(A, g["lowest"], "GJD", g["max_iterations"], g["tolerance"], g["max_dim_sub"], B)
    _check_pairs(A, B, ev, vec, r.eigenvalues, r.eigenvectors, g["tolerance"])
    assert np.allclose(ev, g["eigh"])
    # test_dense_properties.f90:25-26: both methods give the same eigenvalues
    ev2, _, _ = fd.generalized_eigensolver(A, g["lowest"], "DPR", g["max_iterations"], g["tolerance"], g["max_dim_sub"], B)
    assert np.linalg.norm(ev - ev2) < 1e-8


def test_dense_gjd_larger():
    for (n, sp, L, md, tol) in [(600, 1e-2, 3, 10, 1e-10), (1000, 5e-2, 4, 40, 1e-8)]:
        A = orc.generate_diagonal_dominant(n, sp, seed=0)
        B = orc.generate_diagonal_dominant(n, sp, 1.0, seed=1)
        for Bm in (None, B):
            r = orc.generalized_eigensolver(A, L, "GJD", 100, tol, md, Bm)
            s = fd.DavidsonSolver()
            s.upload(0, A)
            if Bm is not None:
                s.upload(1, Bm)
            ev, vec, iters = s.solve(L, "GJD", 100, tol, md)
            st = s.stats()
            assert abs(iters - r.iters) <= 1, (n, Bm is not None, iters, r.iters)
            assert np.abs(ev - r.eigenvalues).max() / np.abs(ev).max() < EV_RTOL
            assert 0 < st.gjd_inner_iterations <= 40 * iters
            s.close()


@pytest.mark.parametrize("impl", [dv.MATVEC_SIMT, dv.MATVEC_TMA_DMMA])
@pytest.mark.parametrize("name", ["readme_gev_DPR", "collapse_n2000_DPR", "collapse_n1000_gev_DPR"])
def test_dense_handle_trace(name, impl, golden_cases):
    """Same solve through the device-resident handle: basis schedule and residual trace match the oracle."""
    g = golden_cases[name]
    A, B = case_inputs(name)
    s = fd.DavidsonSolver()
    s.set_matvec_impl(impl)
    s.upload(0, A)
    if B is not None:
        s.upload(1, B)
    ev, vec, iters = s.solve(g["lowest"], "DPR", g["max_iterations"], g["tolerance"], g["max_dim_sub"])
    st = s.stats()
    assert iters == g["iters"]
    assert list(st.trace_k[:st.trace_len]) == g["trace_k"]
    te = np.array(st.trace_err[:st.trace_len])
    ge = np.array(g["trace_err"])
    big = ge > 1e-6  # residuals well above round-off agree to several digits
    assert np.allclose(te[big], ge[big], rtol=1e-3)
    assert np.abs(ev - np.array(g["eigenvalues"])).max() / np.abs(ev).max() < EV_RTOL
    assert st.kernel_launches > 0 and st.matvec_launches > 0
    s.close()


def test_dense_device_generated_matches_uploaded():
    n = 1500
    s = fd.DavidsonSolver()
    s.generate_diagonal_dominant(0, n, 1e-3, None, 3)
    s.generate_diagonal_dominant(1, n, 1e-3, 1.0, 4)
    assert np.array_equal(s.download(0), orc.generate_diagonal_dominant(n, 1e-3, None, 3))
    ev, vec, iters = s.solve(4, "DPR", 100, 1e-9)
    A, B = orc.generate_diagonal_dominant(n, 1e-3, None, 3), orc.generate_diagonal_dominant(n, 1e-3, 1.0, 4)
    r = orc.generalized_eigensolver(A, 4, "DPR", 100, 1e-9, None, B)
    assert abs(iters - r.iters) <= 1
    _check_pairs(A, B, ev, vec, r.eigenvalues, r.eigenvectors, 1e-9)
    s.close()


def _sym_cases():
    rng = np.random.default_rng(11)
    out = []
    for k in (48, 64, 97, 128, 160, 200, 256, 320):
        a = rng.standard_normal((k, k))
        out.append(("random%d" % k, (a + a.T) / 2, True))
    k = 128
    out.append(("diagdom", np.diag(np.arange(1.0, k + 1)) + 1e-4 * (lambda r: (r + r.T) / 2)(rng.random((k, k))), True))
    g = np.diag(np.concatenate([np.arange(1.0, 65), np.linspace(50, 1e5, 64)])) + 1e-2 * rng.standard_normal((k, k))
    out.append(("graded", (g + g.T) / 2, True))
    out.append(("near_identity", np.eye(k) + 1e-4 * (lambda r: (r + r.T) / 2)(rng.random((k, k))), True))
    out.append(("identity", np.eye(k), False))                         # exactly degenerate -> Jacobi fallback
    q, _ = np.linalg.qr(rng.standard_normal((k, k)))
    out.append(("degenerate_pairs", (q * np.repeat(np.arange(1.0, 65), 2)) @ q.T, None))  # either path, must be right
    out.append(("zero", np.zeros((64, 64)), None))
    blk = np.zeros((96, 96)); blk[:48, :48] = out[0][1]; blk[48:, 48:] = out[0][1] + 3.0 * np.eye(48)
    out.append(("block_diagonal", blk, None))
    return out


@pytest.mark.parametrize("case", _sym_cases(), ids=lambda c: c[0])
def test_sym_eigh_tridiagonal_path(case):
    """The Rayleigh-Ritz eigensolver (tridiagonalisation + multisection + twisted factorisation, guarded, Jacobi
    fallback) against numpy's LAPACK: eigenvalues, orthonormality and residual at round-off level whichever
    path produced them (lapack_wrapper.f90:14-91 contract: ascending eigenvalues, orthonormal vectors)."""
    name, S, expect_fast = case
    k = S.shape[0]
    w, v, info, _ = lw.sym_eigh_info(np.triu(S))  # only the upper triangle may be read
    ref = np.linalg.eigvalsh(S)
    scale = max(np.abs(S).max(), 1e-300)
    assert np.abs(w - ref).max() <= 1e-12 * k * scale, (name, info)
    assert np.abs(v.T @ v - np.eye(k)).max() < 1e-12, (name, info)
    assert np.abs(S @ v - v * w).max() <= 1e-12 * k * scale, (name, info)
    if expect_fast is True:
        assert info["accepted"] == 1, (name, info)
    if expect_fast is False:
        assert info["accepted"] == 0, (name, info)


def test_upload_rows_matches_upload():
    """dav_matrix_upload_rows (a rank's own row block, its own leading dimension) == dav_matrix_upload."""
    n = 700
    A = orc.generate_diagonal_dominant(n, 1e-2, None, 5)
    s = fd.DavidsonSolver()
    r0, r1 = 0, n
    blk = np.zeros((r1 - r0 + 5, n), order="F")  # ld = rows + 5
    blk[:r1 - r0] = A[r0:r1]
    s.upload_rows_ptr(0, n, blk.ctypes.data, blk.shape[0])
    assert np.array_equal(s.download(0), A)
    ev, vec, iters = s.solve(3, "DPR", 200, 1e-9)
    s.upload(0, A)
    ev2, vec2, iters2 = s.solve(3, "DPR", 200, 1e-9)
    assert iters == iters2 and np.array_equal(ev, ev2) and np.array_equal(vec, vec2)
    s.close()


def test_pinned_result_array_matches_pageable():
    """Eigenvectors written by DMA into a page-locked array (dav_alloc_pinned) == staged copy into a pageable one."""
    s = fd.DavidsonSolver()
    s.generate_diagonal_dominant(0, 1200, 1e-2, None, 2)
    ev1, vec1, it1 = s.solve(4, "DPR", 200, 1e-9)
    ev2, vec2, it2 = s.solve(4, "DPR", 200, 1e-9, pinned=True)
    assert it1 == it2 and np.array_equal(ev1, ev2) and np.array_equal(vec1, np.array(vec2))
    s.close()


def test_not_converged_semantics(golden_cases):
    g = golden_cases["notconverged_DPR"]
    A, _ = case_inputs("notconverged_DPR")
    ev, vec, iters = fd.generalized_eigensolver(A, g["lowest"], "DPR", g["max_iterations"], g["tolerance"])
    assert iters == g["max_iterations"] + 1  # davidson.f90:232-235
    assert np.allclose(ev, g["eigenvalues"], rtol=1e-9)  # Ritz values of the last Rayleigh-Ritz step (:186)


def test_error_behaviour():
    A = orc.generate_diagonal_dominant(20, 1e-3)
    with pytest.raises(DavidsonError) as ei:
        fd.generalized_eigensolver(A, 2, "XYZ", 10, 1e-8)  # unknown method is rejected
    assert ei.value.code == 1
    with pytest.raises(DavidsonError):
        fd.generalized_eigensolver(A, 11, "DPR", 10, 1e-8)  # 2*lowest > n
    with pytest.raises(DavidsonError) as ei:  # basis would outgrow the matrix (DORGQR failure in the reference)
        fd.generalized_eigensolver(A, 4, "DPR", 50, 1e-30, 16)
    assert ei.value.code == 5
    Bbad = -np.eye(20)
    with pytest.raises(DavidsonError) as ei:  # DSYGV info > n in the reference
        fd.generalized_eigensolver(A, 2, "DPR", 10, 1e-8, None, Bbad)
    assert ei.value.code == 3


def test_unsorted_diagonal_initial_basis():
    """generate_preconditioner on a matrix whose diagonal is not ascending (what test_reorder.f90 exercises)."""
    n = 400
    rng = np.random.default_rng(9)
    perm = rng.permutation(n)
    A0 = orc.generate_diagonal_dominant(n, 1e-3, seed=21)
    A = np.asfortranarray(A0[np.ix_(perm, perm)])
    ev, vec, iters = fd.generalized_eigensolver(A, 6, "DPR", 50, 1e-8, 18)
    r = orc.generalized_eigensolver(A, 6, "DPR", 50, 1e-8, 18)
    assert abs(iters - r.iters) <= 1
    _check_pairs(A, None, ev, vec, r.eigenvalues, r.eigenvectors, 1e-8)


# ---------------------------------------------------------------- matrix free
@pytest.mark.parametrize("name", ["free_test_50", "free_benchmark_300_L8", "free_benchmark_1000"])
def test_free_builtin_parity(name, golden_cases):
    g = golden_cases[name]
    ev, vec, iters = dv.generalized_eigensolver_builtin(g["dim"], g["op_a"], g["op_b"], g["lowest"], "DPR",
                                                        g["max_iterations"], g["tolerance"], g["max_dim_sub"])
    assert iters is not None and abs(iters - g["iters"]) <= 1
    assert np.abs(ev - np.array(g["eigenvalues"])).max() / np.abs(ev).max() < EV_RTOL
    assert np.allclose(ev, g["eigh"])  # test_davidson.py:69
    Ma, Mb = orc.operator_matrix(g["op_a"], g["dim"]), orc.operator_matrix(g["op_b"], g["dim"])
    for j in range(g["lowest"]):  # test_free_properties.f90:30-34 / benchmark_free.f90:104-108
        assert np.linalg.norm(Ma @ vec[:, j] - ev[j] * (Mb @ vec[:, j])) < 1e-8


def test_free_callback_parity(golden_cases):
    """The procedure-argument form (davidson.f90:277-312) with host callbacks."""
    g = golden_cases["free_test_50"]
    Ma, Mb = orc.operator_matrix(g["op_a"], g["dim"]), orc.operator_matrix(g["op_b"], g["dim"])
    calls = {"a": 0, "b": 0}

    def fa(x):
        calls["a"] += 1
        return Ma @ x

    def fb(x):
        calls["b"] += 1
        return Mb @ x

    ev, vec, iters = fd.generalized_eigensolver(fa, g["lowest"], "GJD", g["max_iterations"], g["tolerance"],
                                                g["max_dim_sub"], fun_second_matrix_gemv=fb, dim=g["dim"])
    assert iters is not None and abs(iters - g["iters"]) <= 1  # `method` is ignored: always DPR (:428)
    assert np.abs(ev - np.array(g["eigenvalues"])).max() / np.abs(ev).max() < EV_RTOL
    assert calls["a"] > 0 and calls["b"] > 0


def test_free_not_converged_leaves_iters_unassigned():
    ev, vec, iters = dv.generalized_eigensolver_builtin(200, dv.OP_BENCHMARK_MTX, dv.OP_IDENTITY, 3, "DPR", 1, 1e-14, 20)
    assert iters is None  # davidson.f90:417


# ---------------------------------------------------------------- BASELINE config 2 at full size
def test_config2_full_size_properties():
    """n = 20,000, lowest = 10, DPR, max_dim_sub = 100 (BASELINE.json configs[1]): iteration schedule, residuals
    through an independent block matvec, and eigenvalues against the oracle run on the same generated matrix."""
    n, L = 20000, 10
    s = fd.DavidsonSolver()
    s.generate_diagonal_dominant(0, n, 1e-4, None, 0)
    ev, vec, iters = s.solve(L, "DPR", 1000, 1e-8, 100)
    st = s.stats()
    assert iters == 3 and list(st.trace_k[:st.trace_len]) == [20, 40, 80]
    s.set_matvec_impl(dv.MATVEC_SIMT)
    av = s.block_matvec(0, vec)
    res = np.linalg.norm(av - vec * ev, axis=0)
    assert res.max() < 1e-8
    assert np.abs(vec.T @ vec - np.eye(L)).max() < 1e-10
    A = orc.generate_diagonal_dominant(n, 1e-4, None, 0)
    r = orc.generalized_eigensolver(A, L, "DPR", 1000, 1e-8, 100)
    assert r.iters == iters
    assert np.abs(ev - r.eigenvalues).max() / np.abs(ev).max() < EV_RTOL
    for j in range(L):
        sg = np.sign(vec[:, j] @ r.eigenvectors[:, j])
        assert np.abs(sg * vec[:, j] - r.eigenvectors[:, j]).max() < VEC_ATOL
    s.close()
