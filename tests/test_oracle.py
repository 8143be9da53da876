"""CPU tests that pin the oracle (oracle/davidson_oracle.cpp) the way the reference pins itself:
eigenvalues == scipy.linalg.eigh on the same matrix (src/tests/test_davidson.py:36-40,67-69),
lapack wrapper eigenpairs == eigh (src/tests/test_lapack.py:47-51), residual norms below 1e-8
(src/tests/test_dense_properties.f90:31-39), plus the committed golden fixtures."""
import numpy as np
import pytest
import scipy.linalg as sl

from conftest import case_inputs
from oracle import oracle as orc

DENSE_CASES = ["matrix_txt_DPR", "matrix_txt_GJD", "readme_std_DPR", "readme_std_GJD", "readme_gev_DPR",
               "readme_gev_GJD", "test_dense_numpy_std_DPR", "test_dense_numpy_std_GJD", "test_dense_numpy_gen_DPR",
               "test_dense_numpy_gen_GJD", "main_f90_DPR", "main_f90_GJD", "collapse_n1000_DPR",
               "collapse_n1000_gev_DPR", "collapse_n2000_DPR"]


def test_matrix_txt_known_eigenvalues(matrix_100):
    # lowest eigenvalues of the reference's only deterministic fixture
    es = sl.eigh(matrix_100)[0][:3]
    assert np.allclose(es, [0.99998105, 2.00001545, 2.99997773], atol=1e-8)
    r = orc.generalized_eigensolver(matrix_100, 3, "DPR", 1000, 1e-8)
    assert np.allclose(r.eigenvalues, es, rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", DENSE_CASES)
def test_dense_matches_eigh_and_golden(name, golden_cases):
    g = golden_cases[name]
    A, B = case_inputs(name)
    r = orc.generalized_eigensolver(A, g["lowest"], g["method"], g["max_iterations"], g["tolerance"],
                                    g["max_dim_sub"], B)
    es = sl.eigh(A, B)[0][:g["lowest"]]
    # the reference's own acceptance test (np.allclose defaults)
    assert np.allclose(es, r.eigenvalues)
    # and much tighter: the oracle is converged to the tolerance
    assert np.abs(es - r.eigenvalues).max() < 1e-9
    assert r.iters == g["iters"]
    assert list(r.trace_k) == g["trace_k"]
    assert np.allclose(r.eigenvalues, g["eigenvalues"], rtol=1e-12, atol=0)
    # residual check of test_dense_properties.f90:31-39 / main.f90:64-72
    Bm = B if B is not None else np.eye(A.shape[0])
    for j in range(g["lowest"]):
        res = A @ r.eigenvectors[:, j] - r.eigenvalues[j] * (Bm @ r.eigenvectors[:, j])
        assert np.linalg.norm(res) < max(g["tolerance"], 1e-8)


def test_dpr_and_gjd_agree(golden_cases):
    # test_dense_properties.f90:25-26: ||lambda_GJD - lambda_DPR|| < 1e-8
    a = np.array(golden_cases["test_dense_numpy_std_DPR"]["eigenvalues"])
    b = np.array(golden_cases["test_dense_numpy_std_GJD"]["eigenvalues"])
    assert np.linalg.norm(a - b) < 1e-8


def test_not_converged_sets_iters(golden_cases):
    g = golden_cases["notconverged_DPR"]
    A, _ = case_inputs("notconverged_DPR")
    r = orc.generalized_eigensolver(A, g["lowest"], "DPR", g["max_iterations"], g["tolerance"])
    assert r.iters == g["max_iterations"] + 1  # davidson.f90:232-235


@pytest.mark.parametrize("name", ["free_test_50", "free_benchmark_300_L8"])
def test_free_matches_eigh(name, golden_cases):
    g = golden_cases[name]
    r = orc.generalized_eigensolver_free(g["dim"], g["op_a"], g["op_b"], g["lowest"], "DPR", g["max_iterations"],
                                         g["tolerance"], g["max_dim_sub"])
    Ma, Mb = orc.operator_matrix(g["op_a"], g["dim"]), orc.operator_matrix(g["op_b"], g["dim"])
    assert np.abs(Ma - Ma.T).max() == 0.0  # both branches of benchmark_free.f90:54-58 give a symmetric matrix
    es = sl.eigh(Ma, Mb)[0][:g["lowest"]]
    assert np.allclose(es, r.eigenvalues)  # test_davidson.py:69
    assert np.abs(es - r.eigenvalues).max() < 1e-9
    assert r.iters == g["iters"] and list(r.trace_k) == g["trace_k"]
    for j in range(g["lowest"]):  # test_free_properties.f90:30-34
        res = Ma @ r.eigenvectors[:, j] - r.eigenvalues[j] * (Mb @ r.eigenvectors[:, j])
        assert np.linalg.norm(res) < 1e-8


def test_lapack_wrappers_vs_scipy():
    # test_call_lapack.f90:22-33 + test_lapack.py:30-66
    mtx = orc.generate_diagonal_dominant(50, 1e-3, seed=11)
    stx = orc.generate_diagonal_dominant(50, 1e-3, seed=12)
    w, v = orc.lapack_generalized_eigensolver(mtx)
    es, vs = sl.eigh(mtx)
    assert np.allclose(w, es) and np.allclose(np.abs(v), np.abs(vs))
    w, v = orc.lapack_generalized_eigensolver(mtx, stx)
    es, vs = sl.eigh(mtx, b=stx)
    assert np.allclose(w, es) and np.allclose(np.abs(v), np.abs(vs))
    wl, vl = orc.lapack_generalized_eigensolver_lowest(mtx, stx, 4)
    assert np.allclose(wl, es[:4]) and np.allclose(np.abs(vl), np.abs(vs[:, :4]))
    q = orc.lapack_qr(mtx)
    qn, _ = np.linalg.qr(mtx)
    assert np.allclose(np.abs(q), np.abs(qn)) and np.allclose(q.T @ q, np.eye(50))
    # upper triangle only is read (DSYEV 'U'): garbage in the strict lower triangle changes nothing
    junk = mtx.copy()
    junk[np.tril_indices(50, -1)] = 99.0
    w2, _ = orc.lapack_generalized_eigensolver(junk)
    assert np.allclose(w2, sl.eigh(mtx)[0])


def test_lapack_solver_matmul_sort():
    rng = np.random.default_rng(3)
    a = rng.standard_normal((30, 30)); a = a + a.T
    b = rng.standard_normal(30)
    x = orc.lapack_solver(a, b)
    assert np.allclose(a @ x, b)
    p, q = rng.standard_normal((7, 5)), rng.standard_normal((7, 4))
    assert np.allclose(orc.lapack_matmul("T", "N", p, q), p.T @ q)
    assert np.allclose(orc.lapack_matmul("N", "T", p, rng.standard_normal((3, 5)) * 0 + 1), p @ np.ones((5, 3)))
    assert np.allclose(orc.lapack_matmul("N", "N", p.T, q, 2.0), 2.0 * p.T @ q)
    assert np.allclose(orc.lapack_matrix_vector("N", a, b), a @ b)
    v = np.array([3.0, 1.0, 2.0, 0.5])
    s, keys = orc.lapack_sort("I", v)
    assert list(s) == [0.5, 1.0, 2.0, 3.0] and list(keys) == [4, 2, 3, 1]
    pre = orc.generate_preconditioner(v, 2)
    assert pre.shape == (4, 2) and pre[3, 0] == 1.0 and pre[1, 1] == 1.0 and pre.sum() == 2.0
    # duplicated diagonal values: reference undefined -> stable order by index
    pre = orc.generate_preconditioner(np.array([2.0, 1.0, 1.0, 3.0]), 3)
    assert pre[1, 0] == 1.0 and pre[2, 1] == 1.0 and pre[0, 2] == 1.0


def test_generate_diagonal_dominant_properties():
    a = orc.generate_diagonal_dominant(64, 1e-4, seed=5)
    assert np.array_equal(a, a.T)                      # array_utils.f90:101-102
    assert np.array_equal(np.diag(a), np.arange(1, 65))  # :107
    off = a[~np.eye(64, dtype=bool)]
    assert off.min() >= 0.0 and off.max() < 1e-4
    b = orc.generate_diagonal_dominant(64, 1e-4, 1.0, seed=5)
    assert np.array_equal(np.diag(b), np.ones(64)) and np.array_equal(b - np.diag(np.diag(b)), a - np.diag(np.diag(a)))
    assert not np.array_equal(a, orc.generate_diagonal_dominant(64, 1e-4, seed=6))
    assert abs(off.mean() / 1e-4 - 0.5) < 0.02


def test_free_matmul_matches_operator_matrix():
    x = np.random.default_rng(0).standard_normal((40, 3))
    for op in (orc.OP_BENCHMARK_MTX, orc.OP_IDENTITY, orc.OP_TEST_STX):
        m = orc.operator_matrix(op, 40)
        assert np.allclose(orc.free_matmul(op, x), m @ x, rtol=1e-13, atol=1e-13)
