"""The device algorithm (incremental AV/BV, residuals from stored products, project-out + SVQB
expansion, Jacobi Rayleigh-Ritz) restated in numpy must be subspace-equivalent to the
reference: same iteration count and basis schedule, eigenvalues to 1e-10 relative, eigenvectors
up to sign."""
import numpy as np
import pytest

import device_model as dm
from conftest import case_inputs
from oracle import oracle as orc


def test_jacobi_matches_lapack():
    rng = np.random.default_rng(0)
    for k in (1, 2, 5, 12, 33, 64):
        s = rng.standard_normal((k, k)); s = s + s.T
        w, v = dm.jacobi_eigh(s)
        assert np.allclose(w, np.linalg.eigvalsh(s), rtol=0, atol=1e-12 * max(1.0, np.abs(s).max()) * k)
        assert np.abs(v.T @ v - np.eye(k)).max() < 1e-13 * k
        assert np.abs(s @ v - v * w).max() < 1e-12 * k


def test_round_robin_covers_all_pairs():
    for kp in (2, 4, 6, 12, 32):
        seen = set()
        for r in range(kp - 1):
            pairs = dm.round_robin_pairs(kp, r)
            flat = [x for p in pairs for x in p]
            assert sorted(flat) == list(range(kp))
            seen |= set(pairs)
        assert len(seen) == kp * (kp - 1) // 2


@pytest.mark.parametrize("name", ["matrix_txt_DPR", "readme_std_DPR", "readme_gev_DPR", "test_dense_numpy_gen_DPR",
                                  "main_f90_DPR", "collapse_n1000_DPR", "collapse_n1000_gev_DPR"])
def test_model_parity_with_oracle(name, golden_cases):
    g = golden_cases[name]
    A, B = case_inputs(name)
    ev, X, iters, tk, te = dm.solve_dense(A, g["lowest"], "DPR", g["max_iterations"], g["tolerance"],
                                          g["max_dim_sub"], B)
    assert iters == g["iters"] and list(tk) == g["trace_k"]
    assert np.abs(ev - np.array(g["eigenvalues"])).max() / np.abs(ev).max() < 1e-10
    r = orc.generalized_eigensolver(A, g["lowest"], "DPR", g["max_iterations"], g["tolerance"], g["max_dim_sub"], B)
    for j in range(g["lowest"]):
        s = np.sign(X[:, j] @ r.eigenvectors[:, j])
        assert np.abs(s * X[:, j] - r.eigenvectors[:, j]).max() < 1e-8


GJD_CASES = ["matrix_txt_GJD", "readme_std_GJD", "readme_gev_GJD", "test_dense_numpy_std_GJD",
             "test_dense_numpy_gen_GJD", "main_f90_GJD"]


@pytest.mark.parametrize("name", GJD_CASES)
def test_model_gjd_parity_with_oracle(name, golden_cases):
    """GJD as the device computes it (block MINRES on the projected correction equation): same outer iteration count
    (+-1) and eigenvalues as the reference's dense DSYSV solve."""
    g = golden_cases[name]
    A, B = case_inputs(name)
    ev, X, iters, tk, inner = dm.solve_dense_gjd(A, g["lowest"], g["max_iterations"], g["tolerance"], g["max_dim_sub"], B)
    assert abs(iters - g["iters"]) <= 1
    assert np.abs(ev - np.array(g["eigenvalues"])).max() / np.abs(ev).max() < 1e-10
    assert max(inner) <= dm.gjd_inner_limits(g["tolerance"])[1]


def test_model_gjd_harder_cases():
    for (n, sp, L, md, tol) in [(600, 1e-2, 3, 10, 1e-10), (1000, 5e-2, 4, 40, 1e-8)]:
        A = orc.generate_diagonal_dominant(n, sp, seed=0)
        B = orc.generate_diagonal_dominant(n, sp, 1.0, seed=1)
        for Bm in (None, B):
            r = orc.generalized_eigensolver(A, L, "GJD", 100, tol, md, Bm)
            ev, X, iters, tk, inner = dm.solve_dense_gjd(A, L, 100, tol, md, Bm)
            assert abs(iters - r.iters) <= 1
            assert np.abs(ev - r.eigenvalues).max() / np.abs(ev).max() < 1e-10


@pytest.mark.parametrize("name", ["matrix_txt_DPR", "readme_std_DPR", "readme_gev_DPR", "test_dense_numpy_gen_DPR",
                                  "main_f90_DPR", "collapse_n1000_DPR", "collapse_n1000_gev_DPR"])
def test_model_with_bcgs_pip2_parity_with_oracle(name, golden_cases):
    """r02 block orthonormalisation (two passes of BCGS with the Pythagorean inner product, scaled Cholesky in the
    first pass, (I + E)^-1/2 series in the second, SVQB only when a flag rejects the pass): the expansion spans the
    same subspace as the reference's Householder QR of [V | correction] (davidson.f90:205-209), so iteration count,
    basis schedule, eigenvalues and eigenvectors are the oracle's."""
    g = golden_cases[name]
    A, B = case_inputs(name)
    stats = {}
    ev, X, iters, tk, te = dm.solve_dense(A, g["lowest"], "DPR", g["max_iterations"], g["tolerance"],
                                          g["max_dim_sub"], B, ortho="pip", stats=stats)
    assert iters == g["iters"] and list(tk) == g["trace_k"]
    assert np.abs(ev - np.array(g["eigenvalues"])).max() / np.abs(ev).max() < 1e-10
    r = orc.generalized_eigensolver(A, g["lowest"], "DPR", g["max_iterations"], g["tolerance"], g["max_dim_sub"], B)
    for j in range(g["lowest"]):
        s = np.sign(X[:, j] @ r.eigenvectors[:, j])
        assert np.abs(s * X[:, j] - r.eigenvectors[:, j]).max() < 1e-8
    # these inputs are well conditioned: the fast path must be the one that ran
    assert stats.get("pip_accepted", 0) >= 1 and stats.get("pip_fallbacks", 0) == 0, stats


def test_bcgs_pip2_block_is_orthonormal_and_spans_the_corrections():
    rng = np.random.default_rng(5)
    n, k, b = 400, 24, 12
    V, _ = np.linalg.qr(rng.standard_normal((n, k)))
    C = rng.standard_normal((n, b)) * np.logspace(0, 6, b)[None, :]   # badly scaled columns
    C[:, 3] += 1e3 * V[:, 1]                                          # mostly inside span(V)
    stats = {}
    Q = dm.bcgs_pip2(C, V, stats)
    assert stats == {"pip_accepted": 1}
    assert np.abs(Q.T @ Q - np.eye(b)).max() < 1e-13 and np.abs(V.T @ Q).max() < 1e-13
    # same subspace as a Householder QR of [V | C] restricted to the new block
    Qh = np.linalg.qr(np.hstack([V, C]))[0][:, k:]
    assert np.abs(Qh @ (Qh.T @ Q) - Q).max() < 1e-9


def test_bcgs_pip2_rejects_a_dependent_block_and_falls_back():
    rng = np.random.default_rng(6)
    n, k, b = 300, 16, 8
    V, _ = np.linalg.qr(rng.standard_normal((n, k)))
    C = rng.standard_normal((n, b))
    C[:, 5] = C[:, 2]                       # exactly dependent columns: the Cholesky pivot is unsafe
    stats = {}
    Q = dm.bcgs_pip2(C, V, stats)
    assert stats == {"pip_fallbacks": 1}
    assert np.abs(Q.T @ Q - np.eye(b)).max() < 1e-8 and np.abs(V.T @ Q).max() < 1e-8
    C2 = rng.standard_normal((n, b))
    C2[:, 0] = V[:, 3] + 1e-9 * C2[:, 0]    # a correction that lies in span(V) to round-off: diagonal unsafe
    stats = {}
    dm.bcgs_pip2(C2, V, stats)
    assert stats == {"pip_fallbacks": 1}


def test_tridiagonal_eigensolver_model_matches_lapack():
    rng = np.random.default_rng(11)
    for k in (16, 33, 64, 100):
        M = rng.standard_normal((k, k))
        S = (M + M.T) / 2 + np.diag(np.arange(k, dtype=float))
        stats = {}
        w, Y = dm.tridiag_eigh(S, stats)
        assert stats == {"eigh_tridiag": 1}
        assert np.abs(w - np.linalg.eigvalsh(S)).max() < 1e-12 * np.abs(w).max()
        assert np.abs(Y.T @ Y - np.eye(k)).max() < 1e-13
        assert np.abs(S @ Y - Y * w[None, :]).max() < 1e-12 * np.abs(S).max()
    # the projected matrices of the solver: diagonally dominant with a sorted diagonal, tiny couplings
    S = np.diag(np.sort(rng.uniform(1, 50, 32))) + 1e-4 * (M[:32, :32] + M[:32, :32].T)
    w, Y = dm.tridiag_eigh(S)
    assert np.abs(w - np.linalg.eigvalsh(S)).max() < 1e-13 * np.abs(w).max()


def test_tridiagonal_eigensolver_model_guard_refuses_degenerate_spectra():
    """Independent eigenvector computations cannot separate an exactly degenerate eigenvalue: the guard must see it
    and the Jacobi solver must take over (csrc/trideig.cu, step 4)."""
    rng = np.random.default_rng(12)
    k = 24
    Q, _ = np.linalg.qr(rng.standard_normal((k, k)))
    lam = np.arange(k, dtype=float); lam[5] = lam[6] = lam[7] = 3.0
    S = (Q * lam[None, :]) @ Q.T
    S = (S + S.T) / 2
    stats = {}
    w, Y = dm.tridiag_eigh(S, stats)
    assert stats == {"eigh_jacobi": 1}
    assert np.abs(np.sort(w) - np.sort(lam)).max() < 1e-12 * k
    assert np.abs(Y.T @ Y - np.eye(k)).max() < 1e-12


@pytest.mark.parametrize("name", ["readme_std_DPR", "readme_gev_DPR", "main_f90_DPR", "collapse_n1000_DPR",
                                  "collapse_n1000_gev_DPR"])
def test_model_round2_flow_parity_with_oracle(name, golden_cases):
    """The complete r02 flow -- tridiagonal Rayleigh-Ritz for k >= 16 and BCGS-PIP2 expansion -- against the oracle:
    same iteration count and basis schedule, eigenvalues to 1e-10, eigenvectors to 1e-8 up to sign."""
    g = golden_cases[name]
    A, B = case_inputs(name)
    stats = {}
    ev, X, iters, tk, te = dm.solve_dense(A, g["lowest"], "DPR", g["max_iterations"], g["tolerance"],
                                          g["max_dim_sub"], B, ortho="pip", stats=stats, eigh="tridiag")
    assert iters == g["iters"] and list(tk) == g["trace_k"]
    assert np.abs(ev - np.array(g["eigenvalues"])).max() / np.abs(ev).max() < 1e-10
    r = orc.generalized_eigensolver(A, g["lowest"], "DPR", g["max_iterations"], g["tolerance"], g["max_dim_sub"], B)
    for j in range(g["lowest"]):
        s = np.sign(X[:, j] @ r.eigenvectors[:, j])
        assert np.abs(s * X[:, j] - r.eigenvectors[:, j]).max() < 1e-8
    if max(tk) >= dm.EIGH_TRIDIAG_MIN_K:
        assert stats.get("eigh_tridiag", 0) >= 1, stats


def test_model_gjd_inner_solve_follows_a_tight_outer_tolerance():
    """with fixed inner constants an outer tolerance of 1e-12 was served by a 1e-8 inner solve.  r02 ties
    them: the model (and csrc/gjd.cu) solve the correction equation to min(1e-8, tolerance); the outer iteration
    count stays within one of the oracle's dense DSYSV solve and the result meets the tight tolerance."""
    assert dm.gjd_inner_limits(1e-8) == (1e-8, 40) and dm.gjd_inner_limits(1e-4) == (1e-8, 40)
    assert dm.gjd_inner_limits(1e-12) == (1e-12, 72) and dm.gjd_inner_limits(1e-20)[0] == 1e-14
    A = orc.generate_diagonal_dominant(800, 1e-2, seed=3)
    r = orc.generalized_eigensolver(A, 4, "GJD", 50, 1e-12, 40, None)
    with np.errstate(divide="ignore", invalid="ignore"):   # converged columns are masked after the division
        ev, X, iters, tk, inner = dm.solve_dense_gjd(A, 4, 50, 1e-12, 40, None)
    assert abs(iters - r.iters) <= 1 and max(inner) <= 72
    assert np.abs(ev - r.eigenvalues).max() / np.abs(ev).max() < 1e-12
    assert np.sqrt(((A @ X - X * ev[None, :]) ** 2).sum(axis=0)).max() < 1e-12


def test_model_round2_flow_seeded_sweep_against_oracle():
    """12 seeded problems beyond the named cases (size, `lowest`, sparsity, subspace limit with collapses, with and
    without second_matrix): the r02 device flow takes exactly the oracle's iteration count and basis schedule and
    lands on its eigenvalues to 1e-12."""
    rng = np.random.default_rng(777)
    for t in range(12):
        n = int(rng.integers(200, 900))
        lowest = int(rng.integers(1, 7))
        sp = float(10.0 ** rng.uniform(-3, -1.5))
        gev = bool(t % 2)
        max_dim = int(lowest * rng.integers(2, 9)) if t % 4 == 0 else None
        A = orc.generate_diagonal_dominant(n, sp, seed=1000 + t)
        B = orc.generate_diagonal_dominant(n, sp, 1.0, seed=2000 + t) if gev else None
        r = orc.generalized_eigensolver(A, lowest, "DPR", 80, 1e-8, max_dim, B)
        stats = {}
        ev, X, iters, tk, te = dm.solve_dense(A, lowest, "DPR", 80, 1e-8, max_dim, B, ortho="pip", stats=stats,
                                              eigh="tridiag")
        label = (t, n, lowest, gev, max_dim)
        assert iters == r.iters and list(tk) == [int(k) for k in r.trace_k], label
        assert np.abs(ev - r.eigenvalues).max() <= 1e-12 * np.abs(ev).max(), label
        assert stats.get("pip_fallbacks", 0) == 0, (label, stats)
