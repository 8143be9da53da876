"""GPU parity tests added in round 2: wide expansion blocks, the in-solver widths 32 / 64 / 128 at lowest = 16,
GJD iteration-count equality on a case with >= 4 outer iterations, the register-tile Cholesky + inverse of the block
orthonormalisation, and the multi-GPU parity check under torchrun when the box has more than one GPU.
Tolerances are BASELINE.json's: eigenvalues 1e-10 relative, eigenvectors up to sign 1e-8, residual <= tol."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.linalg as sl

import fortran_davidson_b200 as fd
from fortran_davidson_b200._lib import check, dp, lib
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EV_RTOL = 1e-10
VEC_ATOL = 1e-8


def _solve_handle(A, B, L, method, max_it, tol, md):
    s = fd.DavidsonSolver()
    s.upload(0, A)
    if B is not None:
        s.upload(1, B)
    ev, vec, iters = s.solve(L, method, max_it, tol, md)
    st = s.stats()
    trace = list(st.trace_k[:st.trace_len])
    s.close()
    return ev, vec, iters, trace


def _check_against_oracle(A, B, ev, vec, r, tol):
    assert np.abs(ev - r.eigenvalues).max() / np.abs(r.eigenvalues).max() < EV_RTOL
    Bm = B if B is not None else None
    for j in range(len(ev)):
        bv = Bm @ r.eigenvectors[:, j] if Bm is not None else r.eigenvectors[:, j]
        sg = np.sign(vec[:, j] @ bv)
        assert np.abs(sg * vec[:, j] - r.eigenvectors[:, j]).max() < VEC_ATOL
        bvec = Bm @ vec[:, j] if Bm is not None else vec[:, j]
        assert np.linalg.norm(A @ vec[:, j] - ev[j] * bvec) < max(tol, 1e-8)


# ---------------------------------------------------------------- b x b Cholesky + inverse (block orthonormalisation)
@pytest.mark.parametrize("b", [1, 3, 4, 5, 31, 32, 64, 100, 128, 129, 150, 169, 170, 200, 256])
def test_chol_inv_upper_matches_numpy(b):
    rng = np.random.default_rng(b)
    M = rng.standard_normal((b + 5, b))
    G = np.asfortranarray(M.T @ M / (b + 5) + 0.5 * np.eye(b))
    # only the upper triangle may be read: poison the strictly lower part
    Gp = np.asfortranarray(np.triu(G) + np.tril(np.full((b, b), 7.5), -1))
    T = np.zeros((b, b), order="F")
    flag = C.c_double(-1.0)
    ms = C.c_float(0.0)
    check(lib().dav_debug_chol_inv(C.c_int(b), dp(Gp), dp(T), C.byref(flag), C.byref(ms)))
    assert flag.value == 0.0
    R = np.linalg.cholesky(G).T  # upper, G = R^T R
    ref = np.linalg.inv(R)
    assert np.abs(np.tril(T, -1)).max() == 0.0 if b > 1 else True
    assert np.abs(T - ref).max() <= 1e-12 * np.abs(ref).max() * max(1.0, np.linalg.cond(G))
    assert np.abs(T.T @ G @ T - np.eye(b)).max() < 1e-12


@pytest.mark.parametrize("b", [8, 64, 128, 200])
def test_chol_inv_upper_flags_unsafe_pivot(b):
    rng = np.random.default_rng(1)
    M = rng.standard_normal((b, b - 1))  # rank b-1: the last pivot is round-off
    G = np.asfortranarray(M @ M.T)
    T = np.zeros((b, b), order="F")
    flag = C.c_double(-1.0)
    check(lib().dav_debug_chol_inv(C.c_int(b), dp(G), dp(T), C.byref(flag), None))
    assert flag.value == 1.0


# ---------------------------------------------------------------- the in-solver widths of configs[2]: 32, 64, 128
def test_dense_lowest16_widths_32_64_128_against_oracle():
    """lowest = 16 (BASELINE.json configs[2] at n = 20,000 instead of 100,000): the expansion blocks are 32 and 64
    columns wide, the Rayleigh-Ritz problems 32 / 64 / 128 -- the very kernels of the headline solve."""
    n, L = 20000, 16
    s = fd.DavidsonSolver()
    s.generate_diagonal_dominant(0, n, 1e-4, None, 0)
    ev, vec, iters = s.solve(L, "DPR", 1000, 1e-8, None)
    st = s.stats()
    trace = list(st.trace_k[:st.trace_len])
    s.close()
    A = orc.generate_diagonal_dominant(n, 1e-4, None, 0)
    r = orc.generalized_eigensolver(A, L, "DPR", 1000, 1e-8, None)
    assert iters == r.iters == 3 and trace == [32, 64, 128]
    _check_against_oracle(A, None, ev, vec, r, 1e-8)


def test_dense_lowest16_harder_input_widths_up_to_128():
    """Same widths on an input that needs the 128-column expansion too (sparsity 2e-2: 5 iterations, a collapse)."""
    n, L = 6000, 16
    A = orc.generate_diagonal_dominant(n, 2e-2, None, 3)
    r = orc.generalized_eigensolver(A, L, "DPR", 1000, 1e-8, None)
    ev, vec, iters, trace = _solve_handle(A, None, L, "DPR", 1000, 1e-8, None)
    assert abs(iters - r.iters) <= 1 and trace[:4] == [32, 64, 128, 256] == [int(k) for k in r.trace_k][:4]
    assert max(trace) >= 256  # the k = 128 -> 256 expansion (b = 128) ran
    _check_against_oracle(A, None, ev, vec, r, 1e-8)


# ---------------------------------------------------------------- wide expansion blocks (b >= 170: ADVICE r1, high)
def test_wide_first_expansion_block_lowest_90():
    """lowest = 90: the first expansion block already has 180 columns -- wider than the shared-memory Cholesky of
    the block orthonormalisation can hold; the reference's Householder QR has no such limit."""
    n, L = 2500, 90
    A = orc.generate_diagonal_dominant(n, 1e-3, None, 5)
    r = orc.generalized_eigensolver(A, L, "DPR", 200, 1e-8, None)
    ev, vec, iters, trace = _solve_handle(A, None, L, "DPR", 200, 1e-8, None)
    assert abs(iters - r.iters) <= 1
    assert trace[:2] == [180, 360]
    _check_against_oracle(A, None, ev, vec, r, 1e-8)


def test_wide_third_expansion_block_lowest_32():
    """lowest = 32 with the default max_dim_sub = 320: a solve that is still unconverged at k = 256 expands by a
    256-column block (configs[4] only avoids it by converging earlier)."""
    n, L = 4000, 32
    A = orc.generate_diagonal_dominant(n, 4e-2, None, 6)
    r = orc.generalized_eigensolver(A, L, "DPR", 200, 1e-9, None)
    ev, vec, iters, trace = _solve_handle(A, None, L, "DPR", 200, 1e-9, None)
    assert [int(k) for k in r.trace_k][:4] == [64, 128, 256, 512], r.trace_k
    assert trace[:4] == [64, 128, 256, 512]
    assert abs(iters - r.iters) <= 1
    _check_against_oracle(A, None, ev, vec, r, 1e-9)


# ---------------------------------------------------------------- fused orthonormalisation: passes and the reject path
@pytest.mark.parametrize("kold,b", [(6, 6), (32, 32), (64, 64), (40, 40), (80, 80), (20, 7), (128, 64), (96, 96)])
def test_fused_pip_passes_match_the_separate_kernels(kold, b):
    """dav_debug_pip_small: Z = [-H Tm; Tm] of one BCGS-PIP pass from the fused kernel (mode 0: scaled Cholesky, mode 1:
    series of (I + E)^-1/2) against numpy: Tm^T (C^T C - H^T H) Tm = I and the top block = -H Tm."""
    rng = np.random.default_rng(kold * 131 + b)
    n = 4 * (kold + b) + 50
    V, _ = np.linalg.qr(rng.standard_normal((n, kold)))
    for mode in (0, 1):
        C0 = rng.standard_normal((n, b))
        if mode == 1:  # second pass: already orthonormal to ~1e-9
            C0 = C0 - V @ (V.T @ C0)
            C0, _ = np.linalg.qr(C0)
            C0 = C0 + 1e-9 * rng.standard_normal((n, b)) + 1e-9 * V @ rng.standard_normal((kold, b))
        H = V.T @ C0
        Gall = np.asfortranarray(np.vstack([H, C0.T @ C0]))
        Z = np.zeros((kold + b, b), order="F")
        met = np.zeros(4)
        ok = C.c_int(0)
        check(lib().dav_debug_pip_small(C.c_int(mode), C.c_int(kold), C.c_int(b), dp(Gall), dp(Z), dp(met), C.byref(ok)))
        assert ok.value == 1, (mode, kold, b)
        assert met[2] == 0.0 and met[3] == 0.0, met
        Tm = Z[kold:]
        Gp = C0.T @ C0 - H.T @ H
        assert np.abs(Tm.T @ Gp @ Tm - np.eye(b)).max() < 1e-10 * max(1.0, np.linalg.cond(Gp))
        assert np.abs(Z[:kold] + H @ Tm).max() <= 1e-12 * max(1.0, np.abs(H).max() * np.abs(Tm).max() * b)
        Q = C0 @ Tm + V @ Z[:kold]
        assert np.abs(V.T @ Q).max() < 1e-8 and np.abs(Q.T @ Q - np.eye(b)).max() < 1e-8
        assert abs(met[0] - (np.abs(H) / np.sqrt(np.diag(C0.T @ C0))).max()) <= 1e-12 + 1e-9 * met[0]


def test_fused_pip_flags_a_dependent_block():
    rng = np.random.default_rng(3)
    n, kold, b = 400, 16, 16
    V, _ = np.linalg.qr(rng.standard_normal((n, kold)))
    C0 = rng.standard_normal((n, b))
    C0[:, 5] = V[:, 2] * 3.0  # inside span(V): G' has a round-off diagonal entry
    H = V.T @ C0
    Gall = np.asfortranarray(np.vstack([H, C0.T @ C0]))
    Z = np.zeros((kold + b, b), order="F")
    met = np.zeros(4)
    ok = C.c_int(0)
    check(lib().dav_debug_pip_small(C.c_int(0), C.c_int(kold), C.c_int(b), dp(Gall), dp(Z), dp(met), C.byref(ok)))
    assert ok.value == 1 and met[2] == 1.0
    # second pass input that is not close to orthonormal: |E| >= 1e-5 raises the flag
    C1, _ = np.linalg.qr(C0 - V @ (V.T @ C0) + rng.standard_normal((n, b)))
    C1 = C1 * 1.001
    Gall = np.asfortranarray(np.vstack([V.T @ C1, C1.T @ C1]))
    check(lib().dav_debug_pip_small(C.c_int(1), C.c_int(kold), C.c_int(b), dp(Gall), dp(Z), dp(met), C.byref(ok)))
    assert ok.value == 1 and met[2] == 1.0


def test_rejected_fast_orthonormalisation_is_rebuilt_from_the_corrections():
    """DAV_PIP_FORCE_REJECT=1: every fast pass is judged as failed after its last update has already been enqueued
    (and has overwritten the block); the block must be rebuilt from the corrections by the SVQB loop with the same
    results."""
    n, L = 3000, 10
    A = orc.generate_diagonal_dominant(n, 5e-2, None, 0)
    r = orc.generalized_eigensolver(A, L, "DPR", 1000, 1e-8, 100)
    code = ("import numpy as np, json, sys; sys.path.insert(0, %r)\n"
            "import fortran_davidson_b200 as fd\n"
            "s = fd.DavidsonSolver(); s.generate_diagonal_dominant(0, %d, 5e-2, None, 0)\n"
            "ev, vec, it = s.solve(%d, 'DPR', 1000, 1e-8, 100); st = s.stats()\n"
            "print(json.dumps({'ev': ev.tolist(), 'it': it, 'fb': st.pip_fallbacks}))" % (ROOT, n, L))
    env = dict(os.environ, DAV_PIP_FORCE_REJECT="1")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    import json
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["fb"] >= 3 and abs(d["it"] - r.iters) <= 1
    assert np.abs(np.array(d["ev"]) - r.eigenvalues).max() / np.abs(r.eigenvalues).max() < EV_RTOL


# ---------------------------------------------------------------- GJD: iteration counts equal on >= 4 iterations
@pytest.mark.parametrize("gev", [False, True])
def test_gjd_iteration_count_equal_on_longer_runs(gev):
    n, sp, L, md, tol = 700, 1e-2, 3, 10, 1e-9
    A = orc.generate_diagonal_dominant(n, sp, None, 0)
    B = orc.generate_diagonal_dominant(n, sp, 1.0, 1) if gev else None
    r = orc.generalized_eigensolver(A, L, "GJD", 200, tol, md, B)
    assert r.iters >= 4
    ev, vec, iters, trace = _solve_handle(A, B, L, "GJD", 200, tol, md)
    assert iters == r.iters, (iters, r.iters)
    assert trace == [int(k) for k in r.trace_k]
    _check_against_oracle(A, B, ev, vec, r, tol)


# ---------------------------------------------------------------- multi-GPU parity (collected; skipped on 1 GPU)
def _visible_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("script", ["dist_collectives_check.py", "dist_gpu_check.py"])
def test_sharded_solver_parity_under_torchrun(script):
    """Spawns tests/<script> under torchrun with min(visible GPUs, 8) ranks: the row-block sharded solver (peer-memory
    exchanges) against the oracle on every rank -- iteration counts, eigenvalues 1e-10, eigenvectors 1e-8."""
    ng = min(_visible_gpus(), 8)
    if ng < 2:
        pytest.skip("one GPU visible: the sharded path needs >= 2 (the driver's scaling run covers it)")
    port = 29600 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ng),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script)]
    if script == "dist_collectives_check.py":
        cmd.append("20000")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "_CHECK_PASSED world=%d" % ng in p.stdout


# ---------------------------------------------------------------- lapack_wrapper mirrors entirely on the device (r02)
def test_lapack_qr_rank_deficient_basis_never_fails():
    """DGEQRF + DORGQR (lapack_wrapper.f90:176-236) return an orthonormal basis for any input; the CholeskyQR2 of r01
    raised on a rank-deficient one.  Now: orthonormal, and it contains the span of the input."""
    from fortran_davidson_b200 import lapack_wrapper as lw
    rng = np.random.default_rng(0)
    m, n = 500, 12
    X = rng.standard_normal((m, n))
    X[:, 7] = X[:, 2] - 3.0 * X[:, 5]   # exactly dependent
    X[:, 10] = 0.0                       # zero column
    Q = lw.lapack_qr(X)
    assert np.abs(Q.T @ Q - np.eye(n)).max() < 1e-12
    resid = X - Q @ (Q.T @ X)
    assert np.abs(resid).max() < 1e-10 * np.abs(X).max()
    # full rank: still the QR factor (up to column signs), like the reference
    Y = rng.standard_normal((m, n))
    Qy = lw.lapack_qr(Y)
    Qr, _ = np.linalg.qr(Y)
    assert np.abs(np.abs(Qy.T @ Qr) - np.eye(n)).max() < 1e-10


@pytest.mark.parametrize("n", [1, 7, 33, 300, 2000])
def test_lapack_solver_device_lu(n):
    """lapack_solver (DSYSV 'U', lapack_wrapper.f90:238-277) on the device, no size cap: symmetric INDEFINITE systems,
    only the upper triangle may be read."""
    from fortran_davidson_b200 import lapack_wrapper as lw
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    A = (A + A.T) / 2                      # indefinite
    b = rng.standard_normal(n)
    junk = np.triu(A) + np.tril(np.full((n, n), 123.0), -1)
    x = lw.lapack_solver(junk, b)
    ref = np.linalg.solve(A, b)
    assert np.abs(A @ x - b).max() <= 1e-10 * max(1.0, np.abs(b).max()) * max(1.0, np.linalg.cond(A) * 1e-3)
    assert np.abs(x - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()) * max(1.0, np.linalg.cond(A) * 1e-3)


def test_lapack_solver_singular_matrix_is_reported():
    from fortran_davidson_b200 import lapack_wrapper as lw
    A = np.ones((5, 5))
    with pytest.raises(fd.DavidsonError):
        lw.lapack_solver(np.zeros((4, 4)), np.ones(4))
    assert A.shape == (5, 5)


@pytest.mark.parametrize("n", [1, 2, 5, 1000, 1025, 70001])
def test_lapack_sort_device_bitonic(n):
    from fortran_davidson_b200 import lapack_wrapper as lw
    rng = np.random.default_rng(n)
    v = rng.integers(0, max(2, n // 3), size=n).astype(np.float64)  # many ties
    for id_ in ("I", "D"):
        s, keys = lw.lapack_sort(id_, v)
        ref = np.sort(v) if id_ == "I" else np.sort(v)[::-1]
        assert np.array_equal(s, ref)
        assert sorted(keys.tolist()) == list(range(1, n + 1))
        assert np.array_equal(s[keys - 1], v)      # keys[original position] = rank
        order = np.argsort(keys)                    # original positions in sorted order: ties keep their order
        same = s[1:] == s[:-1]
        assert np.all(order[1:][same] > order[:-1][same])


def test_lapack_matmul_transposed_b_on_device():
    from fortran_davidson_b200 import lapack_wrapper as lw
    rng = np.random.default_rng(4)
    A, B = rng.standard_normal((130, 57)), rng.standard_normal((41, 57))
    out = lw.lapack_matmul("N", "T", A, B, 2.0)
    assert np.abs(out - 2.0 * A @ B.T).max() < 1e-12 * 57
    out = lw.lapack_matmul("T", "T", rng.standard_normal((57, 130)), B)
    assert out.shape == (130, 41)


def test_default_device_selection():
    check(lib().dav_set_default_device(C.c_int(0)))
    assert lib().dav_set_default_device(C.c_int(lib().dav_device_count())) != 0
    check(lib().dav_set_default_device(C.c_int(-1)))


# ---------------------------------------------------------------- device functor operators
class _DevView:
    """Raw device address -> __cuda_array_interface__ (column-major rows x cols view with leading dimension ld)."""

    def __init__(self, ptr, rows, cols, ld):
        self.__cuda_array_interface__ = {"shape": (rows, cols), "typestr": "<f8", "data": (ptr, False), "version": 3,
                                         "strides": (8, 8 * ld)}


def _torch_functor(torch, Adev):
    def fun(xp, ldx, yp, ldy, n, b, r0, nr, stream):
        if nr == 0:
            return
        with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
            x = torch.as_tensor(_DevView(xp, n, b, ldx), device="cuda")
            y = torch.as_tensor(_DevView(yp, nr, b, ldy), device="cuda")
            y.copy_(Adev[r0:r0 + nr] @ x)
    return fun


@pytest.mark.parametrize("give_diag", [True, False])
def test_device_functor_operator_matches_the_dense_solve(give_diag):
    """A user operator that stays on the GPU (dav_matrix_set_device_callback): here a torch matmul enqueued on the
    solver's stream.  Same Ritz pairs and iteration count as the host-callback path and as the oracle's
    matrix-free loop (davidson.f90:277-460: always generalized, DPR, non-sticky convergence)."""
    import torch
    n, L = 1500, 4
    A = orc.generate_diagonal_dominant(n, 1e-2, None, 2)
    B = orc.generate_diagonal_dominant(n, 1e-3, 1.0, 3)
    Ad, Bd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    s = fd.DavidsonSolver()
    s.set_device_callback(0, n, _torch_functor(torch, Ad), np.diag(A).copy() if give_diag else None)
    s.set_device_callback(1, n, _torch_functor(torch, Bd), np.diag(B).copy() if give_diag else None)
    ev, vec, iters = s.solve(L, "DPR", 200, 1e-9, 24)
    s.close()
    h = fd.DavidsonSolver()
    h.set_callback(0, n, lambda x: A @ x, np.diag(A).copy())
    h.set_callback(1, n, lambda x: B @ x, np.diag(B).copy())
    ev2, vec2, iters2 = h.solve(L, "DPR", 200, 1e-9, 24)
    h.close()
    assert iters == iters2
    assert np.abs(ev - ev2).max() / np.abs(ev2).max() < 1e-12
    es = sl.eigh(A, b=B, eigvals_only=True)[:L]
    assert np.abs(ev - es).max() / np.abs(es).max() < EV_RTOL
    for j in range(L):
        assert np.linalg.norm(A @ vec[:, j] - ev[j] * (B @ vec[:, j])) < 1e-8


# ---------------------------------------------------------------- cached handle of the drop-in call
def test_dropin_calls_reuse_the_handle_without_mixing_problems():
    """dav_generalized_eigensolver_dense keeps one handle per process (device block, TMA plan, workspace): repeated
    calls must see the NEW matrix (same size), forget a second_matrix that is no longer passed, survive a size
    change, and dav_release_cache() must leave a working library."""
    n, L = 900, 3
    A1 = orc.generate_diagonal_dominant(n, 1e-2, None, 1)
    A2 = orc.generate_diagonal_dominant(n, 1e-2, None, 2) + 5.0 * np.eye(n)
    B = orc.generate_diagonal_dominant(n, 1e-3, 1.0, 3)
    e1, v1, i1 = fd.generalized_eigensolver(A1, L, "DPR", 200, 1e-9)
    e2, v2, i2 = fd.generalized_eigensolver(A2, L, "DPR", 200, 1e-9)
    eg, vg, ig = fd.generalized_eigensolver(A1, L, "DPR", 200, 1e-9, None, B)
    e1b, v1b, i1b = fd.generalized_eigensolver(A1, L, "DPR", 200, 1e-9)          # no second_matrix any more
    assert np.abs(e1 - sl.eigh(A1, eigvals_only=True)[:L]).max() < 1e-9
    assert np.abs(e2 - sl.eigh(A2, eigvals_only=True)[:L]).max() < 1e-9
    assert np.abs(eg - sl.eigh(A1, b=B, eigvals_only=True)[:L]).max() < 1e-9
    assert np.array_equal(e1, e1b) and np.array_equal(v1, v1b) and i1 == i1b
    As = orc.generate_diagonal_dominant(300, 1e-2, None, 4)                        # another size
    es, _, _ = fd.generalized_eigensolver(As, L, "DPR", 200, 1e-9)
    assert np.abs(es - sl.eigh(As, eigvals_only=True)[:L]).max() < 1e-9
    check(lib().dav_release_cache())
    e1c, _, _ = fd.generalized_eigensolver(A1, L, "DPR", 200, 1e-9)
    assert np.array_equal(e1, e1c)


def test_per_phase_spans_are_opt_in():
    """dav_set_profiling: by default only solve_ms is timed (the ~80 event records of the phase spans cost ~0.1 ms per
    solve); switched on, the phase times add up to (almost) the total."""
    s = fd.DavidsonSolver()
    s.generate_diagonal_dominant(0, 4000, 1e-3, None, 1)
    s.solve(6, "DPR", 100, 1e-9)
    st = s.stats()
    assert st.solve_ms > 0 and st.matvec_ms == 0 and st.rr_ms == 0 and st.kernel_launches > 0
    s.set_profiling(True)
    s.solve(6, "DPR", 100, 1e-9)
    st = s.stats()
    parts = st.matvec_ms + st.rr_ms + st.orth_ms + st.resid_ms + st.proj_ms + st.init_ms + st.output_ms
    # (lower bound loose on purpose: host stalls between two spans count towards solve_ms only)
    assert st.matvec_ms > 0 and st.rr_ms > 0 and 0.2 * st.solve_ms < parts <= 1.05 * st.solve_ms
    s.set_profiling(False)
    s.solve(6, "DPR", 100, 1e-9)
    assert s.stats().rr_ms == 0
    s.close()


def test_symmetric_upload_moves_one_triangle_and_is_exact():
    """One GPU, symmetric host matrix: only the upper triangle crosses PCIe, the device mirrors it -- the resident matrix
    must be bit-identical to the host matrix; an asymmetric host matrix must take the full upload unchanged."""
    n = 3000
    A = orc.generate_diagonal_dominant(n, 1e-3, None, 9)
    s = fd.DavidsonSolver()
    s.upload(0, A)
    moved = C.c_double(0.0)
    check(lib().dav_upload_bytes(s._h, C.byref(moved)))
    assert 0.5 * 8 * n * n <= moved.value <= 0.52 * 8 * n * n
    assert np.array_equal(s.download(0), A)
    # padded leading dimension on the host side
    blk = np.zeros((n + 7, n), order="F")
    blk[:n] = A
    s.upload_ptr(0, n, blk.ctypes.data, n + 7)
    assert np.array_equal(s.download(0), A)
    # not symmetric: everything is uploaded as given
    N = A.copy()
    N[np.tril_indices(n, -1)] += 1.0
    s.upload(0, N)
    check(lib().dav_upload_bytes(s._h, C.byref(moved)))
    assert moved.value == 8.0 * n * n
    assert np.array_equal(s.download(0), N)
    ev, vec, it = fd.generalized_eigensolver(A, 4, "DPR", 200, 1e-9)
    assert np.abs(ev - sl.eigh(A, eigvals_only=True)[:4]).max() < 1e-9
    s.close()
