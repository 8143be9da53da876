"""Multi-GPU parity check (run under torchrun on the GPU box, not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_gpu_check.py

Every rank solves the same problems through the row-block sharded handle API and compares with the oracle
(iteration count +-1, eigenvalues 1e-10 relative, eigenvectors up to sign, residual <= tol)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fortran_davidson_b200 as fd  # noqa: E402
from fortran_davidson_b200 import dist as fdist  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    rank, world, _ = fdist.env_world()
    s = fdist.create_solver()
    cases = [(50, 1e-4, 3, 20, 1e-8, False, "DPR"), (1000, 1e-2, 3, 10, 1e-10, False, "DPR"),
             (1000, 1e-2, 3, 10, 1e-10, True, "DPR"), (3000, 5e-2, 10, 100, 1e-8, False, "DPR"),
             (700, 1e-2, 3, 10, 1e-9, True, "GJD"), (5000, 1e-3, 8, None, 1e-8, True, "DPR")]
    for (n, sp, L, md, tol, gev, method) in cases:
        s.clear(1)
        s.generate_diagonal_dominant(0, n, sp, None, 0)
        if gev:
            s.generate_diagonal_dominant(1, n, sp, 1.0, 1)
        ev, vec, iters = s.solve(L, method, 200, tol, md)
        A = orc.generate_diagonal_dominant(n, sp, None, 0)
        B = orc.generate_diagonal_dominant(n, sp, 1.0, 1) if gev else None
        r = orc.generalized_eigensolver(A, L, method, 200, tol, md, B)
        assert abs(iters - r.iters) <= 1, (n, iters, r.iters)
        assert np.abs(ev - r.eigenvalues).max() / np.abs(ev).max() < 1e-10
        Bm = B if gev else np.eye(n)
        for j in range(L):
            sg = np.sign(vec[:, j] @ (Bm @ r.eigenvectors[:, j]))
            assert np.abs(sg * vec[:, j] - r.eigenvectors[:, j]).max() < 1e-8
            assert np.linalg.norm(A @ vec[:, j] - ev[j] * (Bm @ vec[:, j])) < max(tol, 1e-8)
        r0, r1 = s.rows()
        evl, vecl, itl = s.solve(L, method, 200, tol, md, local=True)   # row-sharded result
        assert itl == iters and np.array_equal(evl, ev) and np.array_equal(vecl[:r1 - r0], vec[r0:r1])
        blk = s.download(0)
        assert np.array_equal(blk, A[r0:r1])
        if rank == 0:
            print("ok n=%d gev=%s %s iters=%d (oracle %d) rows/rank0=%d" % (n, gev, method, iters, r.iters, r1 - r0))
    # matrix-free, sharded
    s.clear(1)
    s.set_operator(0, 1000, fd.OP_BENCHMARK_MTX)
    s.set_operator(1, 1000, fd.OP_IDENTITY)
    ev, vec, iters = s.solve(3, "DPR", 1000, 1e-8, 20)
    r = orc.generalized_eigensolver_free(1000, orc.OP_BENCHMARK_MTX, orc.OP_IDENTITY, 3, "DPR", 1000, 1e-8, 20)
    assert abs(iters - r.iters) <= 1 and np.abs(ev - r.eigenvalues).max() / np.abs(ev).max() < 1e-10
    if rank == 0:
        print("ok free benchmark n=1000 iters=%d" % iters)
        print("comm", s.comm_info())
        print("DIST_GPU_CHECK_PASSED world=%d" % world)
    s.close()


if __name__ == "__main__":
    main()
