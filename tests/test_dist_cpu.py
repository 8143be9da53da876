"""world_size-2/3 gloo tests (CPU) of the host-side logic of the row-block sharded path: the row partition,
the staged all-gather layout, the bootstrap broadcast, the exchange pattern of one expansion step
(all-gather of the new block, all-reduce of projection / Gram / norm partials) and the WHOLE sharded driver loop
(numpy model of solver.cu with comm.active()) reproduce the single-process numbers and the oracle's golden trace."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fortran_davidson_b200 import dist as fdist
        from oracle import oracle as orc

        # bootstrap broadcast (what carries the NCCL unique id on the GPU box)
        payload = fdist.broadcast_bytes(bytes(range(128)) if rank == 0 else None)
        assert payload == bytes(range(128))
        assert fdist.max_over_ranks(float(rank), device="cpu") == float(world - 1)

        r0, r1 = fdist.partition_rows(n, world, rank)
        chunk = fdist.chunk_rows(n, world)
        spans = [None] * world
        dist.all_gather_object(spans, (r0, r1))
        assert spans[0][0] == 0 and spans[-1][1] == n
        for a, b in zip(spans[:-1], spans[1:]):
            assert a[1] == b[0] and (a[1] - a[0]) == chunk

        # one expansion step, sharded: A rows, V rows, new block Q rows
        A = orc.generate_diagonal_dominant(n, 1e-3, None, 5)
        k, b = 6, 6
        rng = np.random.default_rng(0)
        Vfull = np.linalg.qr(rng.standard_normal((n, k)))[0]
        Cfull = rng.standard_normal((n, b))
        A_loc, V_loc, C_loc = A[r0:r1], Vfull[r0:r1], Cfull[r0:r1]

        def allreduce(x):
            t = torch.from_numpy(np.ascontiguousarray(x))
            dist.all_reduce(t)
            return t.numpy()

        def allgather_rows(x_loc):
            send = torch.from_numpy(np.ascontiguousarray(fdist.stage_block(x_loc, chunk)))
            recv = [torch.empty_like(send) for _ in range(world)]
            dist.all_gather(recv, send)
            return fdist.unstage_allgather([t.numpy() for t in recv], n)

        # project out V (partial Gram all-reduced), column norms, all-gather, block matvec, projection
        G = allreduce(V_loc.T @ C_loc)
        C_loc = C_loc - V_loc @ G
        n2 = allreduce((C_loc * C_loc).sum(axis=0))
        Q_loc = C_loc / np.sqrt(n2)
        Qfull = allgather_rows(Q_loc)
        AQ_loc = A_loc @ Qfull
        P = allreduce(np.hstack([V_loc, Q_loc]).T @ AQ_loc)

        Cref = Cfull - Vfull @ (Vfull.T @ Cfull)
        Qref = Cref / np.linalg.norm(Cref, axis=0)
        assert np.allclose(Qfull, Qref, rtol=1e-13, atol=1e-14)
        Pref = np.hstack([Vfull, Qref]).T @ (A @ Qref)
        assert np.allclose(P, Pref, rtol=1e-12, atol=1e-12)
        assert np.allclose(AQ_loc, (A @ Qref)[r0:r1], rtol=1e-12, atol=1e-13)
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [300, 1000])
def test_sharded_exchange_pattern_gloo(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n % 7
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_bind_cpu_to_gpu_is_best_effort():
    """No NVML / no GPU here: the NUMA binding helper of the multi-rank upload must simply do nothing."""
    import os

    from fortran_davidson_b200.dist import bind_cpu_to_gpu
    before = os.sched_getaffinity(0)
    r = bind_cpu_to_gpu(0)
    assert r is None or set(r) <= set(before)
    assert os.sched_getaffinity(0) == before or r is not None


def _solve_worker(rank, world, port, case, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import json
        import device_model as dm
        from conftest import case_inputs
        from fortran_davidson_b200 import dist as fdist

        g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_cases.json")))[case]
        A, B = case_inputs(case)
        n = A.shape[0]
        r0, r1 = fdist.partition_rows(n, world, rank)
        chunk = fdist.chunk_rows(n, world)
        counts = {"allreduce": 0, "allgather": 0}

        def allreduce(x):
            counts["allreduce"] += 1
            t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
            dist.all_reduce(t)
            return t.numpy()

        def allgather_rows(x_loc):
            counts["allgather"] += 1
            send = torch.from_numpy(np.ascontiguousarray(fdist.stage_block(x_loc, chunk)))
            recv = [torch.empty_like(send) for _ in range(world)]
            dist.all_gather(recv, send)
            return np.array(fdist.unstage_allgather([t.numpy() for t in recv], n))

        stats = {}
        ev, X, iters, tk = dm.solve_dense_sharded(A[r0:r1], r0, n, g["lowest"], g["max_iterations"], g["tolerance"],
                                                  g["max_dim_sub"], None if B is None else B[r0:r1], allreduce,
                                                  allgather_rows, stats=stats)
        # the oracle's trace and eigenpairs (golden file), and the single-process model to round-off
        assert iters == g["iters"] and list(tk) == g["trace_k"], (iters, list(tk))
        assert np.abs(ev - np.array(g["eigenvalues"])).max() / np.abs(ev).max() < 1e-10
        ev1, X1, it1, tk1, _ = dm.solve_dense(A, g["lowest"], "DPR", g["max_iterations"], g["tolerance"],
                                              g["max_dim_sub"], B, ortho="pip", eigh="tridiag")
        assert it1 == iters and np.abs(ev1 - ev).max() < 1e-12 * np.abs(ev).max()
        for j in range(g["lowest"]):
            s = np.sign(X[:, j] @ X1[:, j])
            assert np.abs(s * X[:, j] - X1[:, j]).max() < 1e-8
        resid = A @ X - (X if B is None else B @ X) * ev[None, :]
        assert np.sqrt((resid ** 2).sum(axis=0)).max() < g["tolerance"]
        # every rank ends with the same replicated numbers (deterministic reductions)
        sig = torch.tensor([float(iters), float(ev.sum()), float(np.abs(X).sum())], dtype=torch.float64)
        sigs = [torch.empty_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        assert all(torch.equal(sigs[0], t) for t in sigs)
        assert stats.get("pip_accepted", 0) >= 1
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        import traceback
        q.put((rank, repr(ex) + traceback.format_exc()[-600:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case", [(2, "readme_std_DPR"), (2, "collapse_n1000_gev_DPR"), (3, "collapse_n1000_DPR")])
def test_sharded_solver_flow_gloo(world, case):
    """The whole row-block sharded driver loop (tests/device_model.py::solve_dense_sharded = solver.cu with
    comm.active()) over gloo: rows of A / B / V / AV on their owners, projections + Gram blocks + norm partials
    all-reduced, the new basis block all-gathered through the staged layout -- must reproduce the oracle's golden trace
    and the single-process model on uneven row blocks (the partition gives whole 128-row tiles)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + 3 * world + len(case)
    procs = [ctx.Process(target=_solve_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results
