"""Inter-GPU exchange self-check + latency table (run under torchrun on a multi-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
        tests/dist_collectives_check.py [n]

Checks the library's own peer-memory all-reduce / gathers against analytically known results and prints the time per
call next to NCCL's (CUDA events, back-to-back calls on the solver stream)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fortran_davidson_b200 import dist as fdist  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    rank, world, _ = fdist.env_world()
    s = fdist.create_solver()
    s.generate_diagonal_dominant(0, min(n, 4096), 1e-4, None, 0)  # sets the row partition of a small problem first
    out = {"world": world, "n": n, "allreduce_us": {}, "nccl_allreduce_us": {}, "gather_us": {}, "gather_packed_us": {}}
    for count in (16, 1024, 8192, 32768, 131072):
        us, err = s.debug_collective(0, count, 50)
        assert err <= 1e-9 * world * world, ("allreduce", count, err)
        out["allreduce_us"][count] = round(us, 2)
        us, err = s.debug_collective(1, count, 50)
        assert err <= 1e-9 * world * world, ("nccl allreduce", count, err)
        out["nccl_allreduce_us"][count] = round(us, 2)
    info = s.comm_info()
    out["transport"] = "peer" if info["peer"] else "nccl"
    if info["peer"]:
        for nn in sorted(set([min(n, 4096), n])):
            s.clear(0)
            s.set_operator(0, nn, 1)  # only the row partition matters here (no matrix memory)
            for b in (16, 32, 64, 128, 200):
                us, err = s.debug_collective(2, b, 10)
                assert err == 0.0, ("gather", nn, b, err)
                out["gather_us"]["%d x %d" % (nn, b)] = round(us, 2)
                us, err = s.debug_collective(3, b, 10)
                assert err == 0.0, ("gather_packed", nn, b, err)
                out["gather_packed_us"]["%d x %d" % (nn, b)] = round(us, 2)
    if rank == 0:
        print(json.dumps(out))
        print("DIST_COLLECTIVES_CHECK_PASSED world=%d transport=%s" % (world, out["transport"]))
    s.close()


if __name__ == "__main__":
    main()
