"""Host-side plumbing of the row-block sharded solver: one process per GPU, torch.distributed for the
rendezvous (NCCL id broadcast, barriers, max-over-ranks timing).  The data path itself never goes through
torch: the library runs its own NCCL communicator (csrc/comm.cu)."""
import ctypes as C
import os

import numpy as np

from ._lib import check, lib


def partition_rows(n, world_size, rank):
    """[row_begin, row_end) owned by `rank` (dav_partition_rows: contiguous blocks, multiples of 128 rows)."""
    b, e = C.c_int64(), C.c_int64()
    check(lib().dav_partition_rows(C.c_int64(n), C.c_int(world_size), C.c_int(rank), C.byref(b), C.byref(e)))
    return b.value, e.value


def chunk_rows(n, world_size):
    """Rows per rank of the padded all-gather staging buffer (the library's `chunk`)."""
    if world_size <= 1:
        return n
    return ((n + world_size - 1) // world_size + 127) // 128 * 128


def stage_block(x_local, chunk):
    """nl x b row block -> the contiguous chunk x b send buffer of the all-gather (zero padded rows)."""
    nl, b = x_local.shape
    out = np.zeros((chunk, b), order="F")
    out[:nl] = x_local
    return out


def unstage_allgather(stages, n):
    """[rank][chunk x b] receive buffer -> full n x b block (inverse of stage_block on every rank)."""
    chunk = stages[0].shape[0]
    full = np.concatenate(stages, axis=0)
    assert full.shape[0] == chunk * len(stages)
    return np.asfortranarray(full[:n])


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload, src=0):
    """Broadcasts a bytes object from `src` with torch.distributed (any backend)."""
    import torch.distributed as dist
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def bind_cpu_to_gpu(local_rank):
    """Restricts this process to the CPUs next to its GPU (NVML's ideal affinity), so that the page-locked host
    block of a rank is first-touched on the GPU's NUMA node and its upload does not cross the socket link (8 ranks
    uploading 10 GB each).  Best effort: returns the CPU list or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:  # noqa: BLE001  (no NVML, no permission, single socket: nothing to do)
        return None
    return None


def create_solver():
    """DavidsonSolver for this process (RANK / WORLD_SIZE / LOCAL_RANK); initialises torch.distributed (nccl) and
    hands the broadcast NCCL id to the library when WORLD_SIZE > 1."""
    import torch
    import torch.distributed as dist

    from .davidson import DavidsonSolver
    rank, world, local_rank = env_world()
    torch.cuda.set_device(local_rank)
    if world == 1:
        return DavidsonSolver(local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nccl_id = broadcast_bytes(DavidsonSolver.unique_id() if rank == 0 else None)
    return DavidsonSolver(local_rank, rank, world, nccl_id)


def max_over_ranks(value, device="cuda"):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
