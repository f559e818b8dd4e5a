"""torch.distributed plumbing for the one-process-per-GPU launch (bench.py, multi-rank drivers).

The hot path has no collective: reads are independent and each rank decodes its own shard.
What crosses ranks is bookkeeping only -- a barrier, the max over ranks of the device timings,
and the integer read_type / per-barcode tallies (the reference merges them on the host,
barcode_hmm.c:356-384).  Works with backend "nccl" (GPU) and "gloo" (CPU tests).
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None):
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
        global _cpu_group
        # a host-side group: a rank that waits in an NCCL barrier keeps a spinning kernel resident on ITS GPU, which
        # would take SMs away from a rank that drives all GPUs of the box from one process (bench.py's strong-scaling leg)
        _cpu_group = dist.new_group(backend="gloo") if backend != "gloo" else None
    return rank, world, local


_cpu_group = None


def cpu_barrier():
    """Barrier that waits on the host only (gloo): the GPUs of the waiting ranks stay idle."""
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier(group=_cpu_group) if _cpu_group is not None else dist.barrier()


def _dev():
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def barrier():
    if dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(values):
    t = torch.tensor(list(values), dtype=torch.float64, device=_dev())
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def merge_tallies(counts):
    """Sum integer tallies (read_type counts, reads per barcode, ...) over ranks."""
    t = torch.as_tensor(np.asarray(counts, dtype=np.int64), device=_dev())
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def rank_slice(n_total, rank, world):
    """Contiguous [start, end) of a global read set for one rank (same rule as the reference's
    static slices, barcode_hmm.c:1911-1922: floor(n/T) each, the tail goes to the last)."""
    interval = int(n_total / world)
    start = rank * interval
    end = n_total if rank == world - 1 else start + interval
    return start, end


def finalize():
    if dist.is_initialized():
        dist.destroy_process_group()
