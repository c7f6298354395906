"""Multi-GPU plumbing: read pairs are sharded across ranks in contiguous blocks (keeps the
reference's task order on gather), reference / index / cutoffs are replicated, and the only exchange is the
integer sum of the pileup counts — the multi-GPU replacement of the serial partial-matrix sum at
ConsensusBuilderSimple.java:138-145. torch.distributed is plumbing: NCCL over NVLink on GPUs, gloo in CPU tests."""
import numpy as np


def shard_range(n_pairs, rank, world):
    """Contiguous block of pairs of `rank`: [lo, hi). Blocks differ by at most one pair."""
    base, rem = divmod(int(n_pairs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _DevArray:
    """__cuda_array_interface__ view of a device pointer (zero copy)."""

    def __init__(self, ptr, n, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def allreduce_counts_device(dev_ptr, n_int32, device):
    """In-place NCCL all-reduce (sum, int32) of the counts buffer owned by a vg_ctx."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(_DevArray(dev_ptr, n_int32), device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    torch.cuda.current_stream(device).synchronize()
    return t


def allreduce_counts_host(counts):
    """Same exchange for host arrays (gloo; used by the CPU tests of the host logic)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(counts, np.int32))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()


def gather_records(local_records, world):
    """All-gather of per-pair records in rank order == input order (contiguous shards)."""
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, local_records)
    return np.concatenate(out)
