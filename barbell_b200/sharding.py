"""Multi-GPU partitioning of the annotate path: reads are independent (reference src/annotate/annotator.rs:122-135), so
the read stream is split on the host, one rank per GPU, with no data-path collective.  The only collective is the final
sum of the ProgressTracker counters {total, kept, dropped} (annotator.rs:109-113) -- NCCL over NVLink on the GPU box,
gloo in the CPU tests."""
import numpy as np


def shard_range(n_reads: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of the read indices for `rank`."""
    base, rem = divmod(n_reads, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def slice_reads(bases: np.ndarray, offsets: np.ndarray, lo: int, hi: int):
    """Sub-batch [lo, hi) with offsets rebased to 0."""
    b0, b1 = int(offsets[lo]), int(offsets[hi])
    return bases[b0:b1], (offsets[lo:hi + 1] - offsets[lo]).astype(np.uint64)


def all_reduce_counters(total: int, kept: int, backend_device="cuda"):
    """Sum {total, kept, dropped} over all ranks (torch.distributed must be initialised)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([total, kept, total - kept], dtype=torch.int64, device=backend_device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    v = t.tolist()
    return dict(total=v[0], kept=v[1], dropped=v[2])
