"""Multi-GPU partitioning of the annotate path: reads are independent (reference src/annotate/annotator.rs:122-135), so
the read stream is split on the host, one rank per GPU, with no data-path collective.  The only collectives are the final
sums of the ProgressTracker counters {total, kept, dropped} (annotator.rs:109-113) and of the per-barcode row counts
(what `trim` turns into one FASTQ per label) -- NCCL over NVLink on the GPU box, gloo in the CPU tests."""
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


def label_histogram(rows: np.ndarray, groups) -> np.ndarray:
    """Rows per (group, barcode label); slot 0 of every group counts its flank-only rows.  int64[sum(n_barcodes + 1)]."""
    sizes = [len(g["barcodes"]) + 1 for g in groups]
    base = np.concatenate([[0], np.cumsum(sizes)])
    hist = np.zeros(int(base[-1]), dtype=np.int64)
    if len(rows):
        idx = base[rows["group_idx"].astype(np.int64)] + rows["label_idx"].astype(np.int64) + 1
        np.add.at(hist, idx, 1)
    return hist


def all_reduce_label_counts(hist: np.ndarray, backend_device="cuda") -> np.ndarray:
    """Sum the per-barcode row counts over all ranks (one all-reduce of a few hundred int64)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(hist, dtype=np.int64)).to(backend_device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
