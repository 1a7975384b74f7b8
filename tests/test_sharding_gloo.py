"""N>1 path on CPU: reads shard across ranks with no data-path collective; the only collective is the final all-reduce of
the {total, kept, dropped} counters (reference src/annotate/annotator.rs:109-113).  world_size=2, gloo, with the oracle
standing in for the per-rank device (host-side logic only)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from barbell_b200 import sharding


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle_lib as O
    gs, bases, offsets, rows, _ = cases.load_case("rbk_k5")
    lo, hi = sharding.shard_range(len(offsets) - 1, rank, world)
    sb, so = sharding.slice_reads(bases, offsets, lo, hi)
    mine = O.demux_batch(gs.as_dicts(), sb, so, n_threads=2)
    mine["read_idx"] += np.uint32(lo)
    kept = len(np.unique(mine["read_idx"]))
    counters = sharding.all_reduce_counters(hi - lo, kept, backend_device="cpu")
    hist = sharding.all_reduce_label_counts(sharding.label_histogram(mine, gs.as_dicts()), backend_device="cpu")
    q.put((rank, mine.tobytes(), counters, hist.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    gs, bases, offsets, rows, _ = cases.load_case("rbk_k5")
    merged = b"".join(g[1] for g in got)
    assert merged == rows.tobytes()
    n = len(offsets) - 1
    kept = len(np.unique(rows["read_idx"]))
    want_hist = sharding.label_histogram(rows, gs.as_dicts())
    assert want_hist.sum() == len(rows) and (want_hist[1:] > 0).sum() > 10
    for g in got:
        assert g[2] == dict(total=n, kept=kept, dropped=n - kept)
        assert g[3] == want_hist.tolist()


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 1000, 1001):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
