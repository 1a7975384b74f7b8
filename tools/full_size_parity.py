#!/usr/bin/env python3
"""GPU == oracle on the FULL bench batch (100 000 reads x 10 kb = 1 GB of bases) of every bench configuration: the rows of the whole
batch, byte for byte.  python tools/full_size_parity.py [config ...]   (needs a B200; ~1 minute for all five)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
import barbell_b200 as bb
from barbell_b200 import synth
import oracle_lib as O

names = sys.argv[1:] or ["nbd", "rbk_k5", "rbk_ext", "ald384", "nbd_ext"]
bad = 0
for name in names:
    cfg = bench.CONFIGS[name]
    gs = bench.product_groups(cfg)
    G = gs.as_dicts()
    Go = bench.oracle_groups(cfg)
    n = cfg["reads"]
    bases, offsets, _ = bench.make_batch(G, n, synth.SEED0 + 2)
    an = bb.Annotator(gs)
    t0 = time.perf_counter(); rows_gpu = an.annotate(bases, offsets); t_gpu = time.perf_counter() - t0
    an.close()
    t0 = time.perf_counter(); rows_cpu = O.demux_batch(Go, bases, offsets, n_threads=O.lib().orc_max_threads()); t_cpu = time.perf_counter() - t0
    same = rows_gpu.tobytes() == rows_cpu.tobytes()
    bad += not same
    print(f"{name}: {n} reads, {len(bases)} bases, rows gpu {len(rows_gpu)} / oracle {len(rows_cpu)}  identical={same}  "
          f"(host-API GPU call {t_gpu:.2f} s, oracle on {O.lib().orc_max_threads()} threads {t_cpu:.1f} s)", flush=True)
print("ALL IDENTICAL" if not bad else f"{bad} MISMATCHES")
sys.exit(1 if bad else 0)
