#!/usr/bin/env python3
"""Device-resident throughput of SQK-NBD114-96 over read length (about 0.5 GB of bases per batch): python tools/read_length_sweep.py"""
import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import barbell_b200 as bb
from barbell_b200 import synth
gs = bb.GroupSet.from_kit("SQK-NBD114-96")
for L, n in ((500, 1000000), (800, 600000), (1500, 300000), (3000, 150000)):
    b, o, _ = synth.make_reads(gs.as_dicts(), n, L, seed=5)
    tb = torch.from_numpy(b).cuda(); to = torch.from_numpy(o.astype(np.int64)).cuda()
    an = bb.Annotator(gs)
    st = torch.cuda.current_stream().cuda_stream
    for it in range(3):
        nr = an.annotate_device(tb.data_ptr(), to.data_ptr(), n, len(b), st)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for it in range(5):
        an.annotate_device(tb.data_ptr(), to.data_ptr(), n, len(b), st)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"len {L} reads {n} bases {len(b)/1e6:.0f} MB rows {nr}: wall {dt*1e3:.2f} ms ({len(b)/dt/1e9:.1f} Gbases/s, {n/dt/1e6:.1f} M reads/s)", {k: round(v, 3) for k, v in an.stage_ms().items()}, flush=True)
    an.close()
