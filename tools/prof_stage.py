#!/usr/bin/env python3
"""One device-resident batch of a bench.py config through the hot path (for ncu):
   python tools/prof_stage.py [--config nbd] [--reads 20000] [--iters 3]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import barbell_b200 as bb
from barbell_b200 import synth
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="nbd"); ap.add_argument("--reads", type=int, default=20000); ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
cfg = bench.CONFIGS[a.config]
gs = bench.product_groups(cfg)
b, o, _ = synth.make_reads(gs.as_dicts(), a.reads, bench.READ_LEN, seed=synth.SEED0 + 2)
tb = torch.from_numpy(b).cuda(); to = torch.from_numpy(o.astype(np.int64)).cuda()
an = bb.Annotator(gs)
for it in range(a.iters):
    nr = an.annotate_device(tb.data_ptr(), to.data_ptr(), a.reads, len(b), torch.cuda.current_stream().cuda_stream)
print("config", a.config, "reads", a.reads, "rows", nr, an.stage_ms())
an.close()
