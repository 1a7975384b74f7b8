#!/usr/bin/env python3
"""One device-resident batch through the hot path a few times (for ncu): python tools/prof_stage.py [n_reads] [kit] [flank_max_errors]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import barbell_b200 as bb
from barbell_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
kit = sys.argv[2] if len(sys.argv) > 2 else "SQK-NBD114-96"
kw = dict(max_flank_errors=int(sys.argv[3])) if len(sys.argv) > 3 else {}
gs = bb.GroupSet.from_kit(kit, **kw)
b, o, _ = synth.make_reads(gs.as_dicts(), n, 10000, seed=synth.SEED0 + 2)
tb = torch.from_numpy(b).cuda(); to = torch.from_numpy(o.astype(np.int64)).cuda()
an = bb.Annotator(gs)
for it in range(3):
    nr = an.annotate_device(tb.data_ptr(), to.data_ptr(), n, len(b), torch.cuda.current_stream().cuda_stream)
print("rows", nr, an.stage_ms())
an.close()
