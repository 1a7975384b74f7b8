#!/usr/bin/env python3
"""Stage timings (ms, CUDA events inside the library) of one device-resident batch for a few kits, filter on/off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import barbell_b200 as bb
from barbell_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
cases = [("SQK-NBD114-96", {}), ("SQK-RBK114-96", dict(max_flank_errors=5)), ("SQK-RBK114-96", dict(use_extended=True))]
for kit, kw in cases:
    gs = bb.GroupSet.from_kit(kit, **kw)
    b, o, _ = synth.make_reads(gs.as_dicts(), n, 10000, seed=synth.SEED0 + 2)
    tb = torch.from_numpy(b).cuda(); to = torch.from_numpy(o.astype(np.int64)).cuda()
    for uf in (True, False):
        an = bb.Annotator(gs, use_filter=uf)
        for it in range(3):
            l0 = an.kernel_launches()
            nr = an.annotate_device(tb.data_ptr(), to.data_ptr(), n, len(b), torch.cuda.current_stream().cuda_stream)
        st = an.stage_ms()
        print(kit, kw, "filter" if uf else "exact ", "reads", n, "rows", nr, "launches", an.kernel_launches() - l0,
              {k: round(v, 3) for k, v in st.items()}, "total %.2f ms" % sum(st.values()))
        an.close()
