import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import barbell_b200 as bb
from barbell_b200 import synth
for kit, kw in (("SQK-NBD114-96", {}), ("SQK-RBK114-96", dict(max_flank_errors=5))):
    gs = bb.GroupSet.from_kit(kit, **kw)
    n = 6400
    b, o, _ = synth.make_reads(gs.as_dicts(), n, 10000, seed=5)
    tb = torch.from_numpy(b).cuda(); to = torch.from_numpy(o.astype(np.int64)).cuda()
    an = bb.Annotator(gs)
    st = torch.cuda.current_stream().cuda_stream
    for it in range(5):
        an.annotate_device(tb.data_ptr(), to.data_ptr(), n, len(b), st)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for it in range(50):
        an.annotate_device(tb.data_ptr(), to.data_ptr(), n, len(b), st)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
    print(kit, kw, "wall per 64 MB batch %.3f ms" % (dt * 1e3), {k: round(v, 3) for k, v in an.stage_ms().items()}, "sum %.3f" % sum(an.stage_ms().values()))
    an.close()
