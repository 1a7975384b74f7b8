import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); 
import barbell_b200 as bb
from barbell_b200 import synth
gs=bb.GroupSet.from_kit("SQK-NBD114-96"); G=gs.as_dicts()
for n in (20000, 50000, 100000):
    b,o,_=synth.make_reads(G,n,10000,seed=synth.SEED0+2)
    an=bb.Annotator(gs)
    tb=torch.from_numpy(b).cuda(); to=torch.from_numpy(o.astype(np.int64)).cuda()
    for it in range(2):
        l0=an.kernel_launches()
        nr=an.annotate_device(tb.data_ptr(), to.data_ptr(), n, len(b), torch.cuda.current_stream().cuda_stream)
        print(n, "rows", nr, "launches", an.kernel_launches()-l0, an.stage_ms())
    an.close()
