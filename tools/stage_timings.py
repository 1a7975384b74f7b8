#!/usr/bin/env python3
"""Stage timings (ms, CUDA events inside the library) of one device-resident batch for a few kits, filter on/off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import barbell_b200 as bb
from barbell_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
def ald_panel(n_bar=384):
    """configs[3] style: n_bar random 24-mers in the ald_left / ald_right flanks (tests/golden), Ftag + Rtag"""
    gold = os.path.join(ROOT, "tests", "golden")
    rnd = np.random.default_rng(26)
    gl, gr = bb.GroupSet.from_fasta([gold + "/ald_left.fasta", gold + "/ald_right.fasta"], [0, 1]).as_dicts()
    def panel(g):
        pre, suf = g["flank"][:g["bar_region"][0]], g["flank"][g["bar_region"][1] + 1:]
        out = []
        for i in range(n_bar):
            core = bytes(rnd.choice(np.frombuffer(b"ACGT", np.uint8), 24))
            out.append(pre + b"ACGT"[i % 4:i % 4 + 1] + core[1:-1] + b"ACGT"[(i // 4) % 4:(i // 4) % 4 + 1] + suf)
        return out
    return bb.GroupSet.from_seqs([(panel(gl), [f"L{i}" for i in range(n_bar)], 0), (panel(gr), [f"R{i}" for i in range(n_bar)], 1)])
cases = [("SQK-NBD114-96", {}), ("SQK-RBK114-96", dict(max_flank_errors=5)), ("SQK-RBK114-96", {}), ("SQK-RBK114-96", dict(use_extended=True)),
         ("custom dual-end 384-barcode panel", None)]
sel = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else range(len(cases))
modes = (True,) if len(sys.argv) > 3 else (True, False)
for kit, kw in [cases[i] for i in sel]:
    gs = ald_panel() if kw is None else bb.GroupSet.from_kit(kit, **kw)
    b, o, _ = synth.make_reads(gs.as_dicts(), n, 10000, seed=synth.SEED0 + 2)
    tb = torch.from_numpy(b).cuda(); to = torch.from_numpy(o.astype(np.int64)).cuda()
    for uf in modes:
        an = bb.Annotator(gs, use_filter=uf)
        for it in range(3):
            l0 = an.kernel_launches()
            nr = an.annotate_device(tb.data_ptr(), to.data_ptr(), n, len(b), torch.cuda.current_stream().cuda_stream)
        st = an.stage_ms()
        print(kit, kw, "filter" if uf else "exact ", "reads", n, "rows", nr, "launches", an.kernel_launches() - l0,
              {k: round(v, 3) for k, v in st.items()}, "total %.2f ms" % sum(st.values()))
        an.close()
