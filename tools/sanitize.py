#!/usr/bin/env python3
"""Small all-kernels run for compute-sanitizer (memcheck / racecheck / initcheck): both scan paths, 1- and 2-word flanks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import barbell_b200 as bb
from barbell_b200 import synth
for kit, kw in (("SQK-NBD114-96", {}), ("SQK-RBK114-96", {}), ("SQK-RBK114-96", dict(max_flank_errors=5))):
    gs = bb.GroupSet.from_kit(kit, **kw)
    b, o, _ = synth.make_reads(gs.as_dicts(), 300, (0, 2500), seed=5)
    for uf in (True, False):
        an = bb.Annotator(gs, use_filter=uf)
        rows = an.annotate(b, o)
        print(kit, kw, "filter" if uf else "exact", len(rows), "rows")
        an.close()

# the barcode stage's other variants: 60-row patterns (regions of more than 48 bases: 16-byte row records), a barcode count that is not a multiple of 32,
# regions with ambiguity codes / garbage bytes (general variant), long insertions (replayed Lodhi recurrence), packed H2D copy
rnd = np.random.default_rng(33)
acgt = np.frombuffer(b"ACGT", np.uint8)
pre, suf = b"GGTCTAGACCATGCTAGGAT", b"TTGACCGATTCAGGCATCAA"
seqs = []
for i in range(37):
    core = bytes(rnd.choice(acgt, 40))
    seqs.append(pre + b"ACGT"[i % 4:i % 4 + 1] + core[1:-1] + b"ACGT"[(i // 4) % 4:(i // 4) % 4 + 1] + suf)
gs = bb.GroupSet.from_seqs([(seqs, [f"X{i}" for i in range(37)], 0)])
b, o, _ = synth.make_reads(gs.as_dicts(), 200, (150, 1500), seed=34)
b = b.copy()
idx = rnd.integers(0, len(b), len(b) // 50)
b[idx] = rnd.choice(np.frombuffer(b"RYKMSWBDHVNacgtn-*x", np.uint8), len(idx))
for kw in (dict(), dict(pack_h2d=True), dict(pack_h2d="crumbs")):
    an = bb.Annotator(gs, **kw)
    rows = an.annotate(b, o)
    print("custom 60-row panel", kw, len(rows), "rows")
    an.close()
gs = bb.GroupSet.from_kit("SQK-NBD114-96")
big, obig, _ = synth.make_reads(gs.as_dicts(), 600, (2000, 5000), seed=35, n_frac=0.02)
for mode in (True, "crumbs"):                            # 2 % N: the crumb format overflows its exception list and falls back
    an = bb.Annotator(gs, pack_h2d=mode)
    print("NBD packed copy", mode, "2 % N:", len(an.annotate(big, obig)), "rows")
    an.close()
clean, oclean, _ = synth.make_reads(gs.as_dicts(), 600, (2000, 5000), seed=36, n_frac=0.001)
an = bb.Annotator(gs, pack_h2d="crumbs")
print("NBD crumb copy, 0.1 % N:", len(an.annotate(clean, oclean)), "rows")
an.close()
# bb_reserve (all-'A' batch through every engine) followed by a real batch through the pipelined form
import torch
an = bb.Annotator(gs)
an.reserve(1000, 1000 * 3000)
hb = torch.from_numpy(clean).pin_memory(); ho = torch.from_numpy(oclean.astype(np.uint64).view(np.int64)).pin_memory()
for rep in range(5):
    an.submit(hb.data_ptr(), ho.data_ptr(), len(oclean) - 1, tag=rep)
    tag, rows = an.collect()
print("reserve + 5 pipelined batches:", len(rows), "rows each")
an.close()
