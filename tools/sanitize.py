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
