#!/usr/bin/env python3
"""Write a synthetic FASTQ (SURVEY.md 8d recipe) for end-to-end CLI runs: make_fastq.py OUT.fastq N_READS [READ_LEN] [KIT]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import barbell_b200 as bb
from barbell_b200 import synth
out, n = sys.argv[1], int(sys.argv[2])
L = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
kit = sys.argv[4] if len(sys.argv) > 4 else "SQK-NBD114-96"
gs = bb.GroupSet.from_kit(kit)
b, o, _ = synth.make_reads(gs.as_dicts(), n, L, seed=synth.SEED0 + 2)
qual = b"I" * L
with open(out, "wb", buffering=1 << 24) as f:
    for i in range(n):
        s = b[int(o[i]):int(o[i + 1])].tobytes()
        f.write(b"@read_%d ch=1\n" % i); f.write(s); f.write(b"\n+\n"); f.write(qual[:len(s)]); f.write(b"\n")
