"""Shared test-case helpers: rebuild the golden inputs (seeded) and load the frozen outputs."""
import hashlib
import json
import os

import numpy as np

import barbell_b200 as bb
from barbell_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
META = json.load(open(os.path.join(GOLD, "golden.json")))


def groups_for(case):
    if "kit" in case:
        return bb.GroupSet.from_kit(case["kit"], case.get("use_extended", False), case.get("max_flank_errors"))
    return bb.GroupSet.from_fasta([os.path.join(GOLD, f) for f in case["fasta"]], [0, 1])


def load_case(name):
    case = META["cases"][name]
    gs = groups_for(case)
    rl = case["read_len"]
    bases, offsets, truth = synth.make_reads(gs.as_dicts(), case["n"], tuple(rl) if isinstance(rl, list) else rl, seed=case["seed"])
    assert hashlib.sha256(bases.tobytes()).hexdigest() == case["bases_sha256"], "synthetic generator drifted from the golden inputs"
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return gs, bases, offsets, z["rows"], z["hits"]
