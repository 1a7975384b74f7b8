#!/usr/bin/env python3
"""Freeze oracle outputs as golden fixtures under tests/golden/ (run here, committed; nothing reads /root/reference).

The reference itself (Rust + sassy from crates.io) cannot be built in this environment, so these vectors pin the
ORACLE (and through it the GPU path) against regressions; the oracle in turn is pinned against the reference's own
known-answer tests in tests/test_oracle_kats.py.
"""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import barbell_b200 as bb
from barbell_b200 import synth
import oracle_lib as O

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {
    # BASELINE.json configs[0]: SQK-NBD114-96 on a 1k-read synthetic FASTQ
    "nbd_1k": dict(kit="SQK-NBD114-96", n=1000, read_len=(600, 4000), seed=synth.SEED0 + 1),
    "rbk_k5": dict(kit="SQK-RBK114-96", n=300, read_len=(600, 4000), seed=synth.SEED0 + 3, max_flank_errors=5),
    "rbk_ext": dict(kit="SQK-RBK114-96", n=200, read_len=(600, 4000), seed=4, use_extended=True),
    "ald": dict(fasta=["ald_left.fasta", "ald_right.fasta"], n=200, read_len=(600, 4000), seed=5),
}


def groups_for(case):
    if "kit" in case:
        return bb.GroupSet.from_kit(case["kit"], case.get("use_extended", False), case.get("max_flank_errors"))
    return bb.GroupSet.from_fasta([os.path.join(GOLD, f) for f in case["fasta"]], [0, 1])


def main():
    meta = {}
    for name, case in CASES.items():
        gs = groups_for(case)
        bases, offsets, _ = synth.make_reads(gs.as_dicts(), case["n"], case["read_len"], seed=case["seed"])
        rows = O.demux_batch(gs.as_dicts(), bases, offsets)
        hits = O.flank_hits_batch(gs.as_dicts(), bases, offsets)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), rows=rows, hits=hits)
        meta[name] = dict(case, bases_sha256=hashlib.sha256(bases.tobytes()).hexdigest(), n_rows=int(len(rows)), n_hits=int(len(hits)))
        if name == "nbd_1k":
            ids = [f"read_{i}" for i in range(case["n"])]
            with open(os.path.join(GOLD, "nbd_1k.annotation.tsv"), "w") as f:
                f.write(bb.rows_to_tsv(rows, gs, ids))
        print(name, len(rows), "rows", len(hits), "hits")
    # sassy-level vectors
    vecs = []
    import random
    rnd = random.Random(11)
    pats = [b"ATTGCTAAGGTTAA" + b"N" * 24 + b"CAGCACCT", b"AAAAACCCAAAA", b"GCTTGGGTGTTTAACC" + b"N" * 24 + b"GTTTTCGCATTTATCGTGAAACGCTTTCGCGTTTTTCGTGCGCCGCTTCA"]
    for p in pats:
        for _ in range(6):
            t = bytearray(rnd.choice(b"ACGT") for _ in range(rnd.randint(20, 300)))
            s = rnd.randint(-10, max(0, len(t) - len(p) + 10))
            for i, c in enumerate(p):
                if 0 <= s + i < len(t) and c != ord("N") and rnd.random() > 0.08:
                    t[s + i] = c
            k = {46: 4, 12: 3, 90: 20}[len(p)]
            ms = O.search(p, bytes(t), k, alpha=0.4)
            vecs.append(dict(pattern=p.decode(), text=bytes(t).decode(), k=k, alpha=0.4,
                             matches=[dict(ts=m.text_start, te=m.text_end, ps=m.pattern_start, pe=m.pattern_end, cost=m.cost,
                                           strand=m.strand, cigar=m.cigar()) for m in ms]))
    json.dump(dict(cases=meta, search_vectors=vecs), open(os.path.join(GOLD, "golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
