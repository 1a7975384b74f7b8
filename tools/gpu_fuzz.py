#!/usr/bin/env python3
"""Adversarial GPU-vs-oracle fuzz: reads assembled from random DNA, whole / cut / mutated / reverse-complemented tags, repeats of
flank pieces, N runs, lower case and garbage bytes.  python tools/gpu_fuzz.py [batches] [seed]"""
import os, sys, time
os.environ.setdefault("ORC_PER_READ", "2048")      # repeat-heavy reads report hundreds of flank matches per read
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import barbell_b200 as bb
from barbell_b200 import synth
import oracle_lib as O

n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ACGT = np.frombuffer(b"ACGT", np.uint8)
JUNK = np.frombuffer(b"NNNNRYKMSWacgtn-*xU", np.uint8)
KITS = [("SQK-NBD114-96", {}), ("SQK-RBK114-96", dict(max_flank_errors=5)), ("SQK-RBK114-96", {}), ("SQK-16S024", {}), ("EXP-PBC096", {}),
        ("SQK-RBK114-96", dict(use_extended=True)), ("SQK-MAB114-24", {}), ("SQK-LWB001", {})]


def piece(G):
    g = G[int(rng.integers(len(G)))]
    tags = synth.full_tags(g)
    tag = tags[int(rng.integers(len(tags)))]
    kind = int(rng.integers(12))
    if kind == 0:
        p = rng.choice(ACGT, int(rng.integers(0, 600)))
    elif kind == 1:
        p = tag
    elif kind == 2:
        p = synth.mutate(rng, tag, float(rng.choice([0.03, 0.08, 0.2])))
    elif kind == 3:
        p = tag[int(rng.integers(0, len(tag))):]                      # cut at the front
    elif kind == 4:
        p = tag[:int(rng.integers(0, len(tag) + 1))]                  # cut at the back
    elif kind == 5:
        f = np.frombuffer(bytes(g["flank"]), np.uint8)
        a = int(rng.integers(0, len(f) - 4)); b = int(rng.integers(a + 3, min(len(f), a + 30) + 1))
        p = np.tile(np.where(f[a:b] == ord("N"), rng.choice(ACGT, b - a), f[a:b]), int(rng.integers(1, 40)))   # repeat of a flank piece
    elif kind == 6:
        p = np.full(int(rng.integers(1, 120)), ord("N"), np.uint8)
    elif kind == 7:
        p = rng.choice(JUNK, int(rng.integers(1, 60)))
    elif kind == 8:
        p = np.concatenate([tag, tag])                                  # two tags back to back
    elif kind == 9:
        p = np.full(int(rng.integers(1, 200)), rng.choice(ACGT), np.uint8)   # homopolymer
    elif kind == 10:
        p = (tag | 0x20).astype(np.uint8)                               # lower case
    else:
        p = rng.choice(ACGT, int(rng.integers(0, 60)))
    if rng.random() < 0.4:
        p = synth.revcomp(np.asarray(p, np.uint8))
    return np.asarray(p, np.uint8)


bad = 0
t0 = time.time()
for it in range(n_batches):
    kit, kw = KITS[it % len(KITS)]
    gs = bb.GroupSet.from_kit(kit, **kw)
    G = gs.as_dicts()
    reads = []
    for r in range(300):
        n_p = int(rng.integers(0, 7))
        rd = np.concatenate([piece(G) for _ in range(n_p)] + [np.zeros(0, np.uint8)])
        reads.append(rd[:8000])
    bases = np.concatenate(reads)
    offsets = np.concatenate([[0], np.cumsum([len(x) for x in reads])]).astype(np.uint64)
    prm = dict(alpha=float(rng.choice([0.4, 0.5, 1.0])), min_score=float(rng.choice([0.2, 0.0])), min_score_diff=float(rng.choice([0.1, 0.0])))
    out = []
    for uf in (True, False):
        an = bb.Annotator(gs, use_filter=uf, **prm)
        out.append((an.annotate(bases, offsets), an.flank_hits()))
        an.close()
    try:
        rows_o = O.demux_batch(G, bases, offsets, cap_per_read=2048, **prm)
        hits_o = O.flank_hits_batch(G, bases, offsets, alpha=prm["alpha"], cap_per_read=2048)
    except RuntimeError as e:
        print(f"batch {it} [{kit} {kw}]: oracle buffer overflow ({e}); filter==exact: {out[0][0].tobytes() == out[1][0].tobytes()}")
        bad += out[0][0].tobytes() != out[1][0].tobytes()
        continue
    ok = (out[0][0].tobytes() == out[1][0].tobytes() == rows_o.tobytes()) and out[0][1].shape == hits_o.shape and (out[0][1] == hits_o).all()
    print(f"batch {it} [{kit} {kw} {prm}] bases={len(bases)} rows={len(rows_o)} hits={len(hits_o)} {'OK' if ok else 'MISMATCH'}", flush=True)
    bad += not ok
print(f"{n_batches} batches, {bad} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
