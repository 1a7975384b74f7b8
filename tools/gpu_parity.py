#!/usr/bin/env python3
"""Quick GPU-vs-oracle parity probe (development tool; the real parity tests live in tests/test_gpu_parity.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import barbell_b200 as bb
from barbell_b200 import synth
import oracle_lib as O


def compare(name, gs, bases, offs, **kw):
    G = gs.as_dicts()
    t = time.time(); rows_o = O.demux_batch(G, bases, offs, **kw); to = time.time() - t
    hits_o = O.flank_hits_batch(G, bases, offs, alpha=kw.get("alpha", 0.4))
    an = bb.Annotator(gs, alpha=kw.get("alpha", 0.4))
    t = time.time(); rows_g = an.annotate(bases, offs); tg = time.time() - t
    hits_g = an.flank_hits()
    ok_h = hits_o.shape == hits_g.shape and (hits_o == hits_g).all()
    ok_r = rows_o.tobytes() == rows_g.tobytes()
    print(f"[{name}] reads={len(offs)-1} hits oracle={len(hits_o)} gpu={len(hits_g)} {'OK' if ok_h else 'MISMATCH'}; "
          f"rows oracle={len(rows_o)} gpu={len(rows_g)} {'OK' if ok_r else 'MISMATCH'}; oracle {to:.2f}s gpu {tg:.3f}s stages={an.stage_ms()}")
    if not ok_h:
        so = set(map(tuple, hits_o.tolist())); sg = set(map(tuple, hits_g.tolist()))
        print("  only oracle:", sorted(so - sg)[:10]); print("  only gpu:", sorted(sg - so)[:10])
    if not ok_r:
        n = min(len(rows_o), len(rows_g)); shown = 0
        for i in range(n):
            if rows_o[i].tobytes() != rows_g[i].tobytes():
                print("  row", i, "\n   oracle", rows_o[i], "\n   gpu   ", rows_g[i]); shown += 1
                if shown >= 5: break
    an.close()
    return ok_h and ok_r


def main():
    ok = True
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    b, o, _ = synth.make_reads(gs.as_dicts(), 1000, (600, 4000), seed=synth.SEED0 + 1)
    ok &= compare("NBD 1k", gs, b, o)
    b, o, _ = synth.make_reads(gs.as_dicts(), 200, 10000, seed=synth.SEED0 + 2)
    ok &= compare("NBD 10kb", gs, b, o)
    b, o, _ = synth.make_reads(gs.as_dicts(), 400, (0, 120), seed=3)
    ok &= compare("NBD tiny reads", gs, b, o)
    gs = bb.GroupSet.from_kit("SQK-RBK114-96", max_flank_errors=5)
    b, o, _ = synth.make_reads(gs.as_dicts(), 500, (600, 4000), seed=synth.SEED0 + 3)
    ok &= compare("RBK k=5", gs, b, o)
    gs = bb.GroupSet.from_kit("SQK-RBK114-96", use_extended=True)
    b, o, _ = synth.make_reads(gs.as_dicts(), 300, (600, 4000), seed=4)
    ok &= compare("RBK ext auto-k", gs, b, o)
    ex = os.path.join(ROOT, "tests", "golden")
    gs = bb.GroupSet.from_fasta([os.path.join(ex, "ald_left.fasta"), os.path.join(ex, "ald_right.fasta")], [0, 1])
    b, o, _ = synth.make_reads(gs.as_dicts(), 300, (600, 4000), seed=5)
    ok &= compare("ALD dual", gs, b, o)
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
