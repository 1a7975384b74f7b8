#!/usr/bin/env python3
"""GPU-vs-oracle parity sweep over EVERY kit preset (39 names, 21 presets; with and without --use-extended) and randomised
parameters (alpha, thresholds, --flank-max-errors, read-length mix, mutation rate, N fraction).  Development tool:
python tools/gpu_parity_sweep.py [reads per case] [seed]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import barbell_b200 as bb
from barbell_b200 import synth
import oracle_lib as O

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 300
names = json.load(open(os.path.join(ROOT, "barbell_b200", "data", "kits.json")))["kit_names"]
names = sorted(n[0] if isinstance(n, (list, tuple)) else n for n in names)
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
bad, done = [], 0
t0 = time.time()
for kit in names:
    for ext in (False, True):
        kw = {}
        if rng.random() < 0.4:
            kw["max_flank_errors"] = int(rng.integers(2, 9))
        try:
            gs = bb.GroupSet.from_kit(kit, use_extended=ext, **kw)
        except bb.BarbellError as e:
            print(f"[{kit} ext={ext} {kw}] not constructible: {e}")
            continue
        G = gs.as_dicts()
        if ext and len(G) == len(bb.GroupSet.from_kit(kit, **kw).as_dicts()):
            continue                                   # no Extended template for this kit: same groups as ext=False
        lens = np.clip(rng.lognormal(7.3, 0.9, n_reads).astype(int), 0, 60000)
        b, o, _ = synth.make_reads(G, n_reads, (200, 3000), seed=int(rng.integers(1, 1 << 30)), p_mut=float(rng.choice([0.0, 0.04, 0.08, 0.15])),
                                   n_frac=float(rng.choice([0.0, 0.001, 0.02])))
        prm = dict(alpha=float(rng.choice([0.4, 0.5, 1.0, 0.25])), min_score=float(rng.choice([0.2, 0.0, 0.5])),
                   min_score_diff=float(rng.choice([0.1, 0.0, 0.3])))
        pol = int(rng.choice([0, 0, 0, int(rng.integers(0, 16)) | int(rng.choice([0, 16, 32]))]))     # search-policy bits (oracle and product alike)
        try:
            an = bb.Annotator(gs, policy=pol, **prm)
        except bb.BarbellError as e:
            print(f"[{kit} ext={ext} {kw}] rejected by bb_set_groups: {e}")
            continue
        rows_g = an.annotate(b, o); hits_g = an.flank_hits(); an.close()
        with O.policy(pol):
            rows_o = O.demux_batch(G, b, o, cap_per_read=32, **prm)
            hits_o = O.flank_hits_batch(G, b, o, alpha=prm["alpha"], cap_per_read=64)
        ok = rows_o.tobytes() == rows_g.tobytes() and hits_o.shape == hits_g.shape and (hits_o == hits_g).all()
        done += 1
        print(f"[{kit} ext={ext} {kw} {prm} policy={pol}] groups={len(G)} k={[g['k_flank'] for g in G]} rows={len(rows_g)} hits={len(hits_g)} {'OK' if ok else 'MISMATCH'}", flush=True)
        if not ok:
            bad.append((kit, ext, kw, prm))
print(f"{done} cases, {len(bad)} mismatches, {time.time() - t0:.0f} s")
print("ALL OK" if not bad else f"FAILURES: {bad}")
sys.exit(1 if bad else 0)
