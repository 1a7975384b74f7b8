"""CPU check of the lemma behind the GPU pre-filter, using only the oracle's exact cost rows:

every end position whose full-flank cost is <= k must lie inside (a) a read-end window or (b) a window derived from a
candidate run of the N-free run Q that also survives the second-run pre-check.  This is a restatement in numpy of what
k_flank_filter computes (scan phase and pre-check phase), so a flaw in the argument shows up here without a GPU."""
import numpy as np
import pytest

import barbell_b200 as bb
from barbell_b200 import synth
import oracle_lib as O


def rc(s):
    return bytes({65: 84, 84: 65, 67: 71, 71: 67, 78: 78}.get(c, 78) for c in reversed(s))


def runs_of(flags):
    idx = np.flatnonzero(flags)
    if len(idx) == 0:
        return []
    cuts = np.flatnonzero(np.diff(idx) > 1)
    starts = np.concatenate([[idx[0]], idx[cuts + 1]])
    ends = np.concatenate([idx[cuts], [idx[-1]]])
    return list(zip(starts.tolist(), ends.tolist()))


def n_free_runs(flank):
    out, s = [], None
    for i, c in enumerate(flank + b"N"):
        if c != ord("N") and s is None:
            s = i
        elif c == ord("N") and s is not None:
            out.append((s, i - s)); s = None
    return out


@pytest.mark.parametrize("kit,kw", [("SQK-NBD114-96", {}), ("SQK-RBK114-96", dict(max_flank_errors=5))])
def test_prefilter_windows_cover_every_subthreshold_position(kit, kw):
    gs = bb.GroupSet.from_kit(kit, **kw)
    G = gs.as_dicts()[0]
    P, k = G["flank"], G["k_flank"]
    m = len(P)
    runs = sorted(n_free_runs(P), key=lambda r: -r[1])
    q0, qlen = runs[0]
    q = min(qlen, 15)
    assert 3 * k <= q
    Q = P[q0:q0 + q]
    rest = [(s, l) for s, l in n_free_runs(P[:q0] + b"N" * q + P[q0 + q:])]
    s0, sl = max(rest, key=lambda r: r[1])
    qs = min(sl, 15)
    s0 = s0 if s0 > q0 else s0 + (sl - qs)
    S = P[s0:s0 + qs]
    Df, Dr = m - (q0 + q), m - q0
    d_f, d_r = (s0 + qs) - (q0 + q), q0 - s0
    b, o, _ = synth.make_reads([G], 400, (300, 2500), seed=11, p_mut=0.08)
    span = m + k
    needed = missing = kept_runs = all_runs = 0
    for r in range(len(o) - 1):
        t = b[int(o[r]):int(o[r + 1])].tobytes()
        n = len(t)
        rows = {0: O.bottom_row(P, t, 0.4), 1: O.bottom_row(P, rc(t), 0.4)}           # exact, with overhang, per strand frame
        blockQ = {0: O.bottom_row(Q, t), 1: O.bottom_row(rc(Q), t)}                    # what k_flank_filter tracks (forward text)
        blockS = {0: O.bottom_row(S, t), 1: O.bottom_row(rc(S), t)}
        for strand in (0, 1):
            cov = np.zeros(n + m + 1, bool)
            cov[:span + 1] = True
            cov[max(0, n - span):] = True
            cq, cs = blockQ[strand], blockS[strand]
            flags = cq <= k
            flags[0] = False
            for ps, pe in runs_of(flags):
                all_runs += 1
                d = d_f if strand == 0 else d_r
                slo, shi = ps + d - k, pe + d + k
                if slo >= 1 and shi <= n and cq[ps:pe + 1].min() + cs[slo:shi + 1].min() > k:
                    continue                                                            # dropped by the pre-check
                kept_runs += 1
                lo, hi = (ps + Df - k, pe + Df + k) if strand == 0 else (n - pe + Dr - k, n - ps + Dr + k)
                lo, hi = max(lo, 1), min(hi, n)
                if lo <= hi:
                    cov[lo:hi + 1] = True
            need = np.flatnonzero(rows[strand] <= k)
            needed += len(need)
            missing += int((~cov[need]).sum())
    assert needed > 500
    assert missing == 0, f"{missing} of {needed} sub-threshold positions fall outside every verification window"
    assert kept_runs < 0.5 * all_runs          # the pre-check really removes most random candidates
