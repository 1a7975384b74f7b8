"""The oracle against every known-answer vector the reference holds for the hot path."""
import random

import numpy as np
import pytest

import oracle_lib as O

P = b"AAAAACCCAAAA"


def rc(s):
    return bytes({65: 84, 84: 65, 67: 71, 71: 67}.get(c, 78) for c in reversed(s))


# both DP back-ends of the oracle x every setting of the search policies that the reference's tests leave open (S1 plateau side,
# S2 text-only vs pattern-only, S5, S6, S3 rounding): the five known-answer tests must hold under ALL of them -- they are what
# makes a setting admissible, and they are why those policies are switches and not facts
POLICY_FLAGS = [0, 1, 2, 4, 8, 16, 32, 1 | 2, 1 | 2 | 4 | 8, 1 | 2 | 4 | 8 | 16, 1 | 2 | 4 | 8 | 32]


@pytest.fixture(params=[(m, f) for m in (1, 0) for f in POLICY_FLAGS], ids=lambda p: f"{'bitvector' if p[0] else 'naive'}-pol{p[1]}")
def mode(request):
    O.set_policy(request.param[0], request.param[1])
    yield request.param
    O.set_policy(1, 0)


# ---- reference src/annotate/cigar_parse.rs:104-176 (the only KATs at the sassy boundary) ----
def test_cost_extraction_no_edits(mode):
    t = b"GGGGAAAAACCCAAAAGGGGG"
    m = O.search(P, t, 0)[0]
    assert m.map_pat_to_text_with_cost(5, 8)[2] == 0
    m = O.search(rc(P), rc(t), 0)[0]
    assert m.map_pat_to_text_with_cost(5, 8)[2] == 0


def test_cost_extraction_1_edits(mode):
    m = O.search(P, b"GGGGAAAAACGCAAAA", 1)[0]
    assert m.map_pat_to_text_with_cost(5, 8)[2] == 1


def test_cost_extraction_1_edits_overhang_left_flank(mode):
    m = O.search(P, b"ACGCAAAAGGGGGGGGGGGG", 5)[0]
    _, (ts, te), cost = m.map_pat_to_text_with_cost(5, 8)
    assert (cost, ts, te) == (1, 1, 4)


def test_cost_extraction_1_edits_overhang_right_flank(mode):
    m = O.search(P, b"GAAAAACGC", 5)[0]
    _, (ts, te), cost = m.map_pat_to_text_with_cost(5, 8)
    assert (cost, ts, te) == (1, 6, 9)


def test_cost_overhang_including_bar(mode):
    m = O.search(P, b"GCAAAAGGGGGGGGGGGG", 8)[0]
    _, (ts, te), cost = m.map_pat_to_text_with_cost(5, 8)
    assert (cost, ts, te) == (2, 0, 2)


# ---- Lodhi: paper Appendix B worked examples + the perfect scores quoted in SURVEY.md section 8 a8 ----
def test_lodhi_paper_examples():
    assert O.lodhi("XX===") == 0.125
    assert O.lodhi("=X=X=") == 0.03125


def test_lodhi_perfect_scores():
    assert repr(O.lodhi("=" * 44)) == "20.000000000002615"
    assert repr(O.lodhi("=" * 43)) == "19.500000000005116"
    assert repr(O.lodhi("=" * 41)) == "18.500000000019554"


def test_lodhi_matches_definition():
    """S_3(C, 1/2) = sum over triples of match positions of 2^-(span) -- brute force on random op strings."""
    rnd = random.Random(7)
    for _ in range(50):
        ops = "".join(rnd.choice("==X=ID") for _ in range(rnd.randint(0, 30)))
        pos = [i for i, c in enumerate(ops) if c == "="]
        want = sum(0.5 ** (pos[c] - pos[a] + 1) for a in range(len(pos)) for b in range(a + 1, len(pos))
                   for c in range(b + 1, len(pos)))
        assert abs(O.lodhi(ops) - want) < 1e-12


# ---- edit_model.rs:2-11, paper Appendix C: kappa(66) = 20 ----
def test_edit_cut_off():
    L = O.lib()
    assert L.orc_edit_cut_off(66) == 20
    assert L.orc_edit_cut_off(22) == 4
    assert L.orc_edit_cut_off(92) == 31 and L.orc_edit_cut_off(91) == 30
    assert L.orc_edit_cut_off(0) == 0 and L.orc_edit_cut_off(1) == 0


# ---- search semantics: costs are true edit distances, both DP back-ends agree ----
def _edit_row(p, t):
    m, n = len(p), len(t)
    prev = list(range(m + 1))
    out = [m]
    for j in range(1, n + 1):
        cur = [0] * (m + 1)
        for i in range(1, m + 1):
            cur[i] = min(prev[i - 1] + (p[i - 1] != t[j - 1]), prev[i] + 1, cur[i - 1] + 1)
        out.append(cur[m])
        prev = cur
    return out


def test_bottom_row_is_edit_distance():
    rnd = random.Random(1)
    for m in (5, 31, 64, 65, 90, 128, 130):
        p = bytes(rnd.choice(b"ACGT") for _ in range(m))
        t = bytes(rnd.choice(b"ACGT") for _ in range(300))
        assert O.bottom_row(p, t).tolist() == _edit_row(p, t)


def test_matches_are_valid_alignments(mode):
    rnd = random.Random(2 + mode[0])
    pol = mode[1]
    for it in range(60):
        m = rnd.choice([8, 20, 46, 70, 115])
        p = bytes(rnd.choice(b"ACGTN") if rnd.random() < 0.3 else rnd.choice(b"ACGT") for _ in range(m))
        t = bytearray(rnd.choice(b"ACGT") for _ in range(rnd.randint(0, 400)))
        if len(t) > m + 10:     # implant a noisy copy
            s = rnd.randint(0, len(t) - m)
            for i in range(m):
                if rnd.random() > 0.1:
                    t[s + i] = p[i] if p[i] != ord("N") else t[s + i]
        k = rnd.randint(0, m // 3)
        alpha = rnd.choice([-1.0, 0.4, 0.5])
        for mt in O.search(p, bytes(t), k, alpha=alpha):
            assert mt.cost <= k
            # re-score the alignment: ops must be consistent with the sequences and add up to the cost
            path = mt.path()
            tt = bytes(t)
            frame = tt if mt.strand == O.FWD else None
            edits = 0
            for (i, j), op in zip(path, mt.ops):
                if op in (O.OP_MATCH, O.OP_SUB):
                    a, b = p[i:i + 1], tt[j:j + 1]
                    if mt.strand == O.RC:
                        b = rc(b)
                    same = a == b or a == b"N" or b == b"N"
                    assert same == (op == O.OP_MATCH)
                edits += op != O.OP_MATCH
            over = mt.pattern_start + (m - mt.pattern_end)
            if alpha >= 0:
                import math

                def over_cost(t):                     # policy S3: floor (default) / round-to-nearest / ceil of the f32 product
                    v = np.float32(t) * np.float32(alpha)
                    return math.floor(v + np.float32(0.5)) if pol & O.POL_S3_ROUND else math.ceil(v) if pol & O.POL_S3_CEIL else math.floor(v)
                extra = over_cost(mt.pattern_start) + over_cost(m - mt.pattern_end)
            else:
                extra = 0
                assert over == 0
            assert edits + extra == mt.cost, (mt, alpha)


def test_backends_agree_on_random_and_adversarial_text():
    rnd = random.Random(3)
    texts = [bytes(rnd.choice(b"ACGT") for _ in range(500)), b"A" * 300, b"N" * 120, b"ACGT" * 60, b"", b"A", b"acgtnACGTN" * 20]
    pats = [b"ATTGCTAAGGTTAA" + b"N" * 24 + b"CAGCACCT", b"AAAAAAAAAAAAAAAA", b"ACGTACGTAC", bytes(rnd.choice(b"ACGT") for _ in range(100))]
    for t in texts:
        for p in pats:
            for k in (0, 2, 6):
                for alpha in (-1.0, 0.4):
                    for pol in (0, 1 | 2 | 4 | 8 | 16):
                        O.set_policy(1, pol); a = [repr(m) for m in O.search(p, t, k, alpha=alpha)]
                        O.set_policy(0, pol); b = [repr(m) for m in O.search(p, t, k, alpha=alpha)]
                        O.set_policy(1, 0)
                        assert a == b, (t[:20], p[:20], k, alpha, pol)


def test_rc_search_is_forward_search_of_reverse_complement():
    rnd = random.Random(4)
    p = b"ATTGCTAAGGTTAA" + b"N" * 24 + b"CAGCACCT"
    t = bytearray(rnd.choice(b"ACGT") for _ in range(300))
    t[100:146] = b"ATTGCTAAGGTTAACACAAAGACACCGACAACTTTCTTCAGCACCT"
    t = bytes(t)
    fw = [m for m in O.search(p, t, 4, alpha=0.4) if m.strand == O.FWD]
    rv = [m for m in O.search(p, rc(t), 4, alpha=0.4) if m.strand == O.RC]
    assert len(fw) == 1 and len(rv) == 1
    n = len(t)
    assert (rv[0].text_start, rv[0].text_end) == (n - fw[0].text_end, n - fw[0].text_start)
    assert rv[0].ops == fw[0].ops and rv[0].cost == fw[0].cost
    # mask region maps onto the same bases (get_matching_region takes min/max)
    a = fw[0].matching_region(14, 37)
    b = rv[0].matching_region(14, 37)
    assert (n - 1 - b[1], n - 1 - b[0]) == a
