"""The text-along-word barcode routine (barbell_b200/csrc/barcode_rows.cuh -- the source the GPU kernel k_barcode_rows compiles)
built for the host and compared with the oracle: best local minimum per pattern (searcher.rs:282-301), traceback, Lodhi score
and map_pat_to_text_with_cost (cigar_parse.rs:6-68), for every number of shared leading rows and under every search policy
(S1 plateau side, S2 traceback order, S5 tie between equal minima).  No GPU needed."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "barcode_rows_emu.cpp")
SO = os.path.join(HERE, "emu", "libbarcode_rows_emu.so")
HDRS = [os.path.join(HERE, "..", "barbell_b200", "csrc", "barcode_rows.cuh")]


@pytest.fixture(scope="module")
def emu():
    newest = max([os.path.getmtime(SRC)] + [os.path.getmtime(h) for h in HDRS])
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO, SRC], check=True)
    lib = C.CDLL(SO)
    lib.emu_barcode_rows.restype = C.c_int
    lib.emu_barcode_rows.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    return lib


def mutate(rng, seq, p_sub, p_ins, p_del, alphabet=b"ACGT"):
    out = bytearray()
    for ch in seq:
        r = rng.random()
        if r < p_del:
            continue
        if r < p_del + p_sub:
            out.append(rng.choice(alphabet))
        else:
            out.append(ch)
        while rng.random() < p_ins:
            out.append(rng.choice(alphabet))
    return bytes(out)


def rand_seq(rng, n, alphabet=b"ACGT"):
    return bytes(rng.choice(alphabet) for _ in range(n))


def run_rows(lib, pattern, region, pb0, pb1, lane=0, P=0, pol=0, words=0):
    out = (C.c_int32 * 12)()
    sc = C.c_double()
    rc = lib.emu_barcode_rows(pattern, len(pattern), region, len(region), pb0, pb1, lane, P, pol, words, out, C.byref(sc))
    assert rc == 0, rc
    keys = ["cbest", "jend", "ts", "cnt", "i_first", "i_last", "j_first", "j_last", "sub_cost", "n_ops", "variant", "replayed"]
    d = dict(zip(keys, list(out)))
    d["s"] = sc.value
    return d


def expect(pattern, region, pol=0):
    with O.policy(pol):
        ms = O.search(pattern, region, len(pattern), alpha=-1.0, rc=False)
    best = None
    for m in ms:                                   # lowest cost, first seen wins (searcher.rs:294-300); S5 flips the tie
        if best is None or m.cost < best.cost or (m.cost == best.cost and (pol & O.POL_S5_LAST)):
            best = m
    return best


def check(lib, pattern, region, pb0, pb1, lane=0, P=0, pol=0, words=0):
    best = expect(pattern, region, pol)
    got = run_rows(lib, pattern, region, pb0, pb1, lane, P, pol, words)
    mitm = run_rows(lib, pattern, region, pb0, pb1, lane, P, pol, words | 16)       # half of the records resident at a time
    for k in ("cbest", "jend", "ts", "cnt", "i_first", "i_last", "j_first", "j_last", "sub_cost", "n_ops", "s", "replayed"):
        assert mitm[k] == got[k], (k, pattern, region, pb0, pb1, P, pol, got, mitm)
    assert best is not None
    ctx = (pattern, region, pb0, pb1, P, pol, best, got)
    assert got["cbest"] == best.cost, ctx
    assert got["jend"] == best.text_end, ctx
    assert got["ts"] == best.text_start, ctx
    assert got["n_ops"] == len(best.ops), ctx
    want_s = O.lib().orc_lodhi(best.ops, len(best.ops))
    assert np.float64(got["s"]).tobytes() == np.float64(want_s).tobytes(), ctx + (want_s,)
    mp = best.map_pat_to_text_with_cost(pb0, pb1)
    if mp is None:
        assert got["cnt"] == 0, ctx
    else:
        assert got["cnt"] > 0, ctx
        assert ((got["i_first"], got["i_last"] + 1), (got["j_first"], got["j_last"] + 1), got["sub_cost"]) == mp, ctx + (mp,)
    return got


POLICIES = [0, 1, 2, 4, 3, 5, 6, 7]


@pytest.mark.parametrize("pol", POLICIES)
def test_native_geometry_random(emu, pol):
    """SQK-NBD114-96 geometry: 10 + 24 + 8 pattern rows, region = 10 + barcode + 9 bases around the mask; shared rows 0..L."""
    rng = random.Random(1234 + pol)
    left, right = b"AAGGTTAA"[-10:].rjust(10, b"T"), b"CAGCACCT"
    n_long = n_replayed = 0
    for it in range(2500 if pol == 0 else 800):
        bar = rand_seq(rng, 24)
        pattern = left + bar + right
        kind = it % 6
        if kind == 0:
            text_bar = bar
        elif kind == 1:
            text_bar = mutate(rng, bar, 0.04, 0.02, 0.02)
        elif kind == 2:
            text_bar = mutate(rng, bar, 0.15, 0.08, 0.08)
        elif kind == 3:
            text_bar = rand_seq(rng, rng.randint(18, 30))           # some other barcode
        elif kind == 4:
            text_bar = mutate(rng, bar, 0.02, 0.35, 0.02)           # many inserted bases: paths longer than 48 ops
        else:
            text_bar = mutate(rng, bar, 0.3, 0.0, 0.3)
        region = mutate(rng, left, 0.05, 0.02, 0.02) + text_bar + mutate(rng, right + b"A", 0.05, 0.02, 0.02)
        if rng.random() < 0.1:
            region = region[rng.randint(0, 12):]
        if rng.random() < 0.1:
            region = region[:max(1, len(region) - rng.randint(0, 12))]
        if rng.random() < 0.1:
            region = bytes(rng.choice(b"NRYacgtn") if rng.random() < 0.1 else c for c in region)
        region = region[:64]
        if not region:
            continue
        P = rng.choice([0, 8, 10, 10, 10, rng.randint(0, 42)])
        got = check(emu, pattern, region, 10, 33, lane=it % 32, P=P, pol=pol)
        n_long += got["n_ops"] > 48
        n_replayed += got["replayed"]
    assert n_long > 15 and n_replayed > 5


@pytest.mark.parametrize("L", [1, 2, 8, 24, 41, 42, 43, 44, 47, 48, 49, 56, 63, 64])
def test_pattern_lengths_and_ranges(emu, L):
    rng = random.Random(77 + L)
    for it in range(400):
        alphabet = b"ACGT" if it % 3 else b"ACGTNRYKM"
        pattern = rand_seq(rng, L, alphabet)
        rn = rng.randint(1, 64)
        if it % 2:
            core = mutate(rng, pattern, 0.1, 0.05, 0.05)
            region = (rand_seq(rng, rng.randint(0, 10)) + core + rand_seq(rng, rng.randint(0, 10)))[:rn] or b"A"
        else:
            region = rand_seq(rng, rn, b"ACGTN")
        pb0 = rng.randint(0, L - 1)
        pb1 = rng.randint(pb0, L)
        check(emu, pattern, region, pb0, pb1, lane=it % 32, P=rng.randint(0, L), pol=rng.choice(POLICIES))


def test_degenerate_regions(emu):
    for pattern, region in [(b"ACGTACGTAC", b"A"), (b"ACGTACGTAC", b"T"), (b"AAAAAAAAAA", b"CCCCCCCCCCCC"), (b"ACGT" * 10, b"ACGT" * 10),
                            (b"N" * 42, b"ACGT" * 11), (b"ACGT" * 10 + b"AC", b"N" * 50), (b"A" * 42, b"A" * 64), (b"A" * 64, b"C" * 64),
                            (b"A" * 64, b"X" * 64), (b"ACGT" * 16, b"ACGT" * 16), (b"A" * 30, b"")]:
        for P in (0, 3, len(pattern)):
            for pol in POLICIES:
                if region:
                    check(emu, pattern, region, 3, len(pattern) - 2, P=P, pol=pol)
                else:
                    got = run_rows(emu, pattern, region, 3, len(pattern) - 2, P=P, pol=pol)
                    assert got["cbest"] == len(pattern) and got["jend"] == 0 and got["ts"] == 0 and got["n_ops"] == len(pattern)


def test_long_regions_three_text_words(emu):
    """Regions of 65..160 bases (large automatic flank k: custom 115-bp tags) run with three text words; short regions forced
    through the three-word variant must give the same answers as the one-word one."""
    rng = random.Random(4242)
    for it in range(700):
        L = rng.choice([24, 42, 43, 44, 64])
        pattern = rand_seq(rng, L, b"ACGT" if it % 4 else b"ACGTNRY")
        rn = rng.randint(65, 160) if it % 3 else rng.randint(1, 64)
        if it % 2:
            core = mutate(rng, pattern, 0.1, 0.08, 0.05)
            lead = rand_seq(rng, rng.randint(0, max(0, rn - len(core))))
            region = (lead + core + rand_seq(rng, 160))[:rn]
        else:
            region = rand_seq(rng, rn, b"ACGTN")
        pb0 = rng.randint(0, L - 1)
        pb1 = rng.randint(pb0, L)
        pol = rng.choice(POLICIES)
        got = check(emu, pattern, region, pb0, pb1, lane=it % 32, P=rng.randint(0, L), pol=pol, words=3)
        assert got["variant"] == 6
        if rn <= 64:
            one = run_rows(emu, pattern, region, pb0, pb1, lane=it % 32, P=0, pol=pol, words=1)
            for k in ("cbest", "jend", "ts", "cnt", "j_first", "j_last", "sub_cost", "n_ops", "s"):
                assert got[k] == one[k], (k, got, one)


def test_reversed_lodhi_is_exact_whenever_the_criterion_says_so(emu):
    """lodhi_exact(s, n_ops) must imply bit-equality of the reversed accumulation with the reference's forward recurrence --
    on adversarial op strings too (long match runs = large scores, lengths around the 53-bit budget)."""
    emu.emu_lodhi_reversed.restype = C.c_int
    emu.emu_lodhi_reversed.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_double)]
    rng = random.Random(99)
    n_ok = n_long_ok = n_rejected = 0
    for it in range(60000):
        n = rng.choice([3, 10, 30, 40, 44, 46, 47, 48, 49, 50, 51, 52, 53, 54, 56, 60, 64, 70])
        p = rng.choice([0.3, 0.6, 0.8, 0.9, 0.97, 1.0])
        ops = bytes(1 if rng.random() < p else 0 for _ in range(n))
        sc = C.c_double()
        ok = emu.emu_lodhi_reversed(ops, n, C.byref(sc))
        want = O.lib().orc_lodhi(bytes(0 if o else 1 for o in ops), n)       # oracle op codes: 0 = match
        if ok:
            assert np.float64(sc.value).tobytes() == np.float64(want).tobytes(), (ops, sc.value, want)
            n_ok += 1
            n_long_ok += n > 48
        else:
            n_rejected += 1
    assert n_ok > 20000 and n_long_ok > 1000 and n_rejected > 5000
