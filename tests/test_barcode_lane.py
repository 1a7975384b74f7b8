"""The barcode stage's per-lane routine (barbell_b200/csrc/barcode_lane.cuh -- the source the GPU kernel compiles) built for
the host and compared with the oracle: best local minimum per pattern (searcher.rs:282-301), traceback, Lodhi score and
map_pat_to_text_with_cost (cigar_parse.rs:6-68).  No GPU needed: the routine is __host__ __device__."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "barcode_lane_emu.cpp")
SO = os.path.join(HERE, "emu", "libbarcode_lane_emu.so")
HDR = os.path.join(HERE, "..", "barbell_b200", "csrc", "barcode_lane.cuh")


@pytest.fixture(scope="module")
def emu():
    newest = max(os.path.getmtime(SRC), os.path.getmtime(HDR))
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO, SRC], check=True)
    lib = C.CDLL(SO)
    lib.emu_barcode_lane.restype = C.c_int
    lib.emu_barcode_lane.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    return lib


def run_lane(lib, pattern, region, pb0, pb1, lane=0, hist_cols=None):
    out = (C.c_int32 * 12)()
    sc = C.c_double()
    rc = lib.emu_barcode_lane(pattern, len(pattern), region, len(region), pb0, pb1, lane, hist_cols or max(1, len(region)), out, C.byref(sc))
    assert rc == 0
    keys = ["cbest", "jend", "ts", "cnt", "i_first", "i_last", "j_first", "j_last", "sub_cost", "n_ops", "packed", "variant"]
    d = dict(zip(keys, list(out)))
    d["s"] = sc.value
    return d


def expect(pattern, region, pb0, pb1):
    ms = O.search(pattern, region, len(pattern), alpha=-1.0, rc=False)
    best = None
    for m in ms:                                   # lowest cost, first seen wins (searcher.rs:294-300)
        if best is None or m.cost < best.cost:
            best = m
    return best


def check(lib, pattern, region, pb0, pb1, lane=0):
    best = expect(pattern, region, pb0, pb1)
    got = run_lane(lib, pattern, region, pb0, pb1, lane)
    assert best is not None
    ctx = (pattern, region, pb0, pb1, best, got)
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


def test_native_geometry_random(emu):
    """SQK-NBD114-96 geometry: 10 + 24 + 8 pattern rows, region = 10 + barcode + 9 bases around the mask."""
    rng = random.Random(1234)
    left, right = b"AAGGTTAA"[-10:].rjust(10, b"T"), b"CAGCACCT"
    n_long, variants = 0, set()
    for it in range(3000):
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
        got = check(emu, pattern, region, 10, 33, lane=it % 32)
        n_long += got["n_ops"] > 48
        variants.add(got["variant"])
    assert n_long > 50
    assert variants == {1, 2, 5, 6}   # FAST / general (other base sets), each with and without the replayed forward recurrence


@pytest.mark.parametrize("L", [8, 24, 41, 42, 43, 44, 47, 48, 49, 56, 63, 64])
def test_pattern_lengths_and_ranges(emu, L):
    rng = random.Random(77 + L)
    for it in range(400):
        alphabet = b"ACGT" if it % 3 else b"ACGTNRYKM"
        pattern = rand_seq(rng, L, alphabet)
        rn = rng.randint(1, 80)
        if it % 2:
            core = mutate(rng, pattern, 0.1, 0.05, 0.05)
            region = (rand_seq(rng, rng.randint(0, 10)) + core + rand_seq(rng, rng.randint(0, 10)))[:rn] or b"A"
        else:
            region = rand_seq(rng, rn, b"ACGTN")
        pb0 = rng.randint(0, L - 1)
        pb1 = rng.randint(pb0, L)
        check(emu, pattern, region, pb0, pb1, lane=it % 32)


def test_degenerate_regions(emu):
    for pattern, region in [(b"ACGTACGTAC", b"A"), (b"ACGTACGTAC", b"T"), (b"AAAAAAAAAA", b"CCCCCCCCCCCC"), (b"ACGT" * 10, b"ACGT" * 10),
                            (b"N" * 42, b"ACGT" * 11), (b"ACGT" * 10 + b"AC", b"N" * 50), (b"A" * 42, b"A" * 64)]:
        check(emu, pattern, region, 3, len(pattern) - 2)


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
