"""filter / inspect / trim through the C ABI: the reference's own unit-test vectors
(src/filter/pattern.rs:389-937, src/trim/trim.rs:538-802) replayed against the C++ build AND the Python oracle
(oracle/post_oracle.py), plus randomised C++ == oracle comparisons.  Host-only: no GPU needed."""
import ctypes as C
import gzip
import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import post_oracle as P  # noqa: E402

import barbell_b200 as bb  # noqa: E402


class TrimOpts(C.Structure):
    _fields_ = [("add_labels", C.c_int32), ("add_orientation", C.c_int32), ("add_flank", C.c_int32), ("sort_labels", C.c_int32),
                ("only_side", C.c_int32), ("write_full_header", C.c_int32), ("skip_trim", C.c_int32), ("flip", C.c_int32),
                ("gzip", C.c_int32), ("failed_out", C.c_char_p), ("threads", C.c_int32)]


@pytest.fixture(scope="module")
def lib():
    L = bb.lib()
    L.bb_pattern_parse.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    L.bb_filter.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int32, C.POINTER(C.c_uint64), C.c_char_p, C.c_size_t]
    L.bb_inspect.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_size_t]
    L.bb_trim.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int32, C.c_char_p, C.POINTER(TrimOpts), C.POINTER(C.c_uint64), C.c_char_p, C.c_size_t]
    L.bb_kit_filter_patterns.argtypes = [C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_char_p)), C.POINTER(C.c_int32)]
    L.bb_kit_info.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
    return L


def canonical(els):
    out = []
    for e in els:
        ori = "any" if e.orientation is None else ("fw" if e.orientation == "Fwd" else "rc")
        cuts = "|".join(f"{d}({g})" for d, g in e.cuts)
        out.append(f"{e.match_type}[ori={ori},label={e.label if e.label is not None else '*'},ph={e.placeholder if e.placeholder is not None else '-'},"
                   f"rel={e.relative_to or 'none'},range={e.range[0]}..{e.range[1]},cuts={cuts}]")
    return "__".join(out)


def c_parse(lib, s):
    buf, err = C.create_string_buffer(8192), C.create_string_buffer(1024)
    rc = lib.bb_pattern_parse(s.encode(), buf, len(buf), err, len(err))
    return (buf.value.decode(), None) if rc == 0 else (None, err.value.decode())


def row(start, end, mtype="Ftag", label="XXX", strand="Fwd", read_len=500, read_id="test", fs=None, fe=None, rel=0, cuts=None):
    return P.Row(read_id, read_len, rel, start, end, start if fs is None else fs, end if fe is None else fe, 0, 24, mtype, 0, 0, label, strand,
                 list(cuts or []))


def c_filter(lib, tmp_path, rows, patterns, with_dropped=True):
    src, out, drop = tmp_path / "a.tsv", tmp_path / "f.tsv", tmp_path / "d.tsv"
    src.write_text(P.to_tsv(rows))
    arr = (C.c_char_p * len(patterns))(*[p.encode() for p in patterns])
    counts, err = (C.c_uint64 * 3)(), C.create_string_buffer(1024)
    rc = lib.bb_filter(str(src).encode(), str(out).encode(), str(drop).encode() if with_dropped else None, arr, len(patterns), counts, err, len(err))
    assert rc == 0, err.value
    return out.read_text(), (drop.read_text() if with_dropped else None), list(counts)


def both_match(lib, tmp_path, rows, pattern):
    """match_pattern through bb_filter (pattern length == number of annotations) and through the oracle; returns (is_match, cuts)."""
    assert len(P.parse_pattern(pattern)) == len(rows)
    import copy
    ok_o, cuts_o = P.match_pattern(copy.deepcopy(rows), P.parse_pattern(pattern))
    kept, dropped, counts = c_filter(lib, tmp_path, rows, [pattern])
    ok_c = counts[1] == 1
    assert ok_c == ok_o
    got = P.parse_tsv(kept if ok_c else dropped)
    cuts_c = [(i, c) for i, r in enumerate(got) for c, _ in r.cuts]
    assert cuts_c == cuts_o
    return ok_c, cuts_c


# ---------------------------------------------------------------- pattern.rs tests
def test_pattern_macro(lib):                                             # pattern.rs:390-432
    s = "Ftag[fw, *, @left(0..250)]__Fflank[fw, @prev_left(5..100)]__Rtag[?1, fw, @right(0..20)]"
    want = ("Ftag[ori=fw,label=*,ph=-,rel=left,range=0..250,cuts=]__Fflank[ori=fw,label=*,ph=-,rel=prev_left,range=5..100,cuts=]__"
            "Rtag[ori=fw,label=*,ph=1,rel=right,range=0..20,cuts=]")
    assert canonical(P.parse_pattern(s)) == want
    assert c_parse(lib, s) == (want, None)


@pytest.mark.parametrize("s", [
    "Ftag[fw, *, @left(0..250), >>]", "Ftag[<<, rc, ?1, @right(0..250)]", 'Ftag[fw, "BC01", >>3, @left(0..250)]__Rtag[rc, ~NB, <<3]',
    "Ftag[ fw ,*,@left( 0 .. 250 )]", "Fflank[@prev_left(-5..+7), >>, <<2, >>x]", "Rflank[]", "Ftag[fw, @left(0-250)]", "Ftag[@middle(0..1)]",
    "Ftag[fw]__", "Flank[fw]", "Ftag[fw]__Btag[fw]", "nonsense", "Ftag[>]", "Ftag[?x, @left(1..2..3)]", "Ftag[fw, *, @left((0..250))]]",
    # the pattern files the reference's own benchmarks use (benchmarks/data/dual_filter.txt, rapid_filter.txt -- the latter with a
    # range the macro does not understand, which it silently drops)
    "Ftag[rc, *, @left(0..250), >>]__Rtag[<<, rc, *, @right(0..250)]", "Rtag[fw, *, @left(0..250), >>]__Ftag[<<, fw, *, @right(0..250)]",
    "Ftag[fw, *, @left(0 to 250)]",
])
def test_pattern_parser_equals_oracle(lib, s):
    got, err = c_parse(lib, s)
    try:
        want = canonical(P.parse_pattern(s))
    except P.PatternError:
        want = None
    assert got == want, (s, got, err, want)
    if want is None:
        assert err


def test_kit_pattern_sets_parse(lib):                                    # kits.rs:175-236
    for dbl, mx, n_want in [(0, 0, 2), (0, 1, 5), (1, 0, 3), (1, 1, 9)]:
        arr, n = C.POINTER(C.c_char_p)(), C.c_int32()
        assert lib.bb_kit_filter_patterns(dbl, mx, C.byref(arr), C.byref(n)) == 0 and n.value == n_want
        for i in range(n.value):
            s = arr[i].decode()
            assert c_parse(lib, s)[0] == canonical(P.parse_pattern(s))
    name, ranges, dbl, err = C.create_string_buffer(64), C.create_string_buffer(256), C.c_int(), C.create_string_buffer(256)
    assert lib.bb_kit_info(b"SQK-NBD114-96", name, 64, ranges, 256, C.byref(dbl), err, 256) == 0
    assert (name.value, ranges.value, dbl.value) == (b"NB96", b"NB01 - NB96", 1)
    assert lib.bb_kit_info(b"SQK-RBK114.96", name, 64, ranges, 256, C.byref(dbl), err, 256) == 0 and dbl.value == 0
    assert lib.bb_kit_info(b"SQK-NOPE", name, 64, ranges, 256, C.byref(dbl), err, 256) != 0


def test_distance_to_left_end(lib, tmp_path):                            # pattern.rs:434-470
    for start, want in [(0, True), (100, True), (250, True), (251, False)]:
        assert both_match(lib, tmp_path, [row(start, 100)], "Ftag[fw, *, @left(0..250)]")[0] is want


def test_distance_to_right_end(lib, tmp_path):                           # pattern.rs:472-508
    for end, want in [(500, True), (450, True), (250, True), (249, False)]:
        assert both_match(lib, tmp_path, [row(0, end)], "Ftag[fw, *, @right(0..250)]")[0] is want


def test_distance_to_prev_left(lib, tmp_path):                           # pattern.rs:510-574
    for start, want in [(50, False), (100, False), (105, True), (200, True), (201, False)]:
        rows = [row(0, 100), row(start, 200, "Fflank")]
        assert both_match(lib, tmp_path, rows, "Ftag[fw, *, @left(0..250)]__Fflank[fw, @prev_left(5..100)]")[0] is want


def test_placeholders(lib, tmp_path):                                    # pattern.rs:576-741
    pat = "Ftag[fw, ?1, @left(0..250)]__Rtag[fw, ?1, @right(0..250)]"
    assert both_match(lib, tmp_path, [row(0, 100, read_len=250), row(100, 200, "Rtag", read_len=250)], pat)[0]
    assert not both_match(lib, tmp_path, [row(0, 100, read_len=250), row(100, 200, "Rtag", label="yyyy", read_len=250)], pat)[0]
    pat = "Ftag[fw, ?1, @left(0..250)]__Rtag[fw, ?2, @right(0..250)]"
    assert both_match(lib, tmp_path, [row(0, 100, read_len=250), row(100, 200, "Rtag", label="XXX", read_len=250)], pat)[0]
    pat = "Ftag[fw, ?1, @left(0..250)]__Ftag[fw, ?2, @prev_left(0..250)]__Ftag[fw, ?1, @left(0..250)]"
    assert both_match(lib, tmp_path, [row(0, 100), row(100, 200, label="YYY"), row(200, 250)], pat)[0]


def test_patterns_with_cuts(lib, tmp_path):                              # pattern.rs:743-922
    rows = [row(0, 100), row(105, 200, "Fflank", label="@Nothing")]
    assert both_match(lib, tmp_path, rows, "Ftag[fw, *, >>, @left(0..250)]__Fflank[fw, <<, @prev_left(5..100)]") == \
        (True, [(0, ("After", 0)), (1, ("Before", 0))])
    assert both_match(lib, tmp_path, rows, "Ftag[fw, *, >>1, @left(0..250)]__Fflank[fw, <<1, @prev_left(5..100)]") == \
        (True, [(0, ("After", 1)), (1, ("Before", 1))])
    rows3 = rows + [row(400, 490, "Rtag", label="YYY")]
    assert both_match(lib, tmp_path, rows3, "Ftag[fw, *, >>1, @left(0..250)]__Fflank[fw, <<1, @prev_left(5..100)]__Rtag[fw, *, <<2, @right(0..20)]") == \
        (True, [(0, ("After", 1)), (1, ("Before", 1)), (2, ("Before", 2))])


def test_label_and_orientation_checks(lib, tmp_path):
    assert both_match(lib, tmp_path, [row(0, 50, label="NB07")], "Ftag[NB07]")[0]
    assert not both_match(lib, tmp_path, [row(0, 50, label="NB07")], "Ftag[NB08]")[0]
    assert both_match(lib, tmp_path, [row(0, 50, label="NB07")], "Ftag[~B0]")[0]
    assert not both_match(lib, tmp_path, [row(0, 50, label="NB07")], "Ftag[~B1]")[0]
    assert both_match(lib, tmp_path, [row(0, 50, "Fflank", label="flank")], "Fflank[NB08]")[0]          # flanks ignore labels
    assert not both_match(lib, tmp_path, [row(0, 50, strand="Rc")], "Ftag[fw]")[0]
    assert not both_match(lib, tmp_path, [row(0, 50, "Rtag")], "Ftag[*]")[0]


def rand_rows(rng, n_reads):
    rows = []
    for r in range(n_reads):
        L = rng.choice([300, 1000, 5000])
        pos = 0
        for _ in range(rng.choice([1, 1, 1, 2, 2, 3, 4])):
            fs = pos + rng.choice([0, 1, 30, 200, 400, L])
            fe = fs + rng.randint(20, 90)
            if fe > L:
                break
            bs = fs + rng.randint(0, 10)
            mt = rng.choice(["Ftag", "Ftag", "Ftag", "Fflank", "Rtag"])
            lab = "flank" if "flank" in mt else rng.choice(["NB01", "NB02", "NB03"])
            rows.append(P.Row(f"read{r}", L, (fs if fs <= L // 2 else -(L - fs)) or 1, bs, min(fe, bs + 24), fs, fe, bs - fs, bs - fs + 23, mt, rng.randint(0, 4),
                              rng.randint(0, 6), lab, rng.choice(["Fwd", "Fwd", "Rc"])))
            pos = fe
    return rows


def kit_patterns(lib, dbl, mx):
    arr, n = C.POINTER(C.c_char_p)(), C.c_int32()
    lib.bb_kit_filter_patterns(dbl, mx, C.byref(arr), C.byref(n))
    return [arr[i].decode() for i in range(n.value)]


@pytest.mark.parametrize("dbl,mx", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_filter_random_equals_oracle(lib, tmp_path, dbl, mx):
    rng = random.Random(100 + 2 * dbl + mx)
    rows = rand_rows(rng, 400)
    pats = kit_patterns(lib, dbl, mx)
    import copy
    kept_o, dropped_o = P.filter_rows(copy.deepcopy(rows), [P.parse_pattern(p) for p in pats])
    kept, dropped, counts = c_filter(lib, tmp_path, rows, pats)
    assert kept == P.to_tsv(kept_o) and dropped == P.to_tsv(dropped_o)
    n_reads = len({r.read_id for r in rows})
    assert counts == [n_reads, len({r.read_id for r in kept_o}), len({r.read_id for r in dropped_o})]
    assert 0 < counts[1] < counts[0]


def test_filter_empty_inputs(lib, tmp_path):
    kept, dropped, counts = c_filter(lib, tmp_path, [], ["Ftag[fw]"])
    assert (kept, dropped, counts) == ("", "", [0, 0, 0])               # a 0-byte annotation.tsv (no hits) stays 0 bytes
    kept, _, counts = c_filter(lib, tmp_path, [row(0, 10, "Rtag")], ["Ftag[fw]"])
    assert kept == "" and counts == [1, 0, 1]


# ---------------------------------------------------------------- inspect.rs
def test_inspect_equals_oracle(lib, tmp_path, capfd):
    rng = random.Random(5)
    rows = rand_rows(rng, 300)
    import copy
    kept, _ = P.filter_rows(copy.deepcopy(rows), [P.parse_pattern(p) for p in kit_patterns(lib, 1, 1)])
    for name, rr in [("anno", rows), ("filtered", kept)]:
        src, out = tmp_path / f"{name}.tsv", tmp_path / f"{name}.patterns.tsv"
        src.write_text(P.to_tsv(rr))
        err = C.create_string_buffer(512)
        assert lib.bb_inspect(str(src).encode(), 3, str(out).encode(), 250, err, 512) == 0, err.value
        want = "".join(f"{g[0].read_id}\t{P.group_structure(g, 250)}\n" for g in P.group_reads(rr))
        assert out.read_text() == want
    assert P.group_structure([row(10, 60, rel=10), row(100, 150, strand="Rc", cuts=[(("Before", 0), 1)])], 250) == \
        "Ftag[fw, *, @left(0..250)]__Ftag[rc, *, >>, @prev_left(0..250)]"
    assert P.group_structure([row(4800, 4850, read_len=5000, rel=-200)], 250) == "Ftag[fw, *, @right(0..250)]"


# ---------------------------------------------------------------- trim.rs tests
def c_trim(lib, tmp_path, rows, reads, gz_in=False, **kw):
    tsv, fq, out = tmp_path / "filtered.tsv", tmp_path / ("reads.fastq.gz" if gz_in else "reads.fastq"), tmp_path / "trimmed"
    tsv.write_text(P.to_tsv(rows))
    text = "".join(f"@{h}\n{s}\n+\n{q}\n" for h, s, q in reads)
    (gzip.open(fq, "wt") if gz_in else open(fq, "w")).write(text)
    if out.exists():
        for f in out.iterdir():
            f.unlink()
    side = {None: 0, "left": 1, "right": 2}[kw.get("only_side")]
    failed = tmp_path / "failed.txt"
    o = TrimOpts(int(kw.get("add_labels", True)), int(kw.get("add_orientation", True)), int(kw.get("add_flank", True)), int(kw.get("sort_labels", False)),
                 side, 1, int(kw.get("skip_trim", False)), int(kw.get("flip", False)), int(kw.get("gzip", False)), str(failed).encode())
    paths = (C.c_char_p * 1)(str(fq).encode())
    counts, err = (C.c_uint64 * 4)(), C.create_string_buffer(1024)
    rc = lib.bb_trim(str(tsv).encode(), paths, 1, str(out).encode(), C.byref(o), counts, err, len(err))
    assert rc == 0, err.value
    files = {}
    for f in sorted(out.iterdir()):
        files[f.name] = gzip.open(f, "rt").read() if f.name.endswith(".gz") else f.read_text()
    return files, list(counts), failed.read_text()


def test_single_cut_skip_and_flip(lib, tmp_path):                       # trim.rs:542-592, 692-745, 747-802
    seq, qual = "CCCCCCCCAAAACCCCCCCCCCCC", "________IIII____________"
    rows = [row(4, 8, label="Fbar", read_id="read1", read_len=24, cuts=[(("After", 0), 8)]),
            row(12, 16, "Rtag", label="Rbar", read_id="read1", read_len=24, cuts=[(("Before", 0), 12)])]
    kw = dict(sort_labels=True)
    assert P.process_read_and_anno(seq.encode(), qual.encode(), rows, sort_labels=True) == [(b"AAAA", b"IIII", "Fbar_fw__Rbar_fw", "")]
    files, counts, failed = c_trim(lib, tmp_path, rows, [("read1 extra words", seq, qual), ("other", "ACGT", "IIII")], **kw)
    assert files == {"Fbar_fw__Rbar_fw.trimmed.fastq": "@read1 extra words\nAAAA\n+\nIIII\n"} and counts == [2, 1, 0, 0] and failed == ""
    files, _, _ = c_trim(lib, tmp_path, rows, [("read1", seq, qual)], skip_trim=True, **kw)
    assert files == {"Fbar_fw__Rbar_fw.trimmed.fastq": f"@read1\n{seq}\n+\n{qual}\n"}
    seq2, qual2 = "CCCCCCCCAGGCCCCCCCCCCCCC", "________IIIA____________"
    rows[0].strand = "Rc"
    assert P.process_read_and_anno(seq2.encode(), qual2.encode(), rows, flip=True, sort_labels=True) == [(b"GCCT", b"AIII", "Fbar_rc__Rbar_fw", "")]
    files, _, _ = c_trim(lib, tmp_path, rows, [("read1", seq2, qual2)], flip=True, **kw)
    assert files == {"Fbar_rc__Rbar_fw.trimmed.fastq": "@read1\nGCCT\n+\nAIII\n"}
    rows[0].strand = "Fwd"
    files, _, _ = c_trim(lib, tmp_path, rows, [("read1", seq2, qual2)], flip=True, gzip=True, gz_in=True, **kw)
    assert files == {"Fbar_fw__Rbar_fw.trimmed.fastq.gz": "@read1\nAGGC\n+\nIIIA\n"}


def test_two_cut_groups_produce_two_slices(lib, tmp_path):               # trim.rs:594-690
    seq, qual = "CCCCCCCCAAAAAAAAAAAACCCCCCGGCC", "________IIIIIIIIIIII______II__"
    rows = [row(4, 8, label="F1", read_id="read1", read_len=30, cuts=[(("After", 1), 8)]),
            row(20, 24, "Rtag", label="R1", read_id="read1", read_len=30, cuts=[(("Before", 1), 20)]),
            row(24, 26, label="F2", read_id="read1", read_len=30, cuts=[(("After", 2), 26)]),
            row(28, 30, "Rtag", label="R2", read_id="read1", read_len=30, cuts=[(("Before", 2), 28)])]
    assert P.process_read_and_anno(seq.encode(), qual.encode(), rows, sort_labels=True) == \
        [(b"AAAAAAAAAAAA", b"IIIIIIIIIIII", "F1_fw__R1_fw", ""), (b"GG", b"II", "F2_fw__R2_fw", "_1")]
    files, counts, _ = c_trim(lib, tmp_path, rows, [("read1", seq, qual)], sort_labels=True)
    assert files == {"F1_fw__R1_fw.trimmed.fastq": "@read1\nAAAAAAAAAAAA\n+\nIIIIIIIIIIII\n", "F2_fw__R2_fw.trimmed.fastq": "@read1_1\nGG\n+\nII\n"}
    assert counts == [1, 1, 1, 0]


@pytest.mark.parametrize("kw", [dict(add_orientation=False, add_flank=False, only_side="left"), dict(), dict(sort_labels=True),
                                dict(add_labels=False), dict(only_side="right", add_flank=False), dict(flip=True), dict(skip_trim=True)])
def test_trim_random_equals_oracle(lib, tmp_path, kw):
    rng = random.Random(9)
    rows = rand_rows(rng, 300)
    import copy
    kept, _ = P.filter_rows(copy.deepcopy(rows), [P.parse_pattern(p) for p in kit_patterns(lib, 1, 1)])
    by = {}
    for r in kept:
        by.setdefault(r.read_id, []).append(r)
    reads = []
    for rid in sorted({r.read_id for r in rows}, key=lambda s: int(s[4:])):
        L = next(r.read_len for r in rows if r.read_id == rid)
        reads.append((rid + (" desc=1" if rng.random() < 0.3 else ""), "".join(rng.choice("ACGTN") for _ in range(L)), "".join(chr(33 + rng.randint(0, 40)) for _ in range(L))))
    want, n_trim, n_split, n_fail = {}, 0, 0, 0
    lkw = {k: v for k, v in kw.items() if k in ("add_labels", "add_orientation", "add_flank", "sort_labels", "only_side")}
    for h, s, q in reads:
        rid = h.split()[0]
        if rid not in by:
            continue
        res = P.process_read_and_anno(s.encode(), q.encode(), by[rid], skip_trim=kw.get("skip_trim", False), flip=kw.get("flip", False), **lkw)
        n_trim += bool(res); n_split += len(res) > 1; n_fail += not res
        for ts, tq, lab, suf in res:
            desc = h[len(rid):].strip()
            want.setdefault(lab + ".trimmed.fastq", []).append(f"@{rid}{suf}{' ' + desc if desc else ''}\n{ts.decode()}\n+\n{tq.decode()}\n")
    files, counts, failed = c_trim(lib, tmp_path, kept, reads, **kw)
    assert files == {k: "".join(v) for k, v in want.items()}
    assert counts == [len(reads), n_trim, n_split, n_fail] and n_trim > 20
    assert len(failed.split()) == n_fail
    # the FASTQ is cut into chunks that several threads trim; the files must not depend on the chunking (1 KB chunks: records straddle
    # chunks, chunks without a record start) nor on the input being gzip (one sequential reader)
    os.environ["BB_TRIM_CHUNK_KB"] = "1"
    try:
        files2, counts2, failed2 = c_trim(lib, tmp_path, kept, reads, **kw)
    finally:
        del os.environ["BB_TRIM_CHUNK_KB"]
    assert files2 == files and counts2 == counts and failed2 == failed
    files3, counts3, failed3 = c_trim(lib, tmp_path, kept, reads, gz_in=True, **kw)
    assert files3 == files and counts3 == counts and failed3 == failed


def test_cli_filter_inspect_trim(tmp_path):
    exe = os.path.join(ROOT, "barbell_b200", "barbell")
    rng = random.Random(3)
    rows = rand_rows(rng, 50)
    (tmp_path / "a.tsv").write_text(P.to_tsv(rows))
    (tmp_path / "pats.txt").write_text("Ftag[fw, *, @left(0..250), >>]\n\n  Ftag[<<, rc, *, @right(0..250)]  \n")
    r = subprocess.run([exe, "filter", "-i", str(tmp_path / "a.tsv"), "-o", str(tmp_path / "f.tsv"), "-f", str(tmp_path / "pats.txt"), "--dropped", str(tmp_path / "d.tsv"),
                        "--verbose"], capture_output=True, text=True)
    assert r.returncode == 0 and "Filtering successful!" in r.stdout, r.stdout + r.stderr
    logs = [f for f in tmp_path.iterdir() if f.name.startswith("filter.") and f.name.endswith(".log")]     # progress.rs:102-144
    assert len(logs) == 1
    lines = logs[0].read_text().split("\n")
    assert lines[0] == "step\tmetric\tcount" and lines[1].startswith("filter\tTotal:\t") and lines[3].startswith("filter\tDropped:\t")
    import copy
    kept, dropped = P.filter_rows(copy.deepcopy(rows), [P.parse_pattern("Ftag[fw, *, @left(0..250), >>]"), P.parse_pattern("Ftag[<<, rc, *, @right(0..250)]")])
    assert (tmp_path / "f.tsv").read_text() == P.to_tsv(kept) and (tmp_path / "d.tsv").read_text() == P.to_tsv(dropped)
    r = subprocess.run([exe, "inspect", "-i", str(tmp_path / "f.tsv"), "-n", "2", "-o", str(tmp_path / "p.tsv")], capture_output=True, text=True)
    assert r.returncode == 0 and "Found" in r.stdout and "Inspection complete!" in r.stdout
    (tmp_path / "bad.txt").write_text("Ftag[fw]__Bogus[x]\n")
    r = subprocess.run([exe, "filter", "-i", str(tmp_path / "a.tsv"), "-o", str(tmp_path / "f2.tsv"), "-f", str(tmp_path / "bad.txt")], capture_output=True, text=True)
    assert r.returncode == 101 and "Pattern parse error" in r.stderr      # the reference's macro panics (exit 101)
    reads = "".join(f"@{rid}\n{'A' * L}\n+\n{'I' * L}\n" for rid, L in {r.read_id: r.read_len for r in rows}.items())
    (tmp_path / "r.fastq").write_text(reads)
    r = subprocess.run([exe, "trim", "-i", str(tmp_path / "f.tsv"), "-r", str(tmp_path / "r.fastq"), "-o", str(tmp_path / "out"), "--no-orientation", "--no-flanks",
                        "--only-side", "left"], capture_output=True, text=True)
    assert r.returncode == 0 and "Trimming complete!" in r.stdout, r.stdout + r.stderr
    assert any(f.name.endswith(".trimmed.fastq") for f in (tmp_path / "out").iterdir())


def test_python_mirror_of_the_post_stages(tmp_path, capfd):
    """barbell_b200.post: kit_info / kit_patterns / filter / inspect / trim_matches with the reference's argument names."""
    from barbell_b200 import post
    assert post.kit_info("SQK-RBK114-96") == ("RBK096_kit14", ["RBK01 - RBK96", "RBK01 - RBK96"], False)
    assert post.kit_patterns("SQK-NBD114-96") == ["Ftag[fw, *, @left(0..250), >>]", "Ftag[<<, rc, *, @right(0..250)]",
                                                  "Ftag[fw, ?1, @left(0..250), >>]__Ftag[<<, rc, ?1, @right(0..250)]"]
    assert len(post.kit_patterns("SQK-RBK114-96", maximize=True)) == 5
    with pytest.raises(bb.BarbellError):
        post.kit_info("SQK-NOPE")
    rng = random.Random(21)
    rows = rand_rows(rng, 120)
    (tmp_path / "a.tsv").write_text(P.to_tsv(rows))
    counts = post.filter(tmp_path / "a.tsv", tmp_path / "f.tsv", tmp_path / "d.tsv", post.kit_patterns("SQK-NBD114-96", maximize=True))
    import copy
    kept, dropped = P.filter_rows(copy.deepcopy(rows), [P.parse_pattern(p) for p in post.kit_patterns("SQK-NBD114-96", maximize=True)])
    n_reads = len({r.read_id for r in rows})
    assert (tmp_path / "f.tsv").read_text() == P.to_tsv(kept) and counts["kept"] + counts["dropped"] == counts["total"] == n_reads
    with pytest.raises(bb.BarbellError, match="Pattern parse error"):
        post.filter(tmp_path / "a.tsv", tmp_path / "x.tsv", None, ["Ftag[fw]__What[x]"])
    post.inspect(tmp_path / "f.tsv", top_n=2, read_pattern_out=tmp_path / "p.tsv")
    assert len((tmp_path / "p.tsv").read_text().split("\n")) == len({r.read_id for r in kept}) + 1
    reads = "".join(f"@{rid} d\n{'ACGT' * (L // 4)}{'A' * (L % 4)}\n+\n{'I' * L}\n" for rid, L in {r.read_id: r.read_len for r in rows}.items())
    (tmp_path / "r.fastq").write_text(reads)
    c = post.trim_matches(tmp_path / "f.tsv", [tmp_path / "r.fastq"], tmp_path / "out", add_orientation=False, add_flank=False, only_side="left")
    assert c["total"] == n_reads and c["trimmed"] + c["failed"] == len({r.read_id for r in kept}) and c["trimmed"] > 10
    with pytest.raises(bb.BarbellError, match="ambiguous"):
        post.trim_matches(tmp_path / "f.tsv", [tmp_path / "r.fastq"], tmp_path / "out2", sort_labels=True, only_side="left")


def test_trim_reports_failed_writes(lib, tmp_path):
    """A per-label output that cannot be written (here: a symlink to /dev/full, i.e. ENOSPC on flush) must make bb_trim fail with
    BB_ERR_IO instead of returning full counts over truncated files -- the reference aborts (`expect("Failed to write ...")`,
    trim.rs:420-445)."""
    if not os.path.exists("/dev/full"):
        pytest.skip("/dev/full not available")
    seq, qual = "CCCCCCCCAAAACCCCCCCCCCCC", "________IIII____________"
    rows = [row(4, 8, label="Fbar", read_id="read1", read_len=24, cuts=[(("After", 0), 8)]),
            row(12, 16, "Rtag", label="Rbar", read_id="read1", read_len=24, cuts=[(("Before", 0), 12)])]
    tsv, fq, out = tmp_path / "filtered.tsv", tmp_path / "reads.fastq", tmp_path / "trimmed"
    tsv.write_text(P.to_tsv(rows))
    fq.write_text(f"@read1\n{seq}\n+\n{qual}\n")
    out.mkdir()
    os.symlink("/dev/full", out / "Fbar_fw__Rbar_fw.trimmed.fastq")
    o = TrimOpts(1, 1, 1, 1, 0, 1, 0, 0, 0, None)
    paths = (C.c_char_p * 1)(str(fq).encode())
    counts, err = (C.c_uint64 * 4)(), C.create_string_buffer(1024)
    lib.bb_trim.argtypes = None                          # (another module may have bound its own struct class to this symbol)
    rc = lib.bb_trim(str(tsv).encode(), paths, 1, str(out).encode(), C.byref(o), counts, err, C.c_size_t(len(err)))
    assert rc == -6 and b"Failed to write" in err.value, (rc, err.value)
