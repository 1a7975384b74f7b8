"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle and the golden fixtures.
Bit-exact: every field of every annotation row (integer work; the f64 Lodhi score only decides thresholds)."""
import ctypes as C

import numpy as np
import pytest

import barbell_b200 as bb
from barbell_b200 import api, synth
import cases
import oracle_lib as O

pytestmark = pytest.mark.gpu


def _annotator(gs, **kw):
    return bb.Annotator(gs, device=0, **kw)


def _check(gs, bases, offsets, **kw):
    """Both scan paths (lossless pre-filter + window verification, and the exact full-length scan) against the oracle."""
    out = []
    kw = dict(kw)
    cap_override = kw.pop("_cap", None)
    for use_filter in (True, False):
        an = _annotator(gs, use_filter=use_filter, **kw)
        try:
            out.append((an.annotate(bases, offsets), an.flank_hits()))
        finally:
            an.close()
    assert out[0][0].tobytes() == out[1][0].tobytes(), "pre-filter path differs from the exact scan"
    rows, hits = out[0]
    G = gs.as_dicts()
    cap = cap_override or max(16, int(len(bases) // 100 // max(1, len(offsets) - 1)))      # rows / hits per read the oracle buffers may hold
    want = O.demux_batch(G, bases, offsets, alpha=kw.get("alpha", 0.4), min_score=kw.get("min_score", 0.2),
                         min_score_diff=kw.get("min_score_diff", 0.1), cap_per_read=cap)
    want_h = O.flank_hits_batch(G, bases, offsets, alpha=kw.get("alpha", 0.4), cap_per_read=2 * cap)
    assert hits.shape == want_h.shape and (hits == want_h).all(), "flank hit list differs"
    assert rows.tobytes() == want.tobytes(), "annotation rows differ"
    return rows


@pytest.mark.parametrize("name", sorted(cases.META["cases"]))
def test_golden_cases(name):
    """configs[0] (nbd_1k) and the kit/panel variants: GPU == golden == oracle."""
    gs, bases, offsets, rows, hits = cases.load_case(name)
    got = _check(gs, bases, offsets)
    assert got.tobytes() == rows.tobytes()
    if name == "nbd_1k":
        ids = [f"read_{i}" for i in range(len(offsets) - 1)]
        assert bb.rows_to_tsv(got, gs, ids) == open(cases.GOLD + "/nbd_1k.annotation.tsv").read()


def test_10kb_reads():
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    b, o, _ = synth.make_reads(gs.as_dicts(), 500, 10000, seed=21)
    _check(gs, b, o)


def test_ragged_tiny_and_empty_reads():
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    b, o, _ = synth.make_reads(gs.as_dicts(), 600, (0, 130), seed=22)
    assert (np.diff(o.astype(np.int64)) == 0).any()
    _check(gs, b, o)
    # a batch of only empty reads, and an empty batch
    an = _annotator(gs)
    assert len(an.annotate(np.zeros(0, np.uint8), np.zeros(5, np.uint64))) == 0
    assert len(an.annotate(np.zeros(0, np.uint8), np.zeros(1, np.uint64))) == 0
    an.close()


def test_unaligned_read_boundaries_and_long_read():
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    b1, o1, _ = synth.make_reads(gs.as_dicts(), 64, (1, 700), seed=23)
    b2, o2, _ = synth.make_reads(gs.as_dicts(), 3, 300001, seed=24)      # spans several CTA tiles
    bases = np.concatenate([b1, b2])
    offsets = np.concatenate([o1, o2[1:] + o1[-1]])
    _check(gs, bases, offsets)


def test_tags_straddling_scan_tiles_and_chunks():
    """Tags (both orientations, mutated) placed across the 300-base chunk and 76 800-base CTA tile boundaries of long reads:
    the filter's candidate runs are re-scored against the second flank run on the shared text tile, whose halos must reach."""
    for kit, kw in (("SQK-NBD114-96", {}), ("SQK-RBK114-96", dict(max_flank_errors=5))):
        gs = bb.GroupSet.from_kit(kit, **kw)
        g = gs.as_dicts()[0]
        tags = synth.full_tags(g)
        rng = np.random.default_rng(31)
        reads = []
        for rd in range(8):                                          # (the oracle's batch driver holds 64 rows per read)
            body = rng.choice(np.frombuffer(b"ACGT", np.uint8), 400_000 + 17 * rd).copy()
            for t, base in enumerate(range(300 * 256, 390_000, 300 * 256)):
                for j, delta in enumerate(range(-150, 151, 50)):
                    tag = tags[(7 * t + j + rd) % len(tags)]
                    tag = synth.mutate(rng, tag, 0.04)
                    if (t + j + rd) & 1:
                        tag = synth.revcomp(tag)
                    pos = base + delta * 7 + rd * 3 - len(tag) // 2
                    body[pos:pos + len(tag)] = tag
            for c in range(5, 20):                                   # ... and across plain chunk boundaries
                tag = synth.mutate(rng, tags[(c + rd) % len(tags)], 0.05)
                if c & 1:
                    tag = synth.revcomp(tag)
                pos = c * 300 * 3 + (c * 37) % 300 - 40
                body[pos:pos + len(tag)] = tag
            reads.append(body)
        bases = np.concatenate(reads)
        offsets = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
        rows = _check(gs, bases, offsets)
        assert (rows["match_type"] < 2).sum() > 100


def test_barcode_regions_with_ambiguity_codes_and_many_insertions():
    """The barcode stage's general variant (base sets other than A/C/G/T/N resolved by OR-ing masks) and the replayed forward
    Lodhi recurrence (paths too long for the exact reversed accumulation): tags whose barcode part carries IUPAC codes,
    lower case, non-IUPAC bytes and runs of inserted bases."""
    for kit, kw in (("SQK-NBD114-96", {}), ("SQK-RBK114-96", dict(max_flank_errors=5))):
        gs = bb.GroupSet.from_kit(kit, **kw)
        g = gs.as_dicts()[0]
        tags = synth.full_tags(g)
        b0, b1 = g["bar_region"]
        rng = np.random.default_rng(77)
        junk = np.frombuffer(b"RYKMSWBDHVNacgtn-*x", np.uint8)
        reads = []
        for r in range(600):
            tag = tags[r % len(tags)].copy()
            bar = tag[b0:b1 + 1].copy()
            kind = r % 4
            if kind in (0, 2):                                   # ambiguity codes / garbage inside the barcode
                idx = rng.integers(0, len(bar), rng.integers(1, 5))
                bar[idx] = rng.choice(junk, len(idx))
            if kind in (1, 2):                                   # 6..14 inserted bases inside the barcode: long alignment paths
                ins = rng.choice(np.frombuffer(b"ACGT", np.uint8), rng.integers(6, 15))
                cut = rng.integers(3, len(bar) - 3)
                bar = np.concatenate([bar[:cut], ins, bar[cut:]])
            tag = np.concatenate([tag[:b0], bar, tag[b1 + 1:]])
            if r & 1:
                tag = synth.revcomp(tag)
            body = rng.choice(np.frombuffer(b"ACGT", np.uint8), rng.integers(100, 400))
            reads.append(np.concatenate([tag, body]) if (r >> 1) & 1 else np.concatenate([body, tag, body[:50]]))
        bases = np.concatenate(reads)
        offsets = np.concatenate([[0], np.cumsum([len(x) for x in reads])]).astype(np.uint64)
        rows = _check(gs, bases, offsets)
        assert len(rows) >= 250


def test_adversarial_text():
    """poly-N (matches everything), homopolymers, lower case, non-IUPAC bytes, tags cut by the read ends."""
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    tag = synth.full_tags(gs.as_dicts()[0])[7].tobytes()
    rnd = np.random.default_rng(3)
    body = bytes(rnd.choice(np.frombuffer(b"ACGT", np.uint8), 500))
    reads = [b"N" * 200, b"A" * 300, b"T" * 97, tag, tag[5:], tag[:-9], tag.lower() + body.lower(), body + synth.revcomp(np.frombuffer(tag, np.uint8)).tobytes()[:30],
             b"ACGT-*.@" * 20 + tag, tag[20:] + body + tag[:25], b"N" * 30 + tag[30:] + body, b"ATTGCTAAGGTTAA" * 10, b"CAGCACCT" * 20,
             tag + b"N" * 10 + tag + tag, b"G", b"", tag[:23], b"NNNNACGT" * 40]
    bases = np.frombuffer(b"".join(reads), np.uint8)
    offsets = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    _check(gs, bases, offsets)


def test_rbk_auto_threshold_two_word_patterns():
    gs = bb.GroupSet.from_kit("SQK-RBK114-96")          # flank 90 chars, k = 20
    b, o, _ = synth.make_reads(gs.as_dicts(), 400, (200, 3000), seed=25)
    _check(gs, b, o)


def test_custom_dual_end_384_panel():
    """configs[3] style: 384 random 24-mers in the ald_left / ald_right flanks, Ftag + Rtag."""
    rnd = np.random.default_rng(26)
    gl, gr = bb.GroupSet.from_fasta([cases.GOLD + "/ald_left.fasta", cases.GOLD + "/ald_right.fasta"], [0, 1]).as_dicts()
    def panel(g, n):
        pre, suf = g["flank"][:g["bar_region"][0]], g["flank"][g["bar_region"][1] + 1:]
        seqs = []
        for i in range(n):
            core = bytes(rnd.choice(np.frombuffer(b"ACGT", np.uint8), 24))
            core = b"ACGT"[i % 4:i % 4 + 1] + core[1:-1] + b"ACGT"[(i // 4) % 4:(i // 4) % 4 + 1]
            seqs.append(pre + core + suf)
        return seqs
    gs = bb.GroupSet.from_seqs([(panel(gl, 384), [f"L{i}" for i in range(384)], api.FTAG),
                                (panel(gr, 384), [f"R{i}" for i in range(384)], api.RTAG)])
    b, o, _ = synth.make_reads(gs.as_dicts(), 150, (400, 2500), seed=27)
    rows = _check(gs, b, o)
    assert (rows["match_type"] < 2).sum() > 50


def test_every_kit_preset():
    """One kit name per preset of the reference's table (kits.rs:635-708: 21 presets, flanks of 44..90 rows, k = 3..20,
    4..96 barcodes, one or two templates): rows and flank hits equal the oracle."""
    import json
    import os
    names = json.load(open(os.path.join(os.path.dirname(bb.lib_path()), "data", "kits.json")))["kit_names"]
    first = {}
    for kit, preset in names:
        first.setdefault(preset, kit)
    assert len(first) >= 20
    for preset, kit in sorted(first.items()):
        gs = bb.GroupSet.from_kit(kit)
        b, o, _ = synth.make_reads(gs.as_dicts(), 150, (200, 2500), seed=100 + len(preset))
        _check(gs, b, o)


def test_thresholds_and_alpha_variants():
    gs = bb.GroupSet.from_kit("SQK-NBD114-96", max_flank_errors=8)
    b, o, _ = synth.make_reads(gs.as_dicts(), 300, (100, 1500), seed=28, p_mut=0.12)
    _check(gs, b, o, alpha=0.5, min_score=0.5, min_score_diff=0.3)
    _check(gs, b, o, alpha=0.4, min_score=0.0, min_score_diff=0.0)


def test_device_api_and_pipeline_equal_host_api():
    import torch
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    b, o, _ = synth.make_reads(gs.as_dicts(), 2000, (500, 3000), seed=29)
    an = _annotator(gs)
    want = an.annotate(b, o)
    # device-resident inputs
    tb = torch.from_numpy(b).cuda()
    to = torch.from_numpy(o.astype(np.int64)).cuda()
    n = an.annotate_device(tb.data_ptr(), to.data_ptr(), len(o) - 1, len(b), torch.cuda.current_stream().cuda_stream)
    assert an.fetch_rows(n).tobytes() == want.tobytes()
    # pipelined submit/collect over 5 sub-batches on two streams
    cuts = np.linspace(0, len(o) - 1, 6).astype(int)
    parts = []
    subs = []
    for i in range(5):
        lo, hi = cuts[i], cuts[i + 1]
        sb = np.ascontiguousarray(b[int(o[lo]):int(o[hi])]); so = (o[lo:hi + 1] - o[lo]).astype(np.uint64)
        subs.append((sb, so, lo))
    inflight = 0
    nxt = 0
    while nxt < 5 or inflight:
        while nxt < 5 and inflight < 2:
            sb, so, lo = subs[nxt]
            an.submit(sb.ctypes.data, so.ctypes.data, len(so) - 1, tag=nxt)
            nxt += 1; inflight += 1
        tag, rows = an.collect()
        rows["read_idx"] += np.uint32(subs[tag][2])
        parts.append((tag, rows)); inflight -= 1
    assert [t for t, _ in parts] == [0, 1, 2, 3, 4]
    assert np.concatenate([r for _, r in parts]).tobytes() == want.tobytes()
    c = an.counters()
    assert c["total"] == 3 * (len(o) - 1) and c["kept"] + c["dropped"] == c["total"]
    assert c["kept"] == 3 * len(np.unique(want["read_idx"]))
    assert an.kernel_launches() > 0
    an.close()


def test_full_size_properties():
    """At bench size the oracle is too slow; check size-independent properties instead: the result of a batch equals the
    concatenation of the results of its halves (reads are independent), and a sample of reads agrees with the oracle."""
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    n = 20000
    b, o, _ = synth.make_reads(gs.as_dicts(), n, 10000, seed=30)
    an = _annotator(gs)
    whole = an.annotate(b, o)
    h = n // 2
    a1 = an.annotate(b[:int(o[h])], o[:h + 1])
    a2 = an.annotate(b[int(o[h]):], (o[h:] - o[h]).astype(np.uint64))
    a2["read_idx"] += np.uint32(h)
    assert np.concatenate([a1, a2]).tobytes() == whole.tobytes()
    assert (np.diff(whole["read_idx"].astype(np.int64)) >= 0).all()            # grouped by read, input order
    idx = np.arange(0, n, 97)
    sb = np.concatenate([b[int(o[i]):int(o[i + 1])] for i in idx])
    so = np.concatenate([[0], np.cumsum([int(o[i + 1] - o[i]) for i in idx])]).astype(np.uint64)
    want = O.demux_batch(gs.as_dicts(), sb, so)
    sel = whole[np.isin(whole["read_idx"], idx)].copy()
    remap = {int(v): k for k, v in enumerate(idx)}
    sel["read_idx"] = [remap[int(v)] for v in sel["read_idx"]]
    assert sel.tobytes() == want.tobytes()
    an.close()


def test_set_groups_rejects_unsupported_geometry():
    """What bb_set_groups still refuses: a flank threshold beyond 120, and -- checked on the host side by
    bb_groups_add -- nothing the shipped kits or the reference's example FASTAs need."""
    gs = bb.GroupSet.from_kit("SQK-NBD114-96", max_flank_errors=121)
    with pytest.raises(bb.BarbellError):
        _annotator(gs)


@pytest.mark.parametrize("kit,k", [("SQK-NBD114-96", 16), ("SQK-NBD114-96", 17), ("SQK-NBD114-96", 18), ("SQK-NBD114-96", 23), ("SQK-RBK114-96", 36)])
def test_flank_threshold_at_and_above_the_full_overhang_cost(kit, k, monkeypatch):
    """--flank-max-errors takes any usize in the reference (bin/main.rs:90-91, annotator.rs:216-229).  From k = floor(alpha * len) on
    the flank hanging over a read end completely is itself a match (cost floor(alpha * len) <= k), so every read end reports
    matches on both strands; the GPU path must degrade exactly like the oracle does -- rows and flank hits."""
    monkeypatch.setenv("ORC_PER_READ", "4096")           # such thresholds match almost everywhere: hundreds of flank matches per read
    gs = bb.GroupSet.from_kit(kit, max_flank_errors=k)
    b, o, _ = synth.make_reads(gs.as_dicts(), 40, (0, 400), seed=300 + k)
    rows = _check(gs, b, o, _cap=2048)
    assert len(rows) > 0


def test_cli_fastq_to_annotation_tsv(tmp_path):
    """`barbell annotate --kit ... -i reads.fastq(.gz) -o out.tsv` == golden annotation.tsv of configs[0]; `kit` subcommand too."""
    import gzip
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(bb.lib_path()), "barbell")
    gs, bases, offsets, rows, _ = cases.load_case("nbd_1k")
    n = len(offsets) - 1
    fq1, fq2 = tmp_path / "a.fastq", tmp_path / "b.fastq.gz"
    def rec(i):
        s = bases[int(offsets[i]):int(offsets[i + 1])].tobytes().decode()
        return f"@read_{i} runid=abc ch=7\n{s}\n+\n{'I' * len(s)}\n"
    with open(fq1, "w") as f:
        f.write("".join(rec(i) for i in range(n // 2)))
    with gzip.open(fq2, "wt") as f:
        f.write("".join(rec(i) for i in range(n // 2, n)))
    want = open(cases.GOLD + "/nbd_1k.annotation.tsv").read()
    out = tmp_path / "anno.tsv"
    r = subprocess.run([exe, "annotate", "--kit", "SQK-NBD114-96", "-i", str(fq1), str(fq2), "-o", str(out), "-t", "4", "--batch-mb", "1"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "Annotation complete!" in r.stdout and "Auto edit flank cut off: 4" in r.stdout, r.stdout + r.stderr
    assert open(out).read() == want
    outdir = tmp_path / "kit_out"
    r = subprocess.run([exe, "kit", "-k", "SQK-NBD114-96", "-i", str(fq1), str(fq2), "-o", str(outdir)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "Done!" in r.stdout, r.stdout + r.stderr
    assert open(outdir / "annotation.tsv").read() == want
    # the later stages of `kit` (use_kit.rs:50-105) against the Python restatement of filter / inspect / trim
    import copy
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(cases.GOLD), "..", "oracle"))
    import post_oracle as P
    from barbell_b200 import post
    anno = P.parse_tsv(want)
    assert open(outdir / "pattern_per_read.tsv").read() == "".join(f"{g[0].read_id}\t{P.group_structure(g, 250)}\n" for g in P.group_reads(anno))
    kept, _ = P.filter_rows(copy.deepcopy(anno), [P.parse_pattern(x) for x in post.kit_patterns("SQK-NBD114-96")])
    assert open(outdir / "filtered.tsv").read() == P.to_tsv(kept) and len(kept) > 100
    by = {}
    for row in kept:
        by.setdefault(row.read_id, []).append(row)
    files = {}
    for i in range(n):
        rid = f"read_{i}"
        if rid in by:
            s = bases[int(offsets[i]):int(offsets[i + 1])].tobytes()
            for ts, tq, lab, suf in P.process_read_and_anno(s, b"I" * len(s), by[rid], add_orientation=False, add_flank=False, only_side="left"):
                files.setdefault(lab + ".trimmed.fastq", []).append(f"@{rid}{suf} runid=abc ch=7\n{ts.decode()}\n+\n{tq.decode()}\n")
    got = {f: open(outdir / f).read() for f in os.listdir(outdir) if f.endswith(".trimmed.fastq")}
    assert got == {k: "".join(v) for k, v in files.items()} and len(got) > 20
    # unknown kit: message, exit code 0 (reference bin/main.rs:301-304), empty/no output
    r = subprocess.run([exe, "annotate", "--kit", "SQK-NOPE", "-i", str(fq1), "-o", str(tmp_path / "x.tsv")], capture_output=True, text=True)
    assert r.returncode == 0 and "Error during processing" in r.stdout


def test_prefilter_queue_overflow_falls_back_to_exact_verification():
    """Low-complexity text that keeps matching the filter's run: a periodic repeat of the flank prefix gives a candidate run every
    14 bases -- more than 2048 per CTA tile -- so those tiles are verified exactly (one window per chunk and strand); poly-N
    makes every position a candidate (one long run per chunk).  The result must still equal the oracle."""
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    b, o, _ = synth.make_reads(gs.as_dicts(), 40, (500, 3000), seed=31)
    tag = synth.full_tags(gs.as_dicts()[0])[11]
    rep = np.frombuffer(b"ATTGCTAAGGTTAA" * 12000, np.uint8).copy()      # 168 kb: two whole CTA tiles of the repeat
    rep[100_000:100_000 + len(tag)] = tag                                  # ... with a real tag inside
    junk = [np.frombuffer(b"N" * 6000, np.uint8), np.frombuffer(b"ATTGCTAAGGTTAA" * 300, np.uint8), rep,
            synth.revcomp(np.frombuffer(b"ATTGCTAAGGTTAA" * 7000, np.uint8))]
    bases = np.concatenate([b] + junk)
    offsets = np.concatenate([o, o[-1] + np.cumsum([len(x) for x in junk])]).astype(np.uint64)
    rows = _check(gs, bases, offsets)
    assert (rows["read_idx"] == 42).any()                                  # the tag inside the repeat was found


def test_long_barcodes_unpacked_history_and_odd_counts():
    """Custom panel with 40-base barcodes (padded patterns of 60 rows, regions past 48 bases: the 16-byte row records of k_barcode_rows), 37 barcodes
    (not a multiple of 32), one-sided short flanks, mixed with an exact-scan-only group."""
    rnd = np.random.default_rng(33)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    pre, suf = b"GGTCTAGACCATGCTAGGAT", b"TTGACCGATTCAGGCATCAA"
    seqs = []
    for i in range(37):
        core = bytes(rnd.choice(acgt, 40))
        core = b"ACGT"[i % 4:i % 4 + 1] + core[1:-1] + b"ACGT"[(i // 4) % 4:(i // 4) % 4 + 1]
        seqs.append(pre + core + suf)
    short = [b"ACGTTGCA" + bytes(rnd.choice(acgt, 12)) + b"GT" for _ in range(5)]
    short = [s[:8] + b"ACGT"[i % 4:i % 4 + 1] + s[9:19] + b"ACGT"[(i + 1) % 4:(i + 1) % 4 + 1] + s[20:] for i, s in enumerate(short)]
    gs = bb.GroupSet.from_seqs([(seqs, [f"X{i}" for i in range(37)], api.FTAG), (short, [f"S{i}" for i in range(5)], api.RTAG)])
    G = gs.as_dicts()
    assert G[0]["bar_len"] == 60 and len(G[0]["barcodes"]) == 37
    b, o, _ = synth.make_reads(G, 300, (150, 1500), seed=34)
    rows = _check(gs, b, o)
    assert (rows["match_type"] < 2).sum() > 50


def test_nibble_packed_host_to_device_copy_is_lossless():
    """bb_opts.flags bit 1: bases are packed two per byte on the host and expanded on the device; rows must not change
    (lower case, IUPAC codes, non-IUPAC bytes included)."""
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    b, o, _ = synth.make_reads(gs.as_dicts(), 3000, (200, 3000), seed=41, n_frac=0.01)
    b = b.copy()
    rnd = np.random.default_rng(5)
    idx = rnd.integers(0, len(b), 3000)
    b[idx] = rnd.choice(np.frombuffer(b"acgtnRYKMSWBDHVryU-*.@x", np.uint8), len(idx))
    want = O.demux_batch(gs.as_dicts(), b, o)
    an = _annotator(gs, pack_h2d=True)
    got = an.annotate(b, o)
    # pipelined calls take the same path
    an.submit(b.ctypes.data, o.ctypes.data, len(o) - 1, tag=7)
    tag, got2 = an.collect()
    # the head/tail split adapts from batch to batch (measured pack and link rates): every split must give the same rows
    for _ in range(4):
        assert an.annotate(b, o).tobytes() == want.tobytes()
    moved = an.h2d_bytes()
    an.close()
    assert got.tobytes() == want.tobytes() and got2.tobytes() == want.tobytes() and tag == 7
    assert 6 * (len(b) // 2) < moved < 6 * (len(b) + 8 * len(o))      # fewer bytes than six plain copies


@pytest.mark.gpu
def test_crumb_packed_host_to_device_copy_is_lossless():
    """bb_opts.flags bit 2: 2 bits per base for A/C/G/T plus an exception list for every other byte (N, IUPAC codes, lower case
    is folded, junk bytes); rows must not change.  An N-rich batch overflows the exception list and the context falls back to
    the nibble format -- same rows again."""
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    b, o, _ = synth.make_reads(gs.as_dicts(), 3000, (200, 3000), seed=43, n_frac=0.002)
    b = b.copy()
    rnd = np.random.default_rng(6)
    idx = rnd.integers(0, len(b), 3000)
    b[idx] = rnd.choice(np.frombuffer(b"acgtnRYKMSWBDHVryU-*.@x", np.uint8), len(idx))
    want = O.demux_batch(gs.as_dicts(), b, o)
    an = _annotator(gs, pack_h2d="crumbs")
    got = an.annotate(b, o)
    an.submit(b.ctypes.data, o.ctypes.data, len(o) - 1, tag=9)
    tag, got2 = an.collect()
    for _ in range(4):                                   # the head/tail split moves from batch to batch
        assert an.annotate(b, o).tobytes() == want.tobytes()
    moved = an.h2d_bytes()
    assert got.tobytes() == want.tobytes() and got2.tobytes() == want.tobytes() and tag == 9
    assert 6 * (len(b) // 4) < moved < 6 * (len(b) + 8 * len(o))
    # 6 % N: too many exceptions -> nibble format from here on, still lossless
    b2, o2, _ = synth.make_reads(gs.as_dicts(), 3000, (200, 3000), seed=44, n_frac=0.06)
    want2 = O.demux_batch(gs.as_dicts(), b2, o2)
    for _ in range(3):
        assert an.annotate(b2, o2).tobytes() == want2.tobytes()
    assert an.annotate(b, o).tobytes() == want.tobytes()
    an.close()


POLICY_SETS = [1, 2, 4, 8, 16, 32, 1 | 2 | 4 | 8, 2 | 4 | 16, 1 | 8 | 32, 63 - 32]


@pytest.mark.parametrize("pol", POLICY_SETS)
def test_search_policies_gpu_equals_oracle(pol):
    """The choices of sassy that the reference's tests do not pin (S1 plateau side, S2 traceback order, S3 overhang rounding,
    S5 tie between equal minima, S6 strand order) are run-time switches of BOTH the oracle (orc_policy.flags) and the product
    (bb_opts.policy): under every setting the GPU rows and flank hits must equal the oracle's."""
    for kit, kw, n in (("SQK-NBD114-96", {}, 400), ("SQK-RBK114-96", dict(max_flank_errors=5), 200), ("SQK-RBK114-96", {}, 120)):
        gs = bb.GroupSet.from_kit(kit, **kw)
        b, o, _ = synth.make_reads(gs.as_dicts(), n, (150, 2500), seed=900 + pol)
        with O.policy(pol):
            rows = _check(gs, b, o, policy=pol)
        base = _annotator(gs)
        try:
            default_rows = base.annotate(b, o)
        finally:
            base.close()
        if pol in (1, 2, 1 | 2 | 4 | 8) and kit == "SQK-NBD114-96":
            assert rows.tobytes() != default_rows.tobytes(), "the policy did not change a single row: the knob is not wired"


def test_global_window_queue_overflow_reruns_the_batch_exactly(monkeypatch):
    """A window queue too small for the batch (forced with BB_WIN_CAP): CTAs that cannot reserve slots leave holes, so the
    verify kernel must not decode the queue at all and the engine re-runs the batch with the exact scan -- same rows."""
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    b, o, _ = synth.make_reads(gs.as_dicts(), 3000, (400, 3000), seed=4711)
    an = _annotator(gs)
    want = an.annotate(b, o)
    an.close()
    monkeypatch.setenv("BB_WIN_CAP", "7")
    an = _annotator(gs)
    try:
        for _ in range(3):                              # stale queue contents of an earlier batch must not matter either
            got = an.annotate(b, o)
            assert got.tobytes() == want.tobytes()
    finally:
        an.close()
    assert len(want) > 2000


def test_submit_packed_equals_submit():
    """bb_submit_packed: the batch arrives as the 2-bit stream + exception list built read by read (bb_pack_crumbs_append, as the
    CLI's FASTQ reader does); rows must equal those of the byte form, and the library must move about a quarter of the bytes."""
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    b, o, _ = synth.make_reads(gs.as_dicts(), 2500, (0, 3000), seed=45, n_frac=0.002)
    b = b.copy()
    rnd = np.random.default_rng(8)
    idx = rnd.integers(0, len(b), 2000)
    b[idx] = rnd.choice(np.frombuffer(b"acgtnRYKMryU-*x", np.uint8), len(idx))
    want = O.demux_batch(gs.as_dicts(), b, o)
    n = len(b)
    crumbs = np.zeros((n + 3) // 4 + 64, np.uint8); exc = np.zeros(n // 16 + 64, np.uint64)
    pos, ne = C.c_uint64(0), C.c_uint64(0)
    for r in range(len(o) - 1):
        lo, hi = int(o[r]), int(o[r + 1])
        assert bb.lib().bb_pack_crumbs_append(b[lo:].ctypes.data if hi > lo else None, hi - lo, crumbs.ctypes.data, C.byref(pos), exc.ctypes.data,
                                              len(exc), C.byref(ne)) == 0
    assert pos.value == n
    offs = np.ascontiguousarray(o, dtype=np.uint64)
    an = _annotator(gs)
    try:
        h0 = an.h2d_bytes()
        an.submit_packed(crumbs.ctypes.data, n, exc.ctypes.data, ne.value, offs.ctypes.data, len(o) - 1, tag=5)
        tag, got = an.collect()
        assert tag == 5 and got.tobytes() == want.tobytes()
        assert an.h2d_bytes() - h0 < 0.3 * n + 8 * len(o) + 64
        an.submit_packed(crumbs.ctypes.data, 0, None, 0, np.zeros(1, np.uint64).ctypes.data, 0, tag=6)      # empty batch
        tag, got = an.collect()
        assert tag == 6 and len(got) == 0
    finally:
        an.close()


def test_slot_path_overflows_fall_back_to_the_sorted_path(monkeypatch):
    """The default engine keeps every read's sub-threshold entries in per-read slots and resolves them with one warp per read (no
    radix sort, no host round trip).  Too few slots for a read, or too small a hit list, re-run the batch on the global-sort
    path; BB_GLUE=sort selects that path outright.  Same rows and flank hits in every case, repeat-rich reads included."""
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    g = gs.as_dicts()[0]
    b, o, _ = synth.make_reads(gs.as_dicts(), 1500, (300, 4000), seed=77)
    # a few reads made of many tags in a row: dozens of sub-threshold positions per read
    tags = synth.full_tags(g)
    rng = np.random.default_rng(78)
    extra = []
    for r in range(6):
        extra.append(np.concatenate([synth.mutate(rng, tags[(7 * r + q) % len(tags)], 0.03) for q in range(25)] + [rng.choice(np.frombuffer(b"ACGT", np.uint8), 500)]))
    bases = np.concatenate([b] + extra)
    offsets = np.concatenate([o, o[-1] + np.cumsum([len(x) for x in extra]).astype(np.uint64)])
    want = O.demux_batch(gs.as_dicts(), bases, offsets, cap_per_read=256)
    want_h = O.flank_hits_batch(gs.as_dicts(), bases, offsets, cap_per_read=256)

    def run():
        an = _annotator(gs)
        try:
            rows = an.annotate(bases, offsets)
            return rows, an.flank_hits()
        finally:
            an.close()
    rows, hits = run()
    assert rows.tobytes() == want.tobytes() and (hits == want_h).all()
    for key, val in (("BB_SLOT_CAP", "8"), ("BB_SLOT_CAP", "120"), ("BB_HITS_CAP", "100"), ("BB_GLUE", "sort")):
        monkeypatch.setenv(key, val)
        rows2, hits2 = run()
        monkeypatch.delenv(key)
        assert rows2.tobytes() == want.tobytes() and (hits2 == want_h).all(), key


def test_cli_in_process_multi_gpu(tmp_path):
    """`barbell annotate --gpus 2`: one process, batches dealt round-robin to one context per GPU, rows written in input order --
    byte-identical to the single-GPU output (needs two visible GPUs: gpurun --gpus 2)."""
    import os
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    exe = os.path.join(os.path.dirname(bb.lib_path()), "barbell")
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    b, o, _ = synth.make_reads(gs.as_dicts(), 30000, (500, 6000), seed=91)
    fq = tmp_path / "r.fastq"
    synth.write_fastq(str(fq), b, o)
    outs = []
    for gpus in (1, 2):
        out = tmp_path / f"a{gpus}.tsv"
        r = subprocess.run([exe, "annotate", "--kit", "SQK-NBD114-96", "-i", str(fq), "-o", str(out), "-t", "8", "--batch-mb", "8", "--gpus", str(gpus)],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "Annotation complete!" in r.stdout, r.stdout + r.stderr
        outs.append(open(out).read())
    assert outs[0] == outs[1] and outs[0].count("\n") > 20000
    ids = [f"read_{i}" for i in range(len(o) - 1)]
    an = _annotator(gs)
    try:
        assert bb.rows_to_tsv(an.annotate(b, o), gs, ids) == outs[0]
    finally:
        an.close()


def test_barcodes_with_ambiguity_codes_in_the_patterns():
    """Custom query sets may carry IUPAC codes inside the barcodes themselves (README custom-primer example): the barcode stage's
    text masks are indexed by the pattern character's 4-bit base set, so R / Y / N rows match two / two / four bases.  Also a
    panel whose barcodes share their first bases (the shared leading rows then reach into the barcode)."""
    rnd = np.random.default_rng(123)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    pre, suf = b"GGTCTAGACCATGCTAGGAT", b"TTGACCGATTCAGGCATCAA"
    seqs = []
    for i in range(40):
        core = bytearray(bytes(rnd.choice(acgt, 24)))
        core[0:3] = b"ACG"                                            # common first barcode bases: more shared rows
        core[3] = b"ACGT"[i % 4]
        core[-1] = b"ACGT"[(i // 4) % 4]
        for pos, ch in ((7, b"R"), (12, b"Y"), (15, b"N"), (18, b"K")):
            if (i + pos) % 3 == 0:
                core[pos] = ch[0]
        seqs.append(pre + bytes(core) + suf)
    gs = bb.GroupSet.from_seqs([(seqs, [f"I{i}" for i in range(40)], 0)])
    assert gs.as_dicts()[0]["bar_region"][0] == len(pre) + 3          # the common barcode prefix moved into the flank (barcodes.rs LCP)
    b, o, _ = synth.make_reads(gs.as_dicts(), 800, (200, 2500), seed=124)
    rows = _check(gs, b, o)
    assert (rows["match_type"] < 2).sum() > 200


@pytest.mark.gpu
def test_reserve_changes_nothing_but_the_first_batch_cost():
    """bb_reserve (every engine sized and warmed by an all-'A' batch of the given shape) before the first submit: rows and counters
    of the batches that follow are those of a context that was not reserved; it is refused while batches are in flight."""
    import torch
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    b, o, _ = synth.make_reads(gs.as_dicts(), 3000, 3000, seed=77)
    want = O.demux_batch(gs.as_dicts(), b, o)
    hb = torch.from_numpy(b).pin_memory(); ho = torch.from_numpy(o.astype(np.uint64).view(np.int64)).pin_memory()
    a0 = bb.Annotator(gs); a1 = bb.Annotator(gs)
    a1.reserve(4000, 4000 * 3000)
    assert a1.kernel_launches() == 0 and a1.counters() == dict(total=0, kept=0, dropped=0)
    for an in (a0, a1):
        for rep in range(5):                                      # past BB_MAX_INFLIGHT: every engine runs at least once
            an.submit(hb.data_ptr(), ho.data_ptr(), len(o) - 1, tag=rep)
            tag, rows = an.collect()
            assert tag == rep and rows.tobytes() == want.tobytes()
    assert a1.kernel_launches() <= a0.kernel_launches() and a0.counters() == a1.counters()
    a1.submit(hb.data_ptr(), ho.data_ptr(), len(o) - 1, tag=9)
    with pytest.raises(bb.BarbellError, match="in flight"):
        a1.reserve(10, 1000)
    a1.collect()
    a0.close(); a1.close()


@pytest.mark.gpu
def test_cli_error_paths_end_cleanly(tmp_path):
    """`barbell annotate` with its parser threads, GPU workers and writer thread: an unwritable output and a malformed record in
    the middle of the stream must end the run with the reference's message and without "Annotation complete!" (the exit code stays
    0 like the reference's, bin/main.rs:301-304), not with a hang or a truncated-but-successful file."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(bb.lib_path()), "barbell")
    gs, bases, offsets, rows, _ = cases.load_case("nbd_1k")
    n = len(offsets) - 1
    def rec(i):
        s = bases[int(offsets[i]):int(offsets[i + 1])].tobytes().decode()
        return f"@read_{i}\n{s}\n+\n{'I' * len(s)}\n"
    good = "".join(rec(i) for i in range(n))
    fq = tmp_path / "good.fastq"; fq.write_text(good * 3)
    if os.path.exists("/dev/full"):
        r = subprocess.run([exe, "annotate", "--kit", "SQK-NBD114-96", "-i", str(fq), "-o", "/dev/full", "-t", "4", "--chunk-kb", "256"],
                           capture_output=True, text=True, timeout=120)
        assert "Error during processing: write to /dev/full failed" in r.stdout and "Annotation complete!" not in r.stdout, r.stdout + r.stderr
    bad = tmp_path / "bad.fastq"; bad.write_text(good * 2 + "@broken\nACGT\n+\nII\n" + good)
    for extra in (["-t", "4", "--chunk-kb", "256"], ["--single-reader"]):
        out = tmp_path / "bad.tsv"
        r = subprocess.run([exe, "annotate", "--kit", "SQK-NBD114-96", "-i", str(bad), "-o", str(out)] + extra, capture_output=True, text=True, timeout=120)
        assert "Error during processing" in r.stdout and ("quality length" in r.stdout or "malformed" in r.stdout), r.stdout + r.stderr
        assert "Annotation complete!" not in r.stdout
    # and the good file through the same small chunks gives the golden rows three times over
    out = tmp_path / "good.tsv"
    r = subprocess.run([exe, "annotate", "--kit", "SQK-NBD114-96", "-i", str(fq), "-o", str(out), "-t", "6", "--chunk-kb", "256"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "Annotation complete!" in r.stdout, r.stdout + r.stderr
    want = open(cases.GOLD + "/nbd_1k.annotation.tsv").read().splitlines(keepends=True)
    assert open(out).read() == want[0] + "".join(want[1:]) * 3


@pytest.mark.gpu
def test_large_barcode_panel_1200_patterns():
    """A group may hold up to 4096 barcodes (the reference has no limit, barcodes.rs:106-197): 1200 patterns = 38 rounds of 32
    lanes with a partly filled last round, through the barcode stage's fallback rule and top-two reduction; a group beyond the
    limit is refused with a message."""
    rnd = np.random.default_rng(321)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    pre, suf = b"GGTCTAGACCATGCTAGGAT", b"TTGACCGATTCAGGCATCAA"
    def panel(n):
        seqs = []
        for i in range(n):
            core = bytearray(bytes(rnd.choice(acgt, 24)))
            core[0] = b"ACGT"[i % 4]; core[-1] = b"ACGT"[(i // 4) % 4]
            seqs.append(pre + bytes(core) + suf)
        return seqs
    gs = bb.GroupSet.from_seqs([(panel(1200), [f"P{i}" for i in range(1200)], 0)])
    b, o, _ = synth.make_reads(gs.as_dicts(), 600, (200, 2500), seed=322)
    rows = _check(gs, b, o)
    tags = rows[rows["match_type"] < 2]
    assert len(tags) > 150 and tags["label_idx"].max() >= 1024 and len(np.unique(tags["label_idx"])) > 100
    big = bb.GroupSet.from_seqs([(panel(4100), [f"Q{i}" for i in range(4100)], 0)])
    with pytest.raises(bb.BarbellError, match="barcodes per group"):
        bb.Annotator(big)
