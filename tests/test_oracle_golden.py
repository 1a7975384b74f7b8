"""The oracle reproduces the committed golden fixtures (tests/golden/, made by tools/make_golden.py)."""
import os

import numpy as np
import pytest

import barbell_b200 as bb
import cases
import oracle_lib as O


@pytest.mark.parametrize("name", sorted(cases.META["cases"]))
def test_oracle_rows_match_golden(name):
    gs, bases, offsets, rows, hits = cases.load_case(name)
    got = O.demux_batch(gs.as_dicts(), bases, offsets, n_threads=4)
    assert got.tobytes() == rows.tobytes()
    got_h = O.flank_hits_batch(gs.as_dicts(), bases, offsets, n_threads=4)
    assert (got_h == hits).all()


def test_naive_backend_reproduces_golden_config1():
    gs, bases, offsets, rows, _ = cases.load_case("nbd_1k")
    O.set_policy(0)
    try:
        got = O.demux_batch(gs.as_dicts(), bases[:int(offsets[200])], offsets[:201], n_threads=4)
    finally:
        O.set_policy(1)
    want = rows[rows["read_idx"] < 200]
    assert got.tobytes() == want.tobytes()


def test_search_vectors():
    for v in cases.META["search_vectors"]:
        ms = O.search(v["pattern"].encode(), v["text"].encode(), v["k"], alpha=v["alpha"])
        got = [dict(ts=m.text_start, te=m.text_end, ps=m.pattern_start, pe=m.pattern_end, cost=m.cost, strand=m.strand,
                    cigar=m.cigar()) for m in ms]
        assert got == v["matches"]


def test_tsv_text_golden():
    """annotation.tsv columns and formatting (reference src/annotate/searcher.rs:31-64, annotator.rs:13-26)."""
    gs, bases, offsets, rows, _ = cases.load_case("nbd_1k")
    ids = [f"read_{i}" for i in range(len(offsets) - 1)]
    text = bb.rows_to_tsv(rows, gs, ids)
    assert text == open(os.path.join(cases.GOLD, "nbd_1k.annotation.tsv")).read()
    assert text.splitlines()[0].split("\t") == ["read_id", "read_len", "rel_dist_to_end", "read_start_bar", "read_end_bar",
                                                "read_start_flank", "read_end_flank", "bar_start", "bar_end", "match_type",
                                                "flank_cost", "barcode_cost", "label", "strand", "cuts"]
    assert bb.rows_to_tsv(rows[:0], gs, ids) == ""          # zero hits -> empty file (annotator.rs:20-24)


def test_demux_recovers_implanted_barcodes():
    """Sanity of the whole restatement: group-II reads (one clean tag at the 5' end) get their barcode back."""
    gs, bases, offsets, rows, _ = cases.load_case("nbd_1k")
    from barbell_b200 import synth
    case = cases.META["cases"]["nbd_1k"]
    _, _, truth = synth.make_reads(gs.as_dicts(), case["n"], tuple(case["read_len"]), seed=case["seed"])
    by = {}
    for r in rows:
        by.setdefault(int(r["read_idx"]), []).append(r)
    ok = tot = 0
    for i, (kind, placed) in enumerate(truth):
        if kind == 1:
            tot += 1
            ok += any(int(x["label_idx"]) == placed[0][1] and x["match_type"] < 2 for x in by.get(i, []))
    assert tot > 400 and ok / tot > 0.97
