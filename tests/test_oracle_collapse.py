"""collapse_overlapping_matches: the reference's own unit tests (src/annotate/interval.rs:154-256) on the oracle."""
import random

import numpy as np

import oracle_lib as O


def row(start, end, mtype, bcost, label, fcost=0):
    r = np.zeros(1, dtype=O.ROW_DTYPE)[0]
    r["read_start_bar"], r["read_end_bar"] = start, end
    r["read_start_flank"], r["read_end_flank"] = start, end
    r["bar_end"] = 10
    r["match_type"], r["barcode_cost"], r["flank_cost"], r["label_idx"], r["read_len"] = mtype, bcost, fcost, label, 100
    return r


def collapse(rows, thr):
    arr = np.array(rows, dtype=O.ROW_DTYPE) if rows else np.zeros(0, dtype=O.ROW_DTYPE)
    n = O.lib().orc_collapse(arr.ctypes.data, len(arr), thr)
    return arr[:n]


def test_empty_and_single():
    assert len(collapse([], 0.5)) == 0
    out = collapse([row(0, 10, O.FTAG, 3, 1)], 0.5)
    assert [int(x["label_idx"]) for x in out] == [1]


def test_double_no_overlap():
    out = collapse([row(0, 10, O.FTAG, 3, 1), row(10, 20, O.FTAG, 3, 2)], 0.5)
    assert [int(x["label_idx"]) for x in out] == [1, 2]


def test_collapse_overlapping():
    out = collapse([row(0, 20, O.FTAG, 0, 1), row(15, 20, O.FTAG, 3, 2)], 0.5)
    assert [int(x["label_idx"]) for x in out] == [1]


def test_overlap_threshold():
    rows = [row(0, 20, O.FTAG, 0, 1), row(10, 35, O.FTAG, 3, 2)]
    assert [int(x["label_idx"]) for x in collapse(rows, 0.5)] == [1]
    assert [int(x["label_idx"]) for x in collapse(rows, 0.6)] == [1, 2]


def test_correct_sorting_under_shuffle():
    rows = [row(0, 10, O.FTAG, 0, 1), row(10, 20, O.FTAG, 3, 2), row(0, 15, O.FTAG, 3, 2), row(100, 110, O.FTAG, 3, 3)]
    rnd = random.Random(0)
    for _ in range(10):
        rnd.shuffle(rows)
        assert [int(x["label_idx"]) for x in collapse(list(rows), 0.5)] == [1, 3]


def test_small_overlap_boundary():
    a, b = row(0, 10, O.FTAG, 3, 1), row(10, 20, O.FTAG, 1, 2)
    for _ in range(4):
        b["read_start_flank"] -= 1
        b["read_end_flank"] -= 1
        assert [int(x["label_idx"]) for x in collapse([a, b], 0.5)] == [1, 2]
    b["read_start_flank"] -= 1
    b["read_end_flank"] -= 1
    assert [int(x["label_idx"]) for x in collapse([a, b], 0.5)] == [2]


def test_tag_beats_flank_and_longest_flank_wins():
    out = collapse([row(0, 40, O.FFLANK, 42, -1), row(2, 40, O.FTAG, 5, 7)], 0.8)
    assert int(out[0]["label_idx"]) == 7
    out = collapse([row(0, 40, O.FFLANK, 42, -1), row(0, 46, O.RFLANK, 42, -2)], 0.8)
    assert int(out[0]["label_idx"]) == -2
