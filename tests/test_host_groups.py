"""Host-side pattern sets (C++ behind the C ABI) against the reference's own unit vectors:
src/annotate/barcodes.rs:443-555 and src/kits/kits.rs:1105-1183, plus an independent python restatement."""
import os

import pytest

import barbell_b200 as bb
from barbell_b200 import api

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- kits.rs tests ----
def test_get_barcodes_bc_1_to_12():
    assert api.label_range("BC01", "BC12") == [f"BC{i:02d}" for i in range(1, 13)]


def test_get_barcodes_12a_from_suffix():
    assert api.label_range("BC1A", "BC12A") == [f"BC{i:02d}" for i in range(1, 12)] + ["BC12A"]
    assert api.label_range("BC1A", "BC13A") == [f"BC{i:02d}" for i in range(1, 12)] + ["BC12A", "BC13"]


def test_get_barcodes_nb():
    assert api.label_range("NB01", "NB12") == [f"NB{i:02d}" for i in range(1, 13)]


def test_get_barcodes_rbk_special_relabel():
    assert api.label_range("RBK24", "RBK28") == ["BC24", "BC25", "RBK26", "BC27", "BC28"]


def test_get_barcodes_12a_flag():
    r = api.label_range("BC01", "BC24", True)
    assert len(r) == 24 and r[11] == "BC12A" and r[10] == "BC11" and r[12] == "BC13"
    assert api.label_range("BC01", "BC12", True) == [f"BC{i:02d}" for i in range(1, 12)] + ["BC12A"]


def test_lookup_barcode_seq():
    assert api.lookup_barcode_seq("BC12A") == "GTTGAGTTACAAAGCACCGATCAG"
    assert api.lookup_barcode_seq("BC01") == "AAGAAAGTTGTCGGTGTCTTTGTG"
    assert api.lookup_barcode_seq("NB01") == "CACAAAGACACCGACAACTTTCTT"      # reverse complement of BC01
    assert api.lookup_barcode_seq("RBK26") == "ACTATGCCTTTCCGTGAAACAGTT"
    assert api.lookup_barcode_seq("ZZ01") is None


# ---- barcodes.rs tests ----
def test_barcode_group_small():
    g = bb.GroupSet.from_seqs([([b"AAATTTGGG", b"AAACCCGGG"], ["s1", "s2"], api.FTAG)]).as_dicts()[0]
    assert g["flank"] == b"AAANNNGGG" and g["bar_region"] == (3, 5)
    assert g["barcodes"] == [b"AAATTTGGG", b"AAACCCGGG"]       # padding saturates
    assert g["pad_region"] == (0, 16)                          # pad end is NOT clamped (barcodes.rs:160-163)


def test_barcode_group_errors():
    with pytest.raises(bb.BarbellError):
        bb.GroupSet.from_seqs([([b"@@@@@@@@@", b"AAACCCGGG"], ["a", "b"], api.FTAG)])
    with pytest.raises(bb.BarbellError):
        bb.GroupSet.from_seqs([([b"AAATTTGGG", b"AAAAAAACCCGGG"], ["a", "b"], api.FTAG)])
    with pytest.raises(bb.BarbellError):
        bb.GroupSet.from_seqs([([b"AAATTTGGG"], ["a"], api.FTAG)])                      # single query panics upstream
    with pytest.raises(bb.BarbellError):
        bb.GroupSet.from_seqs([([b"AAATTTGGG", b"CCCGGGTTT"], ["a", "b"], api.FTAG)])  # no anchors
    with pytest.raises(bb.BarbellError):
        bb.GroupSet.from_kit("SQK-NOPE-1")


def test_fasta_read_rapid():
    g = bb.GroupSet.from_fasta([os.path.join(GOLD, "rapid_bars.fasta")], [api.FTAG]).as_dicts()[0]
    assert g["flank"] == (b"GCTTGGGTGTTTAACC" + b"N" * 24 + b"GTTTTCGCATTTATCGTGAAACGCTTTCGCGTTTTTCGTGCGCCGCTTCA")
    assert g["bar_region"] == (16, 39) and len(g["barcodes"]) == 96
    assert g["barcodes"][0][10:34] == b"AAGAAAGTTGTCGGTGTCTTTGTG"
    assert g["k_flank"] == 20                                  # paper App. C: kappa(66) = 20


def test_kits_geometry():
    """SURVEY.md section 8 sizes: NBD 14+24N+8 k=4 pad 42; RBK 16+24N+50 k=20 pad 44; ALD mask 23 k=31/30."""
    g = bb.GroupSet.from_kit("SQK-NBD114-96").as_dicts()
    assert len(g) == 1 and g[0]["flank"] == b"ATTGCTAAGGTTAA" + b"N" * 24 + b"CAGCACCT"
    assert g[0]["k_flank"] == 4 and g[0]["bar_len"] == 42 and g[0]["pad_region"] == (4, 48) and g[0]["labels"][0] == "NB01"
    assert len(bb.GroupSet.from_kit("SQK-NBD114-96", use_extended=True).as_dicts()) == 1   # no Extended template for NB96
    assert len(bb.GroupSet.from_kit("SQK-RBK114-96").as_dicts()) == 1
    g = bb.GroupSet.from_kit("SQK-RBK114-96", use_extended=True).as_dicts()
    assert len(g) == 2 and g[1]["flank"].startswith(b"TTCGTGCGCCGCTTCA") and g[0]["labels"][25] == "RBK26"
    assert bb.GroupSet.from_kit("SQK-RBK114.96").as_dicts()[0]["flank"] == g[0]["flank"]   # '.' retried as '-'
    g = bb.GroupSet.from_kit("SQK-RBK114-96", max_flank_errors=5).as_dicts()
    assert g[0]["k_flank"] == 5
    g = bb.GroupSet.from_fasta([os.path.join(GOLD, "ald_left.fasta"), os.path.join(GOLD, "ald_right.fasta")],
                               [api.FTAG, api.RTAG]).as_dicts()
    assert [x["bar_region"][1] - x["bar_region"][0] + 1 for x in g] == [23, 24]
    assert [x["k_flank"] for x in g] == [31, 30] and [x["match_type"] for x in g] == [0, 1]


def test_every_supported_kit_builds():
    import json
    data = json.load(open(os.path.join(os.path.dirname(api.__file__), "data", "kits.json")))
    names = [k for k, _ in data["kit_names"]]
    assert len(names) == 39
    for k in names:
        gs = bb.GroupSet.from_kit(k, use_extended=True)
        assert len(gs) >= 1


def _py_group(seqs):
    """independent restatement of BarcodeGroup::new geometry"""
    n = len(seqs[0])
    pre = min(next((i for i in range(n) if s[i] != seqs[0][i]), n) for s in seqs)
    suf = min(next((i for i in range(n) if s[n - 1 - i] != seqs[0][n - 1 - i]), n) for s in seqs)
    mask = n - pre - suf
    pad0, pad1 = max(0, pre - 10), pre + mask + 10
    return dict(flank=seqs[0][:pre] + b"N" * mask + seqs[0][n - suf:], bar_region=(pre, pre + mask - 1), pad_region=(pad0, pad1),
                barcodes=[s[pad0:min(pad1, n)] for s in seqs])


def test_group_geometry_matches_python_restatement():
    import random
    rnd = random.Random(5)
    for _ in range(30):
        pre = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(0, 30)))
        suf = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(0 if pre else 1, 30)))
        seqs = [pre + bytes(rnd.choice(b"ACGT") for _ in range(12)) + suf for _ in range(rnd.randint(2, 6))]
        want = _py_group(seqs)
        got = bb.GroupSet.from_seqs([(seqs, [f"q{i}" for i in range(len(seqs))], api.FTAG)]).as_dicts()[0]
        for k in want:
            assert got[k] == want[k], k


def test_python_group_oracle_equals_the_product_on_every_kit():
    """oracle/groups_oracle.py (what bench.py's reference arm builds its query groups with, so that arm never loads the product
    library) restates barcodes.rs:106-197 / kits.rs independently of csrc/host/groups.cpp: both must agree on every preset."""
    import json, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle"))
    import groups_oracle as GO
    names = [k for k, _ in json.load(open(os.path.join(root, "barbell_b200", "data", "kits.json")))["kit_names"]]
    assert len(names) == 39
    for kit in names:
        for ext in (False, True):
            for mfe in (None, 3):
                a = GO.groups_from_kit(kit, ext, mfe)
                b = bb.GroupSet.from_kit(kit, ext, mfe).as_dicts()
                assert len(a) == len(b)
                for x, y in zip(a, b):
                    for k in y:
                        assert x[k] == y[k], (kit, ext, mfe, k)
    gold = os.path.join(root, "tests", "golden")
    a = GO.groups_from_fasta([gold + "/ald_left.fasta", gold + "/ald_right.fasta"], [0, 1])
    b = bb.GroupSet.from_fasta([gold + "/ald_left.fasta", gold + "/ald_right.fasta"], [0, 1]).as_dicts()
    for x, y in zip(a, b):
        for k in y:
            assert x[k] == y[k], k


def test_bench_reference_arm_does_not_load_the_product_library():
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-reads", "40"],
                       capture_output=True, text=True, check=True)
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["product_library_loaded"] is False and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
