"""Pattern-set construction of the reference, restated in plain Python -- ORACLE side, TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; it exists so
that those legs can build the query groups WITHOUT loading the product library (libbarbell_b200.so has its own C++ host code
for the same job, csrc/host/groups.cpp; tests/test_host_groups.py checks the two against each other on every kit).

Follows rickbeeloo/barbell @ 9a2b814:
  BarcodeGroup::new            src/annotate/barcodes.rs:106-197   (LCP / LCS flanks, N mask, padded barcodes, regions)
  BarcodeGroup::new_from_kit   src/annotate/barcodes.rs:251-299
  get_flanks / LCP / LCS       src/annotate/barcodes.rs:323-385
  get_kit_info                 src/kits/kits.rs:635-708            (names with '.' are retried with '-')
  get_barcodes                 src/kits/kits.rs:741-816
  lookup_barcode_seq           src/kits/kits.rs:1074-1103
  get_edit_cut_off             src/annotate/edit_model.rs:2-11
  annotate_with_groups         src/annotate/annotator.rs:207-231   (threshold selection)
The kit DATA (sequences, templates, kit-name table) comes from barbell_b200/data/kits.json, which tools/gen_kit_tables.py
extracted from src/kits/kits.rs.
"""
import json
import math
import os
import re

PADDING = 10                      # src/lib.rs:10
FTAG, RTAG = 0, 1
_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "barbell_b200", "data", "kits.json")
_kits = None


def _data():
    global _kits
    if _kits is None:
        _kits = json.load(open(_DATA))
    return _kits


def edit_cut_off(l: int) -> int:
    """edit_model.rs:2-11"""
    v = math.ceil(0.5100 * l - 1.7312 * math.sqrt(l))
    return int(v) if v > 0 else 0


def parse_label_simple(label: str):
    """kits.rs:710-739: alphabetic prefix, number, optional trailing 'A'"""
    a_flag = label.endswith("A") and len(label) > 1 and label[-2].isdigit()
    core = label[:-1] if a_flag else label
    m = re.match(r"^([A-Za-z]+)(\d+)$", core)
    if not m:
        raise ValueError(f"Invalid numeric part in label {label!r}")
    return m.group(1), int(m.group(2)), a_flag


def get_barcodes(from_label: str, to_label: str, use_12a_flag: bool):
    """kits.rs:741-816"""
    pf_from, from_num, from_a = parse_label_simple(from_label)
    pf_to, to_num, to_a = parse_label_simple(to_label)
    assert pf_from == pf_to, f"Mismatched label prefixes: {pf_from} vs {pf_to}"
    start, end = (from_num, to_num) if from_num <= to_num else (to_num, from_num)
    if pf_from != "AB":
        out = [f"BC{i:02d}" for i in range(start, end + 1)]
    else:
        out = [f"AB{i:02d}" for i in range(start, end + 1)]
    if use_12a_flag or ((from_a or to_a) and start <= 12 <= end):
        out = ["BC12A" if x == "BC12" else x for x in out]
    if pf_from == "NB":
        out = [x.replace("BC", "NB", 1) if x.startswith("BC") else x for x in out]
    if pf_from == "RBK":
        special = {26, 39, 40, 48, 54, 60}
        out = [x.replace("BC", "RBK", 1) if x.startswith("BC") and len(x) >= 4 and x[2:4].isdigit() and int(x[2:4]) in special else x
               for x in out]
    return out


def lookup_barcode_seq(label: str):
    """kits.rs:1074-1103"""
    d = _data()
    prefix, number, is_a = parse_label_simple(label)
    T = d["tables"]

    def get(name):
        i = max(number - 1, 0)
        return T[name][i] if i < len(T[name]) else None
    if prefix == "BC":
        return d["bc12a"] if (is_a and number == 12) else get("BC_SEQS")
    if prefix == "NB":
        return d["bc12a"] if (is_a and number == 12) else get("NB_SEQS")
    if prefix == "AB":
        return get("AB_SEQS")
    if prefix == "BP":
        return get("BP_SEQS")
    if prefix == "RBK":
        return d["rbk_special"].get(str(number)) or get("BC_SEQS")
    return None


def _common_prefix(seqs):
    n = len(seqs[0])
    for s in seqs[1:]:
        k = 0
        while k < min(n, len(s)) and seqs[0][k] == s[k]:
            k += 1
        n = min(n, k)
        if n == 0:
            return 0
    return n


def barcode_group(seqs, labels, match_type):
    """BarcodeGroup::new (barcodes.rs:106-197) -> the dict form the oracle bindings consume (k_flank not yet set)."""
    seqs = [bytes(s) for s in seqs]
    if len(seqs) == 1:
        raise ValueError("For now we only support 'groups'")
    if any(len(s) != len(seqs[0]) for s in seqs):
        raise ValueError("All sequences per group must be equally long")
    prefix_len = _common_prefix(seqs)
    suffix_len = _common_prefix([s[::-1] for s in seqs])
    L = len(seqs[0])
    if prefix_len + suffix_len >= L:
        raise ValueError("No barcode region found")
    if prefix_len == 0 and suffix_len == 0:
        raise ValueError("No prefix or suffix found")
    mask = L - prefix_len - suffix_len
    flank = seqs[0][:prefix_len] + b"N" * mask + (seqs[0][L - suffix_len:] if suffix_len else b"")
    pad_start, pad_end = max(prefix_len - PADDING, 0), prefix_len + mask + PADDING        # pad_end NOT clamped (barcodes.rs:160-163)
    bars = [s[pad_start:min(pad_end, L)] for s in seqs]
    return dict(flank=flank, k_flank=0, bar_region=(prefix_len, prefix_len + mask - 1), pad_region=(pad_start, pad_end),
                match_type=int(match_type), bar_len=len(bars[0]), barcodes=bars, labels=list(labels),
                effective_len=prefix_len + suffix_len)


def set_flank_threshold(groups, max_flank_errors=None):
    """annotate_with_groups (annotator.rs:216-229)"""
    for g in groups:
        g["k_flank"] = int(max_flank_errors) if max_flank_errors is not None else edit_cut_off(g["effective_len"])
    return groups


def groups_from_kit(kit: str, use_extended: bool = False, max_flank_errors=None):
    """BarcodeGroup::new_from_kit (barcodes.rs:251-299) + threshold selection"""
    d = _data()
    names = dict(d["kit_names"])
    key = names.get(kit) or names.get(kit.replace(".", "-"))
    if key is None:
        raise KeyError(f"Unsupported kit: {kit}")
    groups = []
    for t in d["templates"][d["kits"][key]["templates"]]:
        if t["template_type"] == "Extended" and not use_extended:
            continue
        labels = get_barcodes(t["label_from"], t["label_to"], t["use_12a"])
        seqs = []
        for lab in labels:
            bar = lookup_barcode_seq(lab)
            if bar is None:
                raise KeyError("Barcode not found - odd - raise issue")
            seqs.append("".join(bar if p in ("{BAR}", "**") else p for p in t["parts"]).encode())
        groups.append(barcode_group(seqs, labels, FTAG if t["side"] == "Left" else RTAG))
    return set_flank_threshold(groups, max_flank_errors)


def groups_from_seqs(specs, max_flank_errors=None):
    """specs: list of (seqs, labels, match_type)"""
    return set_flank_threshold([barcode_group(s, l, t) for s, l, t in specs], max_flank_errors)


def groups_from_fasta(paths, types, max_flank_errors=None):
    """BarcodeGroup::new_from_fasta (barcodes.rs:302-315): sequences upper-cased (needletail normalize(true))"""
    specs = []
    for p, ty in zip(paths, types):
        labels, seqs, cur = [], [], None
        for line in open(p):
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                labels.append(line[1:].split()[0] if line[1:].split() else "")
                seqs.append(bytearray())
            elif seqs and line:
                seqs[-1] += line.strip().upper().encode()
        specs.append(([bytes(s) for s in seqs], labels, ty))
    return groups_from_seqs(specs, max_flank_errors)
