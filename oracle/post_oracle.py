"""CPU ORACLE for the stages that consume annotation.tsv -- TEST INFRASTRUCTURE ONLY (imported by tests/ alone).

A plain-Python restatement of rickbeeloo/barbell @ 9a2b814:
  pattern_from_str!            src/filter/pattern.rs:242-383
  match_pattern + checks       src/filter/pattern.rs:100-240
  check_filter_pass / filter   src/filter/filter.rs:10-119, 183-214
  get_group_structure          src/inspect/inspect.rs:9-117
  LabelConfig::create_label    src/trim/trim.rs:56-106
  preprocess_cuts              src/trim/trim.rs:127-248
  process_read_and_anno        src/trim/trim.rs:250-315
Pinned by the reference's own unit tests for these functions (pattern.rs:389-937, trim.rs:538-802), which
tests/test_post_stages.py replays against this file and against the C++ build.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

COLUMNS = ["read_id", "read_len", "rel_dist_to_end", "read_start_bar", "read_end_bar", "read_start_flank", "read_end_flank",
           "bar_start", "bar_end", "match_type", "flank_cost", "barcode_cost", "label", "strand", "cuts"]


@dataclass
class Row:
    read_id: str
    read_len: int
    rel_dist_to_end: int
    read_start_bar: int
    read_end_bar: int
    read_start_flank: int
    read_end_flank: int
    bar_start: int
    bar_end: int
    match_type: str
    flank_cost: int
    barcode_cost: int
    label: str
    strand: str
    cuts: List[Tuple[Tuple[str, int], int]] = field(default_factory=list)   # ((direction, group id), annotation index)

    def tsv(self):
        cuts = ",".join(f"{d}({g}):{i}" for (d, g), i in self.cuts)
        return "\t".join(str(x) for x in [self.read_id, self.read_len, self.rel_dist_to_end, self.read_start_bar, self.read_end_bar,
                                          self.read_start_flank, self.read_end_flank, self.bar_start, self.bar_end, self.match_type,
                                          self.flank_cost, self.barcode_cost, self.label, self.strand, cuts])


def parse_tsv(text: str) -> List[Row]:
    lines = [ln for ln in text.split("\n") if ln]
    if not lines:
        return []
    hdr = lines[0].split("\t")
    rows = []
    for ln in lines[1:]:
        f = dict(zip(hdr, ln.split("\t")))
        cuts = []
        if f.get("cuts"):
            for part in f["cuts"].split(","):
                c, pos = part.split(":")[:2]
                d, g = c.rstrip(")").split("(")
                cuts.append(((d, int(g)), int(pos)))
        rows.append(Row(f["read_id"], int(f["read_len"]), int(f["rel_dist_to_end"]), int(f["read_start_bar"]), int(f["read_end_bar"]),
                        int(f["read_start_flank"]), int(f["read_end_flank"]), int(f["bar_start"]), int(f["bar_end"]), f["match_type"],
                        int(f["flank_cost"]), int(f["barcode_cost"]), f["label"], f["strand"], cuts))
    return rows


def to_tsv(rows: List[Row]) -> str:
    return "" if not rows else "\t".join(COLUMNS) + "\n" + "".join(r.tsv() + "\n" for r in rows)


# ---------------- pattern.rs ----------------
@dataclass
class Element:
    match_type: str
    orientation: Optional[str] = None
    label: Optional[str] = None
    placeholder: Optional[int] = None
    range: Tuple[int, int] = (0, 0)
    relative_to: Optional[str] = None
    cuts: List[Tuple[str, int]] = field(default_factory=list)


class PatternError(Exception):
    pass


def _int(s):
    s = s.strip()
    t = s[1:] if s[:1] in "+-" else s
    if not t or not t.isdigit():
        return None
    return int(s)


def _parse_position(p):
    parts = p.split("(")
    if len(parts) != 2:
        return None
    name = parts[0].lstrip("@")
    rel = {"left": "left", "right": "right", "prev_left": "prev_left"}.get(name)
    if rel is None:
        return None
    rng = p[len(parts[0]):].strip().strip("()").split("..")
    if len(rng) != 2:
        return None
    a, b = _int(rng[0]), _int(rng[1])
    if a is None or b is None:
        return None
    return rel, (a, b)


def _parse_element(s):
    parts = s.split("[", 1)
    if len(parts) != 2:
        return None
    t = parts[0].strip()
    if t in ("Flank", "flank"):
        raise PatternError("Flank is not valid, use Fflank or Rflank")
    if t not in ("Ftag", "Rtag", "Fflank", "Rflank"):
        return None
    el = Element(t)
    for param in (x.strip() for x in parts[1].rstrip("]").split(",")):
        if param == "fw":
            el.orientation = "Fwd"
        elif param == "rc":
            el.orientation = "Rc"
        elif param.startswith("@"):
            r = _parse_position(param)
            if r:
                el.relative_to, el.range = r
        elif param.startswith("?"):
            if param[1:].isdigit():
                el.placeholder = int(param[1:])
        elif param.startswith(">") or param.startswith("<"):
            if len(param) < 2:
                raise PatternError("cut marker too short")
            gid = 0 if len(param) == 2 else (int(param[2:]) if param[2:].isdigit() else None)
            if gid is not None and param[:2] == ">>":
                el.cuts.append(("After", gid))
            elif gid is not None and param[:2] == "<<":
                el.cuts.append(("Before", gid))
        elif param == "*":
            pass
        else:
            el.label = param.strip('"')
    return el


def parse_pattern(s: str) -> List[Element]:
    els = [e for e in (_parse_element(x.strip()) for x in s.split("__")) if e is not None]
    if s.count("__") + 1 != len(els):
        raise PatternError(f"Pattern parse error for: {s!r}")
    return els


def _matches(m: Row, el: Element, prev_end, labels) -> bool:
    if m.match_type != el.match_type:
        return False
    if el.match_type in ("Ftag", "Rtag") and el.label is not None:
        if el.label.startswith("~"):
            if el.label[1:] not in m.label:
                return False
        elif el.label != m.label:
            return False
    if el.placeholder is not None:
        if el.placeholder in labels:
            if m.label != labels[el.placeholder]:
                return False
        else:
            labels[el.placeholder] = m.label
    if el.orientation is not None and el.orientation != m.strand:
        return False
    if el.relative_to == "left":
        if not (el.range[0] <= m.read_start_bar <= el.range[1]):
            return False
    elif el.relative_to == "right":
        if not (m.read_len - el.range[1] <= m.read_end_bar <= m.read_len - el.range[0]):
            return False
    elif el.relative_to == "prev_left" and prev_end is not None:
        if not (prev_end + el.range[0] <= m.read_start_bar <= prev_end + el.range[1]):
            return False
    return True


def match_pattern(matches: List[Row], pattern: List[Element]):
    if len(matches) < len(pattern):
        return False, []
    prev_end, labels, cuts = None, {}, []
    for idx, el in enumerate(pattern):
        m = matches[idx]
        if not _matches(m, el, prev_end, labels):
            return False, []
        cuts += [(idx, c) for c in el.cuts]
        prev_end = m.read_end_bar
    return True, cuts


def check_filter_pass(annotations: List[Row], patterns) -> bool:
    best_n, best = 0, None
    for p in patterns:
        ok, cuts = match_pattern(annotations, p)
        if ok and len(p) > best_n:
            best_n, best = len(p), cuts
    if best_n > 0:
        for idx, c in best:
            annotations[idx].cuts.append((c, idx))
    return best_n == len(annotations)


def group_reads(rows: List[Row]):
    groups = []
    for r in rows:
        if groups and groups[-1][0].read_id == r.read_id:
            groups[-1].append(r)
        else:
            groups.append([r])
    return groups


def filter_rows(rows: List[Row], patterns):
    kept, dropped = [], []
    for g in group_reads(rows):
        (kept if check_filter_pass(g, patterns) else dropped).extend(g)
    return kept, dropped


# ---------------- inspect.rs ----------------
def _bucket(pos, size):
    return (max(pos - 1, 0) // size) * size


def group_structure(group: List[Row], bucket: int) -> str:
    out, prev_end = [], None
    for a in group:
        start, end = a.read_start_bar, a.read_end_bar

        def right():
            return f"@right({_bucket(max(a.read_len - end, 0), bucket)}..{_bucket(max(a.read_len - start, 0), bucket) + bucket})"
        if prev_end is not None:
            d_prev, d_right = max(start - prev_end, 0), max(a.read_len - end, 0)
            if d_prev <= d_right:
                g0 = _bucket(d_prev, bucket)
                tag = f"@prev_left({g0}..{g0 + bucket})"
            else:
                tag = right()
        elif a.rel_dist_to_end > 0:
            s0 = _bucket(start, bucket)
            tag = f"@left({s0}..{s0 + bucket})"
        else:
            tag = right()
        cut = "" if not a.cuts else (", <<" if a.strand == "Fwd" else ", >>")
        out.append(f"{a.match_type}[{'fw' if a.strand == 'Fwd' else 'rc'}, *{cut}, {tag}]")
        prev_end = end
    return "__".join(out)


# ---------------- trim.rs ----------------
def create_label(annos: List[Row], add_labels=True, add_orientation=True, add_flank=True, sort_labels=False, only_side=None) -> str:
    if not add_labels:
        return "none"
    parts = []
    for m in annos:
        if not add_flank and "flank" in m.label:
            continue
        parts.append(m.label + (("_fw" if m.strand == "Fwd" else "_rc") if add_orientation else ""))
    if not parts:
        return "none"
    if sort_labels:
        return "__".join(sorted(parts))
    if only_side == "left":
        return parts[0]
    if only_side == "right":
        return parts[-1]
    return "__".join(parts)


def preprocess_cuts(annotations: List[Row], seq_len: int):
    groups = {}
    for a in annotations:
        for (d, g), _ in a.cuts:
            groups.setdefault(g, []).append((a.read_start_flank, a.read_end_flank, d, a))
    ordered = [groups[g] for g in sorted(groups)]                 # ties of the stable sort: group-id order (HashMap order upstream)
    ordered.sort(key=lambda grp: grp[0][0])
    slices = []
    for i, grp in enumerate(ordered):
        if len(grp) == 2:
            (s1, e1, d1, a1), (s2, e2, d2, a2) = grp
            slices.append((s1 if d1 == "Before" else e1, s2 if d2 == "Before" else e2, [a1, a2]))
        elif len(grp) == 1:
            s, e, d, a = grp[0]
            if d == "Before":
                if i > 0:
                    prev = ordered[i - 1]
                    best = max(range(len(prev)), key=lambda q: (prev[q][1], q))      # max_by_key: last maximum
                    slices.append((prev[best][1], s, [prev[best][3], a]))
                else:
                    slices.append((0, s, [a]))
            else:
                if i + 1 < len(ordered):
                    nxt = ordered[i + 1]
                    best = min(range(len(nxt)), key=lambda q: (nxt[q][0], q))        # min_by_key: first minimum
                    slices.append((e, nxt[best][0], [a, nxt[best][3]]))
                else:
                    slices.append((e, seq_len, [a]))
    return slices


_RC = bytes.maketrans(b"ACTGRYSWKMBDHVNXactgryswkmbdhvnx", b"TGACYRSWMKVHDBNXtgacyrswmkvhdbnx")


def process_read_and_anno(seq: bytes, qual: bytes, annotations: List[Row], skip_trim=False, flip=False, **label_kw):
    out = []
    for n, (start, end, annos) in enumerate(preprocess_cuts(annotations, len(seq))):
        if start >= end:
            continue
        s, q = (seq, qual) if skip_trim else (seq[start:end], qual[start:end])
        if flip and any(a.match_type == "Ftag" and a.strand == "Rc" for a in annos):
            s, q = s.translate(_RC)[::-1], q[::-1]
        out.append((s, q, create_label(annos, **label_kw), "" if n == 0 else f"_{n}"))
    return out
