"""ctypes bindings for the CPU oracle (oracle/libbarbell_oracle.so). Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "libbarbell_oracle.so")

OP_MATCH, OP_SUB, OP_TEXT, OP_PAT = 0, 1, 2, 3
FWD, RC = 0, 1
FTAG, RTAG, FFLANK, RFLANK = 0, 1, 2, 3
MATCH_TYPE_NAMES = ["Ftag", "Rtag", "Fflank", "Rflank"]
STRAND_NAMES = ["Fwd", "Rc"]


class OrcMatch(C.Structure):
    _fields_ = [("text_start", C.c_int32), ("text_end", C.c_int32), ("pattern_start", C.c_int32),
                ("pattern_end", C.c_int32), ("cost", C.c_int32), ("strand", C.c_int32), ("n_ops", C.c_int32),
                ("ops", C.POINTER(C.c_uint8))]


class OrcGroup(C.Structure):
    _fields_ = [("flank", C.c_char_p), ("flank_len", C.c_int32), ("k_flank", C.c_int32), ("bar0", C.c_int32),
                ("bar1", C.c_int32), ("pad0", C.c_int32), ("pad1", C.c_int32), ("match_type", C.c_int32),
                ("n_barcodes", C.c_int32), ("bar_len", C.c_int32), ("barcodes", C.c_char_p)]


class OrcParams(C.Structure):
    _fields_ = [("alpha", C.c_float), ("min_score", C.c_double), ("min_score_diff", C.c_double)]


ROW_DTYPE = np.dtype([
    ("read_idx", "<u4"), ("read_len", "<u4"), ("rel_dist_to_end", "<i8"), ("read_start_bar", "<i8"),
    ("read_end_bar", "<i8"), ("read_start_flank", "<i8"), ("read_end_flank", "<i8"), ("bar_start", "<i8"),
    ("bar_end", "<i8"), ("flank_cost", "<i4"), ("barcode_cost", "<i4"), ("label_idx", "<i4"), ("group_idx", "<i4"),
    ("match_type", "u1"), ("strand", "u1"), ("pad_", "u1", (6,))])
assert ROW_DTYPE.itemsize == 88


class OrcPolicy(C.Structure):
    _fields_ = [("use_myers", C.c_int), ("flags", C.c_int)]


# orc_policy.flags == bb_opts.policy bits
POL_S1_LEFT, POL_S2_PAT_FIRST, POL_S5_LAST, POL_S6_RC_FIRST, POL_S3_ROUND, POL_S3_CEIL = 1, 2, 4, 8, 16, 32
_cur_policy = [1, 0]


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_search.restype = C.c_int
        L.orc_search.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_float, C.c_int,
                                 C.POINTER(C.POINTER(OrcMatch))]
        L.orc_free_matches.argtypes = [C.POINTER(OrcMatch), C.c_int]
        L.orc_to_path.argtypes = [C.POINTER(OrcMatch), C.POINTER(C.c_int32)]
        L.orc_get_matching_region.restype = C.c_int
        L.orc_get_matching_region.argtypes = [C.POINTER(OrcMatch), C.c_int, C.c_int, C.POINTER(C.c_int64),
                                              C.POINTER(C.c_int64)]
        L.orc_map_pat_to_text_with_cost.restype = C.c_int
        L.orc_map_pat_to_text_with_cost.argtypes = [C.POINTER(OrcMatch), C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.orc_lodhi.restype = C.c_double
        L.orc_lodhi.argtypes = [C.c_char_p, C.c_int]
        L.orc_bottom_row.restype = C.c_int
        L.orc_bottom_row.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_float, C.c_void_p]
        L.orc_edit_cut_off.restype = C.c_int
        L.orc_edit_cut_off.argtypes = [C.c_int]
        L.orc_collapse.restype = C.c_int
        L.orc_collapse.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.orc_demux_batch.restype = C.c_int64
        L.orc_demux_batch.argtypes = [C.POINTER(OrcGroup), C.c_int, C.POINTER(OrcParams), C.c_void_p, C.c_void_p,
                                      C.c_uint32, C.c_int, C.c_void_p, C.c_int64]
        L.orc_flank_hits_batch.restype = C.c_int64
        L.orc_flank_hits_batch.argtypes = [C.POINTER(OrcGroup), C.c_int, C.POINTER(OrcParams), C.c_void_p, C.c_void_p,
                                           C.c_uint32, C.c_int, C.c_void_p, C.c_int64]
        L.orc_max_threads.restype = C.c_int
        L.orc_set_policy.argtypes = [C.POINTER(OrcPolicy)]
        _lib = L
    return _lib


def set_policy(use_myers=None, flags=None):
    if use_myers is not None:
        _cur_policy[0] = int(use_myers)
    if flags is not None:
        _cur_policy[1] = int(flags)
    p = OrcPolicy(_cur_policy[0], _cur_policy[1])
    lib().orc_set_policy(C.byref(p))


class policy:
    """with O.policy(flags): ... -- runs the oracle under the given ORC_POL_* bits and restores the previous setting"""

    def __init__(self, flags):
        self.flags = flags

    def __enter__(self):
        self.saved = _cur_policy[1]
        set_policy(flags=self.flags)

    def __exit__(self, *a):
        set_policy(flags=self.saved)


class Match:
    """Python copy of an orc_match plus the derived helpers of the reference's cigar_parse.rs."""

    def __init__(self, m):
        self.text_start, self.text_end = m.text_start, m.text_end
        self.pattern_start, self.pattern_end = m.pattern_start, m.pattern_end
        self.cost, self.strand = m.cost, m.strand
        self.ops = bytes(m.ops[i] for i in range(m.n_ops))

    def _c(self):
        buf = (C.c_uint8 * max(1, len(self.ops)))(*self.ops)
        m = OrcMatch(self.text_start, self.text_end, self.pattern_start, self.pattern_end, self.cost, self.strand,
                     len(self.ops), C.cast(buf, C.POINTER(C.c_uint8)))
        m._keep = buf
        return m

    def path(self):
        m = self._c()
        ij = (C.c_int32 * (2 * max(1, len(self.ops))))()
        lib().orc_to_path(C.byref(m), ij)
        return [(ij[2 * q], ij[2 * q + 1]) for q in range(len(self.ops))]

    def matching_region(self, start, end):
        m = self._c()
        a, b = C.c_int64(), C.c_int64()
        ok = lib().orc_get_matching_region(C.byref(m), start, end, C.byref(a), C.byref(b))
        return (a.value, b.value) if ok else None

    def map_pat_to_text_with_cost(self, ps, pe):
        m = self._c()
        out = (C.c_int64 * 5)()
        ok = lib().orc_map_pat_to_text_with_cost(C.byref(m), ps, pe, out)
        return ((out[0], out[1]), (out[2], out[3]), out[4]) if ok else None

    def cigar(self):
        return "".join("=XID"[o] for o in self.ops)

    def __repr__(self):
        return (f"Match(t={self.text_start}..{self.text_end}, p={self.pattern_start}..{self.pattern_end}, "
                f"cost={self.cost}, strand={STRAND_NAMES[self.strand]}, {self.cigar()})")


def search(pattern: bytes, text: bytes, k: int, alpha: float = -1.0, rc: bool = True):
    out = C.POINTER(OrcMatch)()
    n = lib().orc_search(pattern, len(pattern), text, len(text), k, alpha, int(rc), C.byref(out))
    res = [Match(out[i]) for i in range(n)]
    lib().orc_free_matches(out, n)
    return res


def lodhi(cigar: str) -> float:
    ops = bytes("=XID".index(ch) for ch in cigar)
    return lib().orc_lodhi(ops, len(ops))


def bottom_row(pattern: bytes, text: bytes, alpha: float = -1.0):
    c = np.zeros(len(text) + len(pattern) + 2, dtype=np.int32)
    n = lib().orc_bottom_row(pattern, len(pattern), text, len(text), alpha, c.ctypes.data)
    return c[:n].copy()


def make_groups(groups):
    """groups: list of dicts (see barbell_b200.host.Group.as_dict). Returns (ctypes array, keepalive)."""
    arr = (OrcGroup * len(groups))()
    keep = []
    for i, g in enumerate(groups):
        flank = bytes(g["flank"])
        bars = b"".join(bytes(b) for b in g["barcodes"])
        keep += [flank, bars]
        arr[i] = OrcGroup(flank, len(flank), g["k_flank"], g["bar_region"][0], g["bar_region"][1], g["pad_region"][0],
                          g["pad_region"][1], g["match_type"], len(g["barcodes"]), g["bar_len"], bars)
    return arr, keep


def demux_batch(groups, bases: np.ndarray, offsets: np.ndarray, alpha=0.4, min_score=0.2, min_score_diff=0.1,
                n_threads=0, cap_per_read=8):
    arr, keep = make_groups(groups)
    prm = OrcParams(alpha, min_score, min_score_diff)
    n_reads = len(offsets) - 1
    cap = max(16, n_reads * cap_per_read)
    rows = np.zeros(cap, dtype=ROW_DTYPE)
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    if n_threads <= 0:
        n_threads = lib().orc_max_threads()
    n = lib().orc_demux_batch(arr, len(groups), C.byref(prm), bases.ctypes.data, offsets.ctypes.data, n_reads,
                              n_threads, rows.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("oracle row buffer overflow")
    return rows[:n].copy()


def flank_hits_batch(groups, bases, offsets, alpha=0.4, n_threads=0, cap_per_read=16):
    arr, keep = make_groups(groups)
    prm = OrcParams(alpha, 0.2, 0.1)
    n_reads = len(offsets) - 1
    cap = max(64, n_reads * cap_per_read)
    out = np.zeros((cap, 6), dtype=np.int32)
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    if n_threads <= 0:
        n_threads = lib().orc_max_threads()
    n = lib().orc_flank_hits_batch(arr, len(groups), C.byref(prm), bases.ctypes.data, offsets.ctypes.data, n_reads,
                                   n_threads, out.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("oracle hit buffer overflow")
    return out[:n].copy()
