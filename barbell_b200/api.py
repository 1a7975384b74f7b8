"""ctypes mirror of include/barbell_b200.h (names follow the reference: BarcodeGroup -> GroupSet, Demuxer -> Annotator)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FTAG, RTAG, FFLANK, RFLANK = 0, 1, 2, 3
MATCH_TYPE_NAMES = ["Ftag", "Rtag", "Fflank", "Rflank"]      # reference src/annotate/barcodes.rs:26-33
STRAND_NAMES = ["Fwd", "Rc"]                                  # reference src/annotate/searcher.rs:100-142

ROW_DTYPE = np.dtype([
    ("read_idx", "<u4"), ("read_len", "<u4"), ("rel_dist_to_end", "<i8"), ("read_start_bar", "<i8"),
    ("read_end_bar", "<i8"), ("read_start_flank", "<i8"), ("read_end_flank", "<i8"), ("bar_start", "<i8"),
    ("bar_end", "<i8"), ("flank_cost", "<i4"), ("barcode_cost", "<i4"), ("label_idx", "<i4"), ("group_idx", "<i4"),
    ("match_type", "u1"), ("strand", "u1"), ("pad_", "u1", (6,))])
assert ROW_DTYPE.itemsize == 88


class BarbellError(RuntimeError):
    pass


class _Group(C.Structure):
    _fields_ = [("flank", C.c_char_p), ("flank_len", C.c_int32), ("k_flank", C.c_int32), ("bar0", C.c_int32),
                ("bar1", C.c_int32), ("pad0", C.c_int32), ("pad1", C.c_int32), ("match_type", C.c_int32),
                ("n_barcodes", C.c_int32), ("bar_len", C.c_int32), ("barcodes", C.c_void_p)]


class _Opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("alpha", C.c_float), ("min_score", C.c_double), ("min_score_diff", C.c_double),
                ("max_batch_bytes", C.c_uint64), ("max_batch_reads", C.c_uint32), ("flags", C.c_uint32), ("policy", C.c_uint32)]


EXPORTS = ["bb_groups_from_kit", "bb_groups_from_fasta", "bb_groups_add", "bb_groups_set_flank_threshold",
           "bb_groups_count", "bb_groups_data", "bb_groups_label", "bb_groups_free", "bb_edit_cut_off", "bb_label_range",
           "bb_lookup_barcode_seq", "bb_create",
           "bb_destroy", "bb_last_error", "bb_set_groups", "bb_annotate", "bb_annotate_device", "bb_fetch_rows",
           "bb_submit", "bb_submit_packed", "bb_reserve", "bb_pack_crumbs_append", "bb_pack_crumbs_append_line", "bb_collect", "bb_counters", "bb_host_alloc", "bb_host_free", "bb_pack_nibbles", "bb_pack_crumbs", "bb_last_stage_ms", "bb_kernel_launches", "bb_h2d_bytes", "bb_fetch_flank_hits",
           "bb_kit_info", "bb_kit_filter_patterns", "bb_pattern_parse", "bb_filter", "bb_inspect", "bb_trim",
           "bb_abi_version"]

_lib = None


def lib_path():
    return os.path.join(_HERE, "libbarbell_b200.so")


def lib():
    """Load libbarbell_b200.so; fails loudly when it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise BarbellError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(or `make -C barbell_b200`). barbell_b200 has no CPU fallback.")
    L = C.CDLL(p)
    vp, i32, u32, u64 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64
    L.bb_groups_from_kit.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp), C.c_char_p, C.c_size_t]
    L.bb_groups_from_fasta.argtypes = [C.POINTER(C.c_char_p), C.POINTER(i32), i32, C.POINTER(vp), C.c_char_p, C.c_size_t]
    L.bb_groups_add.argtypes = [C.POINTER(vp), C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), i32, i32, C.c_char_p, C.c_size_t]
    L.bb_groups_set_flank_threshold.argtypes = [vp, i32]
    L.bb_groups_count.argtypes = [vp]; L.bb_groups_count.restype = i32
    L.bb_groups_data.argtypes = [vp]; L.bb_groups_data.restype = C.POINTER(_Group)
    L.bb_groups_label.argtypes = [vp, i32, i32]; L.bb_groups_label.restype = C.c_char_p
    L.bb_groups_free.argtypes = [vp]; L.bb_groups_free.restype = None
    L.bb_edit_cut_off.argtypes = [i32]; L.bb_edit_cut_off.restype = i32
    L.bb_label_range.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
    L.bb_lookup_barcode_seq.argtypes = [C.c_char_p]; L.bb_lookup_barcode_seq.restype = C.c_char_p
    L.bb_create.argtypes = [C.POINTER(_Opts), C.POINTER(vp), C.c_char_p, C.c_size_t]
    L.bb_destroy.argtypes = [vp]; L.bb_destroy.restype = None
    L.bb_last_error.argtypes = [vp]; L.bb_last_error.restype = C.c_char_p
    L.bb_set_groups.argtypes = [vp, C.POINTER(_Group), i32]
    L.bb_annotate.argtypes = [vp, vp, vp, u32, vp, u64, C.POINTER(u64)]
    L.bb_annotate_device.argtypes = [vp, vp, vp, u32, u64, vp, C.POINTER(u64)]
    L.bb_fetch_rows.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.bb_submit.argtypes = [vp, vp, vp, u32, u64]
    L.bb_collect.argtypes = [vp, C.POINTER(u64), C.POINTER(vp), C.POINTER(u64)]
    L.bb_submit_packed.argtypes = [vp, vp, u64, vp, u64, vp, u32, u64]
    L.bb_reserve.argtypes = [vp, C.c_uint32, u64]
    L.bb_pack_crumbs_append.argtypes = [vp, u64, vp, C.POINTER(u64), vp, u64, C.POINTER(u64)]
    L.bb_pack_crumbs_append_line.argtypes = [vp, u64, vp, C.POINTER(u64), vp, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_int)]
    L.bb_counters.argtypes = [vp, C.POINTER(u64)]
    L.bb_last_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.bb_kernel_launches.argtypes = [vp]; L.bb_kernel_launches.restype = u64
    L.bb_h2d_bytes.argtypes = [vp]; L.bb_h2d_bytes.restype = u64
    L.bb_fetch_flank_hits.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.bb_host_alloc.argtypes = [C.c_size_t]; L.bb_host_alloc.restype = C.c_void_p
    L.bb_host_free.argtypes = [C.c_void_p]; L.bb_host_free.restype = None
    L.bb_pack_nibbles.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    L.bb_pack_crumbs.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.bb_abi_version.restype = C.c_int
    _lib = L
    return L


def edit_cut_off(effective_len: int) -> int:
    """reference src/annotate/edit_model.rs:2-11"""
    return lib().bb_edit_cut_off(effective_len)


def label_range(from_label: str, to_label: str, use_12a: bool = False):
    """reference src/kits/kits.rs:741-816 (get_barcodes)"""
    buf = C.create_string_buffer(4096)
    n = lib().bb_label_range(from_label.encode(), to_label.encode(), int(use_12a), buf, 4096)
    if n < 0:
        raise BarbellError(buf.value.decode())
    return buf.value.decode().split(",") if n else []


def lookup_barcode_seq(label: str):
    """reference src/kits/kits.rs:1074-1103"""
    s = lib().bb_lookup_barcode_seq(label.encode())
    return s.decode() if s is not None else None


class GroupSet:
    """The query groups of one run (reference Vec<BarcodeGroup>, src/annotate/barcodes.rs:57-71)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_kit(cls, kit: str, use_extended: bool = False, max_flank_errors=None):
        """BarcodeGroup::new_from_kit + annotate_with_groups' threshold selection (barcodes.rs:251, annotator.rs:216)."""
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = lib().bb_groups_from_kit(kit.encode(), int(use_extended), C.byref(h), err, 512)
        if rc != 0:
            raise BarbellError(err.value.decode())
        gs = cls(h)
        gs.set_flank_threshold(max_flank_errors)
        return gs

    @classmethod
    def from_fasta(cls, paths, types, max_flank_errors=None):
        """BarcodeGroup::new_from_fasta per query file (barcodes.rs:302-315; bin/main.rs:78-96)."""
        n = len(paths)
        arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
        ty = (C.c_int32 * n)(*[int(t) for t in types])
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = lib().bb_groups_from_fasta(arr, ty, n, C.byref(h), err, 512)
        if rc != 0:
            raise BarbellError(err.value.decode())
        gs = cls(h)
        gs.set_flank_threshold(max_flank_errors)
        return gs

    @classmethod
    def from_seqs(cls, groups, max_flank_errors=None):
        """groups: list of (seqs, labels, type) -- BarcodeGroup::new (barcodes.rs:106-197)."""
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        for seqs, labels, ty in groups:
            n = len(seqs)
            a = (C.c_char_p * n)(*[s if isinstance(s, bytes) else s.encode() for s in seqs])
            b = (C.c_char_p * n)(*[s.encode() for s in labels])
            rc = lib().bb_groups_add(C.byref(h), a, b, n, int(ty), err, 512)
            if rc != 0:
                if h:
                    lib().bb_groups_free(h)
                raise BarbellError(err.value.decode())
        gs = cls(h)
        gs.set_flank_threshold(max_flank_errors)
        return gs

    def set_flank_threshold(self, max_flank_errors=None):
        lib().bb_groups_set_flank_threshold(self._h, -1 if max_flank_errors is None else int(max_flank_errors))

    def __len__(self):
        return lib().bb_groups_count(self._h)

    def _data(self):
        return lib().bb_groups_data(self._h)

    def label(self, group_idx, label_idx):
        s = lib().bb_groups_label(self._h, int(group_idx), int(label_idx))
        return s.decode() if s is not None else "?"

    def as_dicts(self):
        """Plain-python description of every group (what the oracle bindings consume)."""
        out = []
        d = self._data()
        for g in range(len(self)):
            G = d[g]
            bars = C.string_at(G.barcodes, G.n_barcodes * G.bar_len)
            out.append(dict(flank=C.string_at(G.flank, G.flank_len), k_flank=G.k_flank, bar_region=(G.bar0, G.bar1),
                            pad_region=(G.pad0, G.pad1), match_type=G.match_type, bar_len=G.bar_len,
                            barcodes=[bars[i * G.bar_len:(i + 1) * G.bar_len] for i in range(G.n_barcodes)],
                            labels=[self.label(g, i) for i in range(G.n_barcodes)]))
        return out

    def __del__(self):
        try:
            if self._h:
                lib().bb_groups_free(self._h)
                self._h = None
        except Exception:
            pass


class Annotator:
    """One GPU context = the reference's Demuxer (src/annotate/searcher.rs:12-29, 202-227, 430-490) in batch form."""

    def __init__(self, groups: GroupSet, device=0, alpha=0.4, min_score=0.2, min_score_diff=0.1, use_filter=True,
                 pack_h2d=False, policy=0):
        self._ctx = C.c_void_p()
        self.groups = groups
        o = _Opts(device, alpha, min_score, min_score_diff, 0, 0, (0 if use_filter else 1) | (4 if pack_h2d == "crumbs" else 2 if pack_h2d else 0), policy)
        err = C.create_string_buffer(512)
        rc = lib().bb_create(C.byref(o), C.byref(self._ctx), err, 512)
        if rc != 0:
            raise BarbellError(f"bb_create failed ({rc}): {err.value.decode()}")
        self._check(lib().bb_set_groups(self._ctx, groups._data(), len(groups)))

    def _check(self, rc):
        if rc != 0:
            raise BarbellError(f"barbell_b200 error {rc}: {lib().bb_last_error(self._ctx).decode()}")

    def annotate(self, bases: np.ndarray, offsets: np.ndarray, rows_cap=None) -> np.ndarray:
        """Host buffers in, rows out (bb_annotate)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n_reads = len(offsets) - 1
        cap = rows_cap if rows_cap is not None else max(64, 4 * n_reads)
        rows = np.zeros(cap, dtype=ROW_DTYPE)
        n = C.c_uint64()
        rc = lib().bb_annotate(self._ctx, bases.ctypes.data, offsets.ctypes.data, n_reads, rows.ctypes.data, cap, C.byref(n))
        if rc == -3 and rows_cap is None:        # BB_ERR_OVERFLOW reports the needed size: fetch the rows still on the device
            return self.fetch_rows(n.value)
        self._check(rc)
        return rows[:n.value].copy()

    def annotate_ptr(self, bases_ptr, offsets_ptr, n_reads, rows_ptr, rows_cap) -> int:
        n = C.c_uint64()
        self._check(lib().bb_annotate(self._ctx, bases_ptr, offsets_ptr, n_reads, rows_ptr, rows_cap, C.byref(n)))
        return n.value

    def annotate_device(self, d_bases_ptr, d_offsets_ptr, n_reads, total_bytes, stream=0) -> int:
        """Device-resident inputs (bb_annotate_device); returns the number of rows left on the device."""
        n = C.c_uint64()
        self._check(lib().bb_annotate_device(self._ctx, d_bases_ptr, d_offsets_ptr, n_reads, total_bytes, stream or None,
                                             C.byref(n)))
        return n.value

    def fetch_rows(self, n_rows) -> np.ndarray:
        rows = np.zeros(max(1, n_rows), dtype=ROW_DTYPE)
        n = C.c_uint64()
        self._check(lib().bb_fetch_rows(self._ctx, rows.ctypes.data, len(rows), C.byref(n)))
        return rows[:n.value].copy()

    def submit(self, bases_ptr, offsets_ptr, n_reads, tag=0):
        self._check(lib().bb_submit(self._ctx, bases_ptr, offsets_ptr, n_reads, tag))

    def reserve(self, max_reads, max_bases):
        """Optional: allocate every engine's buffers for batches of this shape and load the kernels before the first submit."""
        self._check(lib().bb_reserve(self._ctx, max_reads, max_bases))

    def submit_packed(self, crumbs_ptr, n_bases, exc_ptr, n_exc, offsets_ptr, n_reads, tag=0):
        self._check(lib().bb_submit_packed(self._ctx, crumbs_ptr, n_bases, exc_ptr, n_exc, offsets_ptr, n_reads, tag))

    def collect(self, copy=True):
        tag, ptr, n = C.c_uint64(), C.c_void_p(), C.c_uint64()
        self._check(lib().bb_collect(self._ctx, C.byref(tag), C.byref(ptr), C.byref(n)))
        if not copy:
            return tag.value, ptr.value, n.value
        if n.value == 0:
            return tag.value, np.zeros(0, dtype=ROW_DTYPE)
        buf = (C.c_char * (n.value * ROW_DTYPE.itemsize)).from_address(ptr.value)
        return tag.value, np.frombuffer(buf, dtype=ROW_DTYPE).copy()

    def flank_hits(self) -> np.ndarray:
        cap = 1 << 22
        out = np.zeros((cap, 6), dtype=np.int32)
        n = C.c_uint64()
        self._check(lib().bb_fetch_flank_hits(self._ctx, out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    def counters(self):
        out = (C.c_uint64 * 3)()
        lib().bb_counters(self._ctx, out)
        return dict(total=out[0], kept=out[1], dropped=out[2])

    def stage_ms(self):
        out = (C.c_float * 5)()
        lib().bb_last_stage_ms(self._ctx, out)
        return dict(zip(["scan", "sort_resolve", "trace", "barcode", "collapse"], [float(x) for x in out]))

    def kernel_launches(self) -> int:
        return int(lib().bb_kernel_launches(self._ctx))

    def h2d_bytes(self) -> int:
        """Bytes copied host -> device by annotate() / submit() so far."""
        return int(lib().bb_h2d_bytes(self._ctx))

    def close(self):
        if self._ctx:
            lib().bb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


TSV_HEADER = ("read_id\tread_len\trel_dist_to_end\tread_start_bar\tread_end_bar\tread_start_flank\tread_end_flank\t"
              "bar_start\tbar_end\tmatch_type\tflank_cost\tbarcode_cost\tlabel\tstrand\tcuts")


def rows_to_tsv(rows: np.ndarray, groups: GroupSet, read_ids) -> str:
    """annotation.tsv text of `rows` (column order of BarbellMatch, reference src/annotate/searcher.rs:31-64;
    the csv writer emits the header with the first row, so zero rows give an empty file: annotator.rs:20-24)."""
    if len(rows) == 0:
        return ""
    lines = [TSV_HEADER]
    for r in rows:
        lines.append("\t".join(str(x) for x in (
            read_ids[int(r["read_idx"])], int(r["read_len"]), int(r["rel_dist_to_end"]), int(r["read_start_bar"]),
            int(r["read_end_bar"]), int(r["read_start_flank"]), int(r["read_end_flank"]), int(r["bar_start"]),
            int(r["bar_end"]), MATCH_TYPE_NAMES[int(r["match_type"])], int(r["flank_cost"]), int(r["barcode_cost"]),
            groups.label(int(r["group_idx"]), int(r["label_idx"])), STRAND_NAMES[int(r["strand"])], "")))
    return "\n".join(lines) + "\n"
