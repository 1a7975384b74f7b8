"""Host-side mirror of the reference's filter / inspect / trim entry points over the C ABI (no GPU involved).

  filter(annotated_file, output_file, dropped_out_file, patterns)      src/filter/filter.rs:10-119
  kit_patterns(kit, maximize)                                          src/kits/kits.rs:175-236, 635-708
  inspect(annotated_file, top_n, read_pattern_out, bucket_size)        src/inspect/inspect.rs:133-208
  trim_matches(filtered_match_file, read_fastq_files, output_folder)   src/trim/trim.rs:317-480
"""
import ctypes as C

from .api import BarbellError, lib


class _TrimOpts(C.Structure):
    _fields_ = [("add_labels", C.c_int32), ("add_orientation", C.c_int32), ("add_flank", C.c_int32), ("sort_labels", C.c_int32),
                ("only_side", C.c_int32), ("write_full_header", C.c_int32), ("skip_trim", C.c_int32), ("flip", C.c_int32),
                ("gzip", C.c_int32), ("failed_out", C.c_char_p), ("threads", C.c_int32)]


def _enc(s):
    return None if s is None else str(s).encode()


def kit_info(kit):
    """(preset name, ["NB01 - NB96", ...], double_label) -- get_kit_info, kits.rs:635-708."""
    L = lib()
    L.bb_kit_info.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
    name, ranges, dbl, err = C.create_string_buffer(64), C.create_string_buffer(512), C.c_int(), C.create_string_buffer(512)
    if L.bb_kit_info(_enc(kit), name, 64, ranges, 512, C.byref(dbl), err, 512) != 0:
        raise BarbellError(err.value.decode())
    return name.value.decode(), ranges.value.decode().split("; "), bool(dbl.value)


def kit_patterns(kit, maximize=False):
    L = lib()
    L.bb_kit_filter_patterns.argtypes = [C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_char_p)), C.POINTER(C.c_int32)]
    arr, n = C.POINTER(C.c_char_p)(), C.c_int32()
    L.bb_kit_filter_patterns(int(kit_info(kit)[2]), int(maximize), C.byref(arr), C.byref(n))
    return [arr[i].decode() for i in range(n.value)]


def filter(annotated_file, output_file, dropped_out_file=None, patterns=()):   # noqa: A001 (the reference's name)
    """Returns {total, kept, dropped} read counts."""
    L = lib()
    L.bb_filter.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int32, C.POINTER(C.c_uint64), C.c_char_p, C.c_size_t]
    arr = (C.c_char_p * max(1, len(patterns)))(*[p.encode() for p in patterns])
    counts, err = (C.c_uint64 * 3)(), C.create_string_buffer(1024)
    if L.bb_filter(_enc(annotated_file), _enc(output_file), _enc(dropped_out_file), arr, len(patterns), counts, err, 1024) != 0:
        raise BarbellError(err.value.decode())
    return dict(total=counts[0], kept=counts[1], dropped=counts[2])


def inspect(annotated_file, top_n=10, read_pattern_out=None, bucket_size=250):
    L = lib()
    L.bb_inspect.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_size_t]
    err = C.create_string_buffer(1024)
    if L.bb_inspect(_enc(annotated_file), top_n, _enc(read_pattern_out), bucket_size, err, 1024) != 0:
        raise BarbellError(err.value.decode())


def trim_matches(filtered_match_file, read_fastq_files, output_folder, add_labels=True, add_orientation=True, add_flank=True,
                 sort_labels=False, only_side=None, failed_out=None, write_full_header=True, skip_trim=False, flip=False, gzip=False):
    """Returns {total, trimmed, trimmed_split, failed} read counts (TrimConfig fields, src/config.rs:19-32)."""
    L = lib()
    L.bb_trim.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int32, C.c_char_p, C.POINTER(_TrimOpts), C.POINTER(C.c_uint64), C.c_char_p, C.c_size_t]
    o = _TrimOpts(int(add_labels), int(add_orientation), int(add_flank), int(sort_labels), {None: 0, "left": 1, "right": 2}[only_side],
                  int(write_full_header), int(skip_trim), int(flip), int(gzip), _enc(failed_out))
    paths = (C.c_char_p * max(1, len(read_fastq_files)))(*[_enc(p) for p in read_fastq_files])
    counts, err = (C.c_uint64 * 4)(), C.create_string_buffer(1024)
    if L.bb_trim(_enc(filtered_match_file), paths, len(read_fastq_files), _enc(output_folder), C.byref(o), counts, err, 1024) != 0:
        raise BarbellError(err.value.decode())
    return dict(total=counts[0], trimmed=counts[1], trimmed_split=counts[2], failed=counts[3])
