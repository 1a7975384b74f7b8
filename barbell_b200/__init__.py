"""barbell_b200 -- B200-native `annotate` hot path of rickbeeloo/barbell behind a C ABI.

The product is `libbarbell_b200.so` (CUDA kernels for sm_100a + C++ host code, see include/barbell_b200.h);
this package is the thin ctypes mirror used by the tests, bench.py and __graft_entry__.py.
There is no CPU fallback: importing works anywhere, creating an Annotator needs a B200.
"""
from .api import (Annotator, BarbellError, GroupSet, ROW_DTYPE, MATCH_TYPE_NAMES, STRAND_NAMES, lib, lib_path,
                  rows_to_tsv, edit_cut_off)

__all__ = ["Annotator", "BarbellError", "GroupSet", "ROW_DTYPE", "MATCH_TYPE_NAMES", "STRAND_NAMES", "lib", "lib_path",
           "rows_to_tsv", "edit_cut_off"]
