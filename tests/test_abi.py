"""The C-ABI library loads and exports every symbol include/barbell_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

import barbell_b200 as bb
from barbell_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "barbell_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    L = bb.lib()
    syms = declared_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/barbell_b200.h but not exported"
    assert set(api.EXPORTS) == set(syms)
    assert L.bb_abi_version() == 2


def test_row_layout_matches_oracle_and_header():
    import oracle_lib as O
    assert bb.ROW_DTYPE == O.ROW_DTYPE and bb.ROW_DTYPE.itemsize == 88


def test_no_cpu_fallback():
    """Without a CUDA device bb_create must fail loudly (the product path never routes through the oracle)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    gs = bb.GroupSet.from_kit("SQK-NBD114-96")
    with pytest.raises(bb.BarbellError, match="no CUDA device|CPU"):
        bb.Annotator(gs)


def test_product_does_not_reference_oracle():
    """No file of the product tree mentions the oracle library or imports tests/oracle_lib."""
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "barbell_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if "oracle_lib" in txt or "libbarbell_oracle" in txt or "barbell_oracle.h" in txt:
                    bad.append(fn)
    assert not bad, bad
    out = os.popen(f"ldd {bb.lib_path()}").read()
    assert "oracle" not in out


def test_pack_nibbles_matches_numpy():
    """bb_pack_nibbles (AVX2 + thread pool) against a numpy restatement, all 256 byte values, odd lengths."""
    import numpy as np
    code = np.zeros(256, np.uint8)
    for ch, v in zip(b"ACGTURYSWKMBDHVN", [1, 2, 4, 8, 8, 5, 10, 6, 9, 12, 3, 14, 13, 11, 7, 15]):
        code[ch] = v; code[ch | 0x20] = v
    rnd = np.random.default_rng(9)
    for n in (0, 1, 63, 64, 65, 1000, 5_000_001):
        src = rnd.integers(0, 256, n, dtype=np.uint8)
        c = code[src]
        if n & 1:
            c = np.concatenate([c, [0]]).astype(np.uint8)
        want = (c[0::2] | (c[1::2] << 4)).astype(np.uint8)
        dst = np.zeros((n + 1) // 2 + 1, np.uint8)
        assert bb.lib().bb_pack_nibbles(src.ctypes.data, n, dst.ctypes.data) == 0
        assert (dst[:(n + 1) // 2] == want).all()


def test_pack_crumbs_matches_numpy():
    """bb_pack_crumbs (AVX2 + thread pool, exception blocks) against a numpy restatement: mostly-ACGT input with every other byte
    value sprinkled in, lengths around the 4- and 128-base steps and past the 4 M-base work item; and the overflow return."""
    import ctypes as C
    import numpy as np
    code = np.zeros(256, np.uint8)
    for ch, v in zip(b"ACGTURYSWKMBDHVN", [1, 2, 4, 8, 8, 5, 10, 6, 9, 12, 3, 14, 13, 11, 7, 15]):
        code[ch] = v; code[ch | 0x20] = v
    crumb_of = np.full(16, 255, np.uint8); crumb_of[[1, 2, 4, 8]] = [0, 1, 2, 3]
    rnd = np.random.default_rng(10)
    for n, frac in ((0, 0), (1, 0.5), (3, 0.5), (127, 0.1), (128, 0.1), (131, 0.1), (1000, 1.0), (9_000_003, 0.004), (9_000_003, 0.0)):
        src = rnd.choice(np.frombuffer(b"ACGTacgtUu", np.uint8), n)
        odd = rnd.random(n) < frac
        src[odd] = rnd.integers(0, 256, int(odd.sum()), dtype=np.uint8)
        sets = code[src]
        cr = crumb_of[sets]
        exc_want = np.sort((np.flatnonzero(cr == 255).astype(np.uint64) << np.uint64(4)) | sets[cr == 255].astype(np.uint64))
        cr = np.where(cr == 255, 0, cr).astype(np.uint8)
        cr = np.concatenate([cr, np.zeros((-n) % 4, np.uint8)])
        want = cr[0::4] | (cr[1::4] << 2) | (cr[2::4] << 4) | (cr[3::4] << 6)
        cap = len(exc_want) + 256 * (n // (4 << 20) + 2)
        dst = np.zeros((n + 3) // 4 + 1, np.uint8)
        exc = np.zeros(cap, np.uint64)
        n_exc = C.c_uint64(0)
        assert bb.lib().bb_pack_crumbs(src.ctypes.data, n, dst.ctypes.data, exc.ctypes.data, cap, C.byref(n_exc)) == 0
        assert (dst[:(n + 3) // 4] == want).all()
        got = exc[:n_exc.value]
        assert n_exc.value % 256 == 0 and n_exc.value <= cap
        got = np.sort(got[got != np.uint64(0xFFFFFFFFFFFFFFFF)])
        assert len(got) == len(exc_want) and (got == exc_want).all()
    # too small an exception list is reported, not silently truncated
    src = np.frombuffer(b"N" * 4096, np.uint8).copy()
    dst = np.zeros(1025, np.uint8); exc = np.zeros(512, np.uint64); n_exc = C.c_uint64(0)
    assert bb.lib().bb_pack_crumbs(src.ctypes.data, 4096, dst.ctypes.data, exc.ctypes.data, 512, C.byref(n_exc)) == -3   # BB_ERR_OVERFLOW
    # ... also with every packing thread hitting the end of the list at once: nothing is written past exc_cap
    n = 9_000_000
    src = np.full(n, ord("N"), np.uint8)
    dst = np.zeros(n // 4 + 8, np.uint8)
    cap = 4096
    exc = np.full(cap + 4096, 0x5A5A5A5A5A5A5A5A, np.uint64)
    assert bb.lib().bb_pack_crumbs(src.ctypes.data, n, dst.ctypes.data, exc.ctypes.data, cap, C.byref(n_exc)) == -3
    assert (exc[cap:] == np.uint64(0x5A5A5A5A5A5A5A5A)).all() and n_exc.value <= cap


def test_pack_crumbs_append_builds_the_same_stream_read_by_read():
    """bb_pack_crumbs_append (what the FASTQ reader of the CLI calls per sequence line): reads of every length mod 4 / 32 / 128
    appended one after the other must give the byte stream and the exception entries of one bb_pack_crumbs call over the
    concatenation; a full exception list is reported."""
    import ctypes as C
    import numpy as np
    rnd = np.random.default_rng(12)
    lens = [0, 1, 2, 3, 5, 31, 32, 33, 127, 128, 129, 130, 131, 255, 1000, 4097] + [int(x) for x in rnd.integers(0, 700, 200)]
    reads = []
    for n in lens:
        r = rnd.choice(np.frombuffer(b"ACGTacgt", np.uint8), n)
        odd = rnd.random(n) < 0.02
        r[odd] = rnd.choice(np.frombuffer(b"NnRYKM-x*", np.uint8), int(odd.sum()))
        reads.append(r)
    allb = np.concatenate(reads)
    n = len(allb)
    want = np.zeros((n + 3) // 4 + 1, np.uint8); exc_w = np.zeros(n // 8 + 1024, np.uint64); nw = C.c_uint64(0)
    assert bb.lib().bb_pack_crumbs(allb.ctypes.data, n, want.ctypes.data, exc_w.ctypes.data, len(exc_w), C.byref(nw)) == 0
    exc_w = exc_w[:nw.value]; exc_w = np.sort(exc_w[exc_w != np.uint64(0xFFFFFFFFFFFFFFFF)])
    dst = np.zeros((n + 3) // 4 + 64, np.uint8); exc = np.full(len(exc_w) + 8, 0x5A5A5A5A5A5A5A5A, np.uint64)
    pos, ne = C.c_uint64(0), C.c_uint64(0)
    for r in reads:
        r = np.ascontiguousarray(r)
        assert bb.lib().bb_pack_crumbs_append(r.ctypes.data if len(r) else None, len(r), dst.ctypes.data, C.byref(pos), exc.ctypes.data, len(exc_w), C.byref(ne)) == 0
    assert pos.value == n and ne.value == len(exc_w)
    assert (dst[:(n + 3) // 4] == want[:(n + 3) // 4]).all() and (dst[(n + 3) // 4:] == 0).all()
    assert (exc[:ne.value] == exc_w).all()                      # appended in stream order
    assert (exc[len(exc_w):] == np.uint64(0x5A5A5A5A5A5A5A5A)).all()
    r = np.frombuffer(b"ACGNNNNT", np.uint8).copy()
    pos, ne = C.c_uint64(0), C.c_uint64(0)
    assert bb.lib().bb_pack_crumbs_append(r.ctypes.data, 8, dst.ctypes.data, C.byref(pos), exc.ctypes.data, 2, C.byref(ne)) == -3


def test_pack_crumbs_append_line_equals_append_of_the_line():
    """bb_pack_crumbs_append_line (one pass: find the end of the line while packing it) against memchr + bb_pack_crumbs_append, for
    lines whose end falls everywhere relative to the 128-base step and the 4-base byte, LF and CRLF, without a terminator, with
    text behind the line (a quality string of the same length, as in a FASTQ record) and with a short readable window."""
    import ctypes as C
    import numpy as np
    rnd = np.random.default_rng(13)
    L = bb.lib()
    lens = list(range(0, 12)) + [31, 32, 33, 63, 64, 65, 125, 126, 127, 128, 129, 130, 131, 132, 255, 256, 257, 383, 384, 385, 1000, 4099]
    for start in (0, 1, 2, 3, 4, 130):
        for term in (b"\n", b"\r\n", b""):
            for n in lens:
                seq = rnd.choice(np.frombuffer(b"ACGTacgtU", np.uint8), n)
                odd = rnd.random(n) < 0.03
                seq[odd] = rnd.choice(np.frombuffer(b"NnRY-x*\r", np.uint8), int(odd.sum()))
                if n and seq[-1] == 13:
                    seq[-1] = ord("A")                            # (a '\r' in front of the terminator would belong to it)
                behind = b"" if not term else (b"+\n" + b"I" * n + b"\n@next\nACGT\n")
                buf = np.frombuffer(seq.tobytes() + term + behind, np.uint8).copy()
                for window in (len(buf), n + len(term), max(0, n - 5)):
                    window = min(window, len(buf))
                    want_len = min(n, window) if window < n + len(term) or not term else n
                    want_found = 1 if (term and window >= n + len(term)) else 0
                    if term == b"\r\n" and window == n + 1:
                        want_len, want_found = n + 1, 0          # the '\r' alone is a (non-base) byte of the window
                    if b"\n" in buf[:want_len].tobytes():
                        continue
                    cap = 4096
                    da = np.zeros(2048, np.uint8); ea = np.zeros(cap, np.uint64); pa, na = C.c_uint64(start), C.c_uint64(0)
                    db = np.zeros(2048, np.uint8); eb = np.zeros(cap, np.uint64); pb, nb = C.c_uint64(start), C.c_uint64(0)
                    assert L.bb_pack_crumbs_append(buf.ctypes.data, want_len, da.ctypes.data, C.byref(pa), ea.ctypes.data, cap, C.byref(na)) == 0
                    ll, found = C.c_uint64(99), C.c_int(9)
                    assert L.bb_pack_crumbs_append_line(buf.ctypes.data, window, db.ctypes.data, C.byref(pb), eb.ctypes.data, cap, C.byref(nb),
                                                        C.byref(ll), C.byref(found)) == 0
                    key = (start, term, n, window)
                    assert (ll.value, found.value) == (want_len, want_found), key
                    assert pb.value == pa.value == start + want_len and nb.value == na.value, key
                    assert (eb[:nb.value] == ea[:na.value]).all(), key
                    assert (db == da).all(), key                  # ... including zeros behind the end of the stream
    # a full exception list is reported
    buf = np.frombuffer(b"NNNNNNNNNN\n", np.uint8).copy()
    d = np.zeros(64, np.uint8); e = np.zeros(4, np.uint64); p, ne = C.c_uint64(0), C.c_uint64(0)
    ll, found = C.c_uint64(0), C.c_int(0)
    assert L.bb_pack_crumbs_append_line(buf.ctypes.data, len(buf), d.ctypes.data, C.byref(p), e.ctypes.data, 4, C.byref(ne), C.byref(ll), C.byref(found)) == -3
