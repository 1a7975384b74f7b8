"""FASTQ(.gz) reader of the `barbell` CLI (reference src/io/io.rs:5-32): plain / gzip / CRLF / multi-file / no trailing newline."""
import gzip
import os
import subprocess

import barbell_b200 as bb

EXE = os.path.join(os.path.dirname(bb.lib_path()), "barbell")


def fnv(records):
    h = 1469598103934665603
    for rid, seq in records:
        for part in (rid, seq):
            for c in part:
                h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
            h = ((h ^ 0xFF) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def stats(paths, *extra, want_reader=None):
    r = subprocess.run([EXE, "fastq-stats", "-i"] + [str(p) for p in paths] + list(extra), capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    kv = dict(x.split("=") for x in r.stdout.split())
    if want_reader:
        assert kv["reader"] == want_reader, kv
    assert kv["form"] == ("bytes" if "--no-pack" in extra else "packed")
    if "--no-pack" not in extra:
        # the default form packs every sequence line to 2 bits per base + exceptions while parsing; decoded, it must hash the same
        assert stats(paths, *extra, "--no-pack", want_reader=want_reader) == (int(kv["records"]), int(kv["bases"]), int(kv["fnv"], 16))
    return int(kv["records"]), int(kv["bases"]), int(kv["fnv"], 16)


def test_reader_variants(tmp_path):
    import random
    rnd = random.Random(3)
    recs = [(f"read{i}".encode(), bytes(rnd.choice(b"ACGTN") for _ in range(rnd.choice([0, 1, 50, 700, 70000])))) for i in range(60)]
    def text(rs, nl="\n", desc=True, final_nl=True):
        t = "".join(f"@{rid.decode()}{' runid=1 ch=2' if desc else ''}{nl}{seq.decode()}{nl}+{nl}{'I' * len(seq)}{nl}" for rid, seq in rs)
        return t if final_nl else t[:-len(nl)]
    want = (len(recs), sum(len(s) for _, s in recs), fnv(recs))
    p1 = tmp_path / "a.fastq"; p1.write_text(text(recs))
    assert stats([p1]) == want
    p2 = tmp_path / "b.fastq.gz"
    with gzip.open(p2, "wt") as f:
        f.write(text(recs))
    assert stats([p2]) == want
    p3 = tmp_path / "c.fastq"; p3.write_bytes(text(recs, nl="\r\n").encode())
    assert stats([p3]) == want
    p4 = tmp_path / "d.fastq"; p4.write_text(text(recs, desc=False, final_nl=False))
    assert stats([p4]) == want
    # two files = one collection, in order (io.rs:27-32)
    pa, pb = tmp_path / "e1.fastq", tmp_path / "e2.fastq.gz"
    pa.write_text(text(recs[:25]))
    with gzip.open(pb, "wt") as f:
        f.write(text(recs[25:]) + "\n\n")
    assert stats([pa, pb]) == want
    # blank lines between records (one or several) are skipped by every reader alike
    p6 = tmp_path / "blank.fastq"
    p6.write_text(text(recs[:10]) + "\n" + text(recs[10:20]) + "\n\n" + text(recs[20:]) + "\n\n\n")
    assert stats([p6], "--single-reader", want_reader="sequential") == want
    assert stats([p6], "--chunk-kb", "2", "-t", "4", want_reader="parallel") == want
    # errors: truncated record, missing file -> message, non-zero exit of this debugging subcommand
    p5 = tmp_path / "bad.fastq"; p5.write_text("@r1\nACGT\n+\n")
    r = subprocess.run([EXE, "fastq-stats", "-i", str(p5)], capture_output=True, text=True)
    assert r.returncode != 0 and "truncated" in r.stdout
    r = subprocess.run([EXE, "fastq-stats", "-i", str(tmp_path / "nope.fastq")], capture_output=True, text=True)
    assert r.returncode != 0 and "Failed to open" in r.stdout


def test_parallel_chunked_reader_equals_sequential(tmp_path):
    """Plain files are cut into chunks at record boundaries and parsed by several threads; batches must come out in input order
    whatever the chunk size (records spanning chunks, chunks without a record start, '@' as the first quality character)."""
    import random
    rnd = random.Random(11)
    recs = [(f"r{i}/x".encode(), bytes(rnd.choice(b"ACGTN") for _ in range(rnd.choice([0, 1, 3, 50, 700, 5000, 70000])))) for i in range(150)]

    def text(rs, nl="\n", final_nl=True):
        out = []
        for i, (rid, seq) in enumerate(rs):
            qual = "".join(rnd.choice("@+I5") for _ in seq)          # quality lines that LOOK like headers / separators
            out.append(f"@{rid.decode()} ch={i}{nl}{seq.decode()}{nl}+{nl}{qual}{nl}")
        t = "".join(out)
        return t if final_nl else t[:-len(nl)]
    want = (len(recs), sum(len(s) for _, s in recs), fnv(recs))
    pa, pb, pc = tmp_path / "a.fastq", tmp_path / "b.fastq", tmp_path / "c.fastq"
    pa.write_text(text(recs[:70])); pb.write_bytes(text(recs[70:110], nl="\r\n").encode()); pc.write_text(text(recs[110:], final_nl=False))
    assert stats([pa, pb, pc], "--single-reader", want_reader="sequential") == want
    for chunk_kb, threads in [(1, 8), (3, 4), (64, 8), (300, 3), (100000, 8)]:
        assert stats([pa, pb, pc], "--chunk-kb", str(chunk_kb), "-t", str(threads), want_reader="parallel") == want, (chunk_kb, threads)
    # -t 1 falls back to the single reader thread; several files with gzip among them are inflated and parsed one file per worker
    assert stats([pa, pb, pc], "-t", "1", want_reader="sequential") == want
    zs = []
    for k in range(7):
        pz = tmp_path / f"z{k}.fastq.gz"
        with gzip.open(pz, "wt") as f:
            f.write(text(recs[20 * k:20 * (k + 1)]))
        zs.append(pz)
    pe = tmp_path / "zz_empty.fastq"; pe.write_text("")
    want_z = (140, sum(len(s) for _, s in recs[:140]), fnv(recs[:140]))
    assert stats(zs[:3] + [pe] + zs[3:], "-t", "4", want_reader="multifile") == want_z
    assert stats(zs[:3] + [pe] + zs[3:], "-t", "3", "--batch-mb", "1", want_reader="multifile") == want_z      # many small batches
    assert stats(zs[:3] + [pe] + zs[3:], "--single-reader", want_reader="sequential") == want_z
    stats([zs[0]], want_reader="sequential")                                                                     # one gzip file: one zlib stream
    badz = tmp_path / "bad.fastq.gz"
    with gzip.open(badz, "wt") as f:
        f.write("@r1\nACGT\n+\nII\n")
    r = subprocess.run([EXE, "fastq-stats", "-i", str(zs[0]), str(badz), str(zs[1]), "-t", "4"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "quality length" in r.stdout, r.stdout
    # errors surface from the parser threads too
    bad = tmp_path / "bad.fastq"; bad.write_text(text(recs[:20]) + "@r1\nACGT\n+\nII\n" + text(recs[20:40]))
    r = subprocess.run([EXE, "fastq-stats", "-i", str(bad), "--chunk-kb", "2", "-t", "4"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and ("quality length" in r.stdout or "malformed" in r.stdout), r.stdout
    empty = tmp_path / "empty.fastq"; empty.write_text("")
    assert stats([empty, pa], "--chunk-kb", "4")[0] == 70


def test_many_tiny_reads_grow_the_offset_arrays(tmp_path):
    """The read-offset arrays of a batch slot start sized for reads of >= 256 bases and grow: a chunk (parallel reader) / a batch
    (sequential reader) with several hundred thousand 0..2-base reads must come through unchanged."""
    import random
    rnd = random.Random(5)
    recs = [(b"r%d" % i, bytes(rnd.choice(b"ACGTN") for _ in range(i % 3))) for i in range(400_000)]
    p = tmp_path / "tiny.fastq"
    p.write_bytes(b"".join(b"@%s\n%s\n+\n%s\n" % (rid, seq, b"I" * len(seq)) for rid, seq in recs))
    want = (len(recs), sum(len(s) for _, s in recs), fnv(recs))
    assert stats([p], "--chunk-kb", "8192", "-t", "2", want_reader="parallel") == want
    assert stats([p], "--single-reader", "--batch-mb", "1", want_reader="sequential") == want
