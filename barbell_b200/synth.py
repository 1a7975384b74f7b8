"""Seeded synthetic nanopore-like reads with implanted barcode tags.

The read groups follow the reference simulator's recipe (benchmarks/src/simulations/sim_data.rs:163-401):
  I    no tag                                   5 %
  II   full tag at the 5' end                  60 %
  III  tag with 1-20 outer bases trimmed off   10 %
  IV   two tags 10 bases apart                  5 %
  V    tag at the 5' end + tag mid-read         5 %
  VI   tag at the 5' end + reverse-complemented tag at the 3' end   15 %
Every implanted tag is mutated per base with p = 0.06 (substitution / insertion / deletion, equal shares); half of the
reads are reverse-complemented as a whole; 0.1 % of all bases are replaced by N.  Only numpy is used, so the same
bytes are produced on the build container and on the GPU box.
"""
import numpy as np

_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    _COMP[a] = b
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
SEED0 = 0xBA5BE11


def revcomp(x: np.ndarray) -> np.ndarray:
    return _COMP[x[::-1]]


def full_tags(group: dict):
    """Full query sequences (front + barcode + rear) of a group dict from GroupSet.as_dicts()."""
    flank = group["flank"]
    b0, b1 = group["bar_region"]
    p0 = group["pad_region"][0]
    out = []
    for pb in group["barcodes"]:
        core = pb[b0 - p0:b0 - p0 + (b1 - b0 + 1)]
        out.append(np.frombuffer(flank[:b0] + core + flank[b1 + 1:], dtype=np.uint8))
    return out


def mutate(rng, seq: np.ndarray, p: float) -> np.ndarray:
    if p <= 0:
        return seq
    r = rng.random(len(seq))
    out = []
    for i, c in enumerate(seq):
        if r[i] < p / 3:
            out.append(_ACGT[(int(np.searchsorted(_ACGT, c)) + 1 + rng.integers(0, 3)) % 4])   # substitution
        elif r[i] < 2 * p / 3:
            out.append(c)
            out.append(_ACGT[rng.integers(0, 4)])                                                # insertion
        elif r[i] < p:
            continue                                                                             # deletion
        else:
            out.append(c)
    return np.array(out, dtype=np.uint8)


def make_reads(groups, n_reads, read_len=10000, seed=SEED0, p_mut=0.06, rc_frac=0.5, n_frac=0.001, plain_frac=None):
    """Returns (bases uint8[total], offsets uint64[n_reads+1], truth list).  read_len: int or (lo, hi) inclusive.
    groups: list of dicts (GroupSet.as_dicts()).  truth[r] = (group label string, list of implanted barcode indices)."""
    rng = np.random.default_rng(seed)
    if isinstance(read_len, (tuple, list)):
        lens = rng.integers(read_len[0], read_len[1] + 1, size=n_reads)
    else:
        lens = np.full(n_reads, int(read_len), dtype=np.int64)
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens).astype(np.uint64)
    total = int(offsets[-1])
    bases = _ACGT[rng.integers(0, 4, size=total, dtype=np.uint8)]
    tags = [full_tags(g) for g in groups]
    f_groups = [i for i, g in enumerate(groups) if g["match_type"] == 0] or [0]
    r_groups = [i for i, g in enumerate(groups) if g["match_type"] == 1]
    cum = np.cumsum([0.05, 0.60, 0.10, 0.05, 0.05, 0.15])
    kinds = np.searchsorted(cum, rng.random(n_reads), side="right").clip(0, 5)
    truth = []
    for r in range(n_reads):
        L = int(lens[r])
        s = int(offsets[r])
        read = bases[s:s + L]
        kind = int(kinds[r])
        gi = f_groups[int(rng.integers(0, len(f_groups)))]
        bi = int(rng.integers(0, len(tags[gi])))
        placed = []

        def put(pos, seq):
            seq = seq[:max(0, L - pos)]
            if pos >= 0 and len(seq):
                read[pos:pos + len(seq)] = seq

        if kind >= 1:
            t = mutate(rng, tags[gi][bi], p_mut)
            if kind == 2:
                t = t[int(rng.integers(1, 21)):]
            put(0, t)
            placed.append((gi, bi))
            if kind == 3:
                b2 = int(rng.integers(0, len(tags[gi])))
                put(len(t) + 10, mutate(rng, tags[gi][b2], p_mut))
                placed.append((gi, b2))
            elif kind == 4:
                b2 = int(rng.integers(0, len(tags[gi])))
                put(L // 2, mutate(rng, tags[gi][b2], p_mut))
                placed.append((gi, b2))
            elif kind == 5:
                if r_groups:
                    g2 = r_groups[int(rng.integers(0, len(r_groups)))]
                    b2 = int(rng.integers(0, len(tags[g2])))
                    t2 = mutate(rng, tags[g2][b2], p_mut)
                else:
                    g2, b2 = gi, bi
                    t2 = revcomp(mutate(rng, tags[gi][bi], p_mut))
                put(max(0, L - len(t2)), t2)
                placed.append((g2, b2))
        if rng.random() < rc_frac:
            read[:] = revcomp(read.copy())
        truth.append((kind, placed))
    if n_frac > 0:
        k = rng.binomial(total, n_frac)
        bases[rng.integers(0, total, size=k)] = ord("N")
    return bases, offsets, truth


def _fasta_seqs(path):
    seqs = []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            seqs.append(bytearray())
        elif line and seqs:
            seqs[-1] += line.upper().encode()
    return [bytes(x) for x in seqs]


def _common_prefix_len(seqs):
    n = min(len(x) for x in seqs)
    k = 0
    while k < n and all(x[k] == seqs[0][k] for x in seqs):
        k += 1
    return k


def dual_end_panel_specs(left_fasta, right_fasta, n_bar=384, seed=26):
    """BASELINE configs[3]: a custom dual-end panel -- n_bar random 24-mers wrapped in the flanks of the reference's example
    tags (examples/ald_left.fasta / ald_right.fasta, copied to tests/golden/), one Ftag and one Rtag group.  First / last barcode
    base are varied so that the common prefix / suffix stay on the flanks.  Returns [(seqs, labels, match_type), ...] for
    GroupSet.from_seqs (product) or the oracle's group builder."""
    rnd = np.random.default_rng(seed)
    specs = []
    for path, tag, ty in ((left_fasta, "L", 0), (right_fasta, "R", 1)):
        src = _fasta_seqs(path)
        pre = _common_prefix_len(src)
        suf = _common_prefix_len([x[::-1] for x in src])
        front, rear = src[0][:pre], src[0][len(src[0]) - suf:]
        seqs = []
        for i in range(n_bar):
            core = bytes(rnd.choice(_ACGT, 24))
            seqs.append(front + b"ACGT"[i % 4:i % 4 + 1] + core[1:-1] + b"ACGT"[(i // 4) % 4:(i // 4) % 4 + 1] + rear)
        specs.append((seqs, [f"{tag}{i}" for i in range(n_bar)], ty))
    return specs


def write_fastq(path, bases, offsets, prefix="read_"):
    """FASTQ text of a batch (quality 'I' throughout)."""
    n = len(offsets) - 1
    qual = b"I" * int(np.diff(offsets.astype(np.int64)).max(initial=0))
    with open(path, "wb", buffering=1 << 24) as f:
        for i in range(n):
            sq = bases[int(offsets[i]):int(offsets[i + 1])].tobytes()
            f.write(b"@%s%d ch=1\n" % (prefix.encode(), i)); f.write(sq); f.write(b"\n+\n"); f.write(qual[:len(sq)]); f.write(b"\n")
