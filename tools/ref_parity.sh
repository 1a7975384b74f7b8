#!/bin/bash
# Pins the oracle's sassy policies S1-S7 against the UPSTREAM binary, wherever a Rust toolchain and crates.io are available
# (neither exists in the build image, so this script has never been run there; DESIGN.md section 5: "parity unpinned").
#
#   tools/ref_parity.sh /path/to/rickbeeloo-barbell-checkout [n_reads] [kit]
#
# 1. builds upstream at the surveyed commit (9a2b814) with cargo,
# 2. writes the synthetic FASTQ of SURVEY.md 8(d) (tools/make_fastq.py: seeded, same reads as tests/golden and bench.py),
# 3. runs `barbell annotate -t 1` upstream (one worker thread: rows come out in input order) and this build's CLI,
# 4. diffs the two annotation.tsv byte for byte (both are sorted by read order; within a read rows are ordered by
#    read_start_flank in both).  With --time it also prints upstream's wall time at -t $(nproc): the true CPU baseline.
set -euo pipefail
REF=${1:?path to a checkout of rickbeeloo/barbell}
N=${2:-1000}
KIT=${3:-SQK-NBD114-96}
HERE=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
( cd "$REF" && git checkout -q 9a2b814 && cargo build --release --locked )
python "$HERE/tools/make_fastq.py" "$TMP/reads.fastq" "$N" 10000 "$KIT"
"$REF/target/release/barbell" annotate --kit "$KIT" -i "$TMP/reads.fastq" -o "$TMP/upstream.tsv" -t 1
"$HERE/barbell_b200/barbell" annotate --kit "$KIT" -i "$TMP/reads.fastq" -o "$TMP/b200.tsv"
if cmp -s "$TMP/upstream.tsv" "$TMP/b200.tsv"; then
  echo "PARITY OK: annotation.tsv identical ($(wc -l < "$TMP/b200.tsv") lines, $N reads, $KIT)"
else
  echo "PARITY DIFFERS: first differences (upstream <, b200 >):"
  diff "$TMP/upstream.tsv" "$TMP/b200.tsv" | head -40
  echo "flip the matching policy knob in oracle/barbell_oracle.c (S1 reporting rule, S2 traceback tie-break, S3 overhang rounding,"
  echo "S4 Rc path orientation) and in the kernels (barbell_b200/csrc/kernels.cuh, barcode_lane.cuh), then re-run the test-suite."
fi
if [ "${4:-}" = "--time" ]; then
  /usr/bin/env time -v "$REF/target/release/barbell" annotate --kit "$KIT" -i "$TMP/reads.fastq" -o "$TMP/upstream_mt.tsv" -t "$(nproc)" 2>&1 | egrep "Elapsed|Maximum resident" || true
fi
rm -rf "$TMP"
