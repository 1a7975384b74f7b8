#!/bin/bash
# Pins the search policies S1-S6 against the UPSTREAM binary, wherever a Rust toolchain and crates.io are available
# (neither exists in the build image, so this script has never been run there; DESIGN.md section 5: "parity unpinned").
#
#   tools/ref_parity.sh /path/to/rickbeeloo-barbell-checkout [n_reads] [kit] [--policy BITS|auto] [--time]
#
# 1. builds upstream at the surveyed commit (9a2b814) with cargo,
# 2. writes the synthetic FASTQ of SURVEY.md 8(d) (tools/make_fastq.py: seeded, same reads as tests/golden and bench.py),
# 3. runs `barbell annotate -t 1` upstream (one worker thread: rows come out in input order) and this build's CLI,
# 4. diffs the two annotation.tsv byte for byte (both are in read order; within a read rows are ordered by read_start_flank).
# The choices of sassy 0.2.1 that the reference's own tests do not pin are RUN-TIME SWITCHES of this build (bb_opts.policy /
# `barbell annotate --policy BITS`, include/barbell_b200.h BB_POL_*) and of the oracle (orc_policy.flags): 1 = S1 left end of a
# plateau, 2 = S2 pattern-only before text-only, 4 = S5 last of equal minima, 8 = S6 Rc matches first, 16 / 32 = S3 round / ceil.
#   --policy BITS   run this build under that setting;   --policy auto   try all 48 settings and report the ones that match.
# A matching setting becomes the default by changing ONE constant per side (BB_POL_DEFAULT in include/barbell_b200.h and
# g_policy in oracle/barbell_oracle.c) -- no kernel is touched; the GPU == oracle suite already runs under every setting.
set -euo pipefail
REF=${1:?path to a checkout of rickbeeloo/barbell}
N=${2:-1000}
KIT=${3:-SQK-NBD114-96}
POL=0; TIME=0
shift $(( $# < 3 ? $# : 3 ))
while [ $# -gt 0 ]; do
  case "$1" in
    --policy) POL=$2; shift 2;;
    --time) TIME=1; shift;;
    *) echo "unknown argument $1"; exit 2;;
  esac
done
HERE=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
( cd "$REF" && git checkout -q 9a2b814 && cargo build --release --locked )
python "$HERE/tools/make_fastq.py" "$TMP/reads.fastq" "$N" 10000 "$KIT"
"$REF/target/release/barbell" annotate --kit "$KIT" -i "$TMP/reads.fastq" -o "$TMP/upstream.tsv" -t 1
try() {   # $1 = policy bits
  "$HERE/barbell_b200/barbell" annotate --kit "$KIT" -i "$TMP/reads.fastq" -o "$TMP/b200_$1.tsv" --policy "$1" > /dev/null
  cmp -s "$TMP/upstream.tsv" "$TMP/b200_$1.tsv"
}
if [ "$POL" = "auto" ]; then
  found=""
  for s3 in 0 16 32; do for lo in $(seq 0 15); do
    p=$(( s3 + lo ))
    if try $p; then found="$found $p"; fi
  done; done
  if [ -n "$found" ]; then echo "PARITY OK under policy setting(s):$found  ($N reads, $KIT)"; else echo "NO policy setting reproduces upstream: see the diff at --policy 0"; POL=0; fi
fi
if [ "$POL" != "auto" ]; then
  if try "$POL"; then
    echo "PARITY OK: annotation.tsv identical under --policy $POL ($(wc -l < "$TMP/b200_$POL.tsv") lines, $N reads, $KIT)"
  else
    echo "PARITY DIFFERS under --policy $POL: first differences (upstream <, b200 >):"
    diff "$TMP/upstream.tsv" "$TMP/b200_$POL.tsv" | head -40 || true
    echo "re-run with --policy auto to search the 48 settings of S1/S2/S3/S5/S6"
  fi
fi
if [ "$TIME" = 1 ]; then
  /usr/bin/env time -v "$REF/target/release/barbell" annotate --kit "$KIT" -i "$TMP/reads.fastq" -o "$TMP/upstream_mt.tsv" -t "$(nproc)" 2>&1 | egrep "Elapsed|Maximum resident" || true
fi
rm -rf "$TMP"
