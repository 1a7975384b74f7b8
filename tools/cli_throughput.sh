#!/bin/bash
# FASTQ -> annotation.tsv wall time of the CLI on synthetic 10 kb reads (configs[1] shape), sequential vs chunk-parallel reader.
# usage: bash tools/cli_throughput.sh [n_reads] ; writes gpurun_out/cli/throughput.txt
set -u
N=${1:-100000}
OUT=gpurun_out/cli; mkdir -p $OUT; TMP=$(mktemp -d)
python tools/make_fastq.py $TMP/reads.fastq $N > /dev/null
ls -la $TMP/reads.fastq | awk '{print "fastq bytes", $5}' | tee $OUT/throughput.txt
for mode in "--single-reader" "-t 4" "-t 8" "-t 8"; do
  s=$(date +%s%N)
  barbell_b200/barbell annotate --kit SQK-NBD114-96 -i $TMP/reads.fastq -o $TMP/out.tsv $mode | tail -1 | tee -a $OUT/throughput.txt
  e=$(date +%s%N)
  echo "| mode=[$mode] wall_ms=$(( (e - s) / 1000000 )) reads=$N" | tee -a $OUT/throughput.txt
  md5sum $TMP/out.tsv | tee -a $OUT/throughput.txt
done
# the usual shape of a run: many small gzip files (one zlib stream each): one reader thread vs one file per worker
python - "$TMP" <<'PY'
import gzip, sys
tmp = sys.argv[1]
lines = open(tmp + "/reads.fastq", "rb").read(16 * 20000 * 4 * 5200).split(b"\n")   # first ~16k reads
recs = [b"\n".join(lines[i:i + 4]) + b"\n" for i in range(0, len(lines) - 4, 4)]
per = 1000
for k in range(len(recs) // per):
    with gzip.open(f"{tmp}/part{k:03d}.fastq.gz", "wb", compresslevel=4) as f:
        f.write(b"".join(recs[k * per:(k + 1) * per]))
print("gz parts:", len(recs) // per, "reads:", (len(recs) // per) * per)
PY
for mode in "--single-reader" "-t 8"; do
  s=$(date +%s%N)
  barbell_b200/barbell annotate --kit SQK-NBD114-96 -i $TMP/part*.fastq.gz -o $TMP/outz.tsv $mode | tail -2 | tr '\n' ' '
  e=$(date +%s%N)
  echo "| gzip parts mode=[$mode] wall_ms=$(( (e - s) / 1000000 ))" | tee -a $OUT/throughput.txt
  md5sum $TMP/outz.tsv | tee -a $OUT/throughput.txt
done
s=$(date +%s%N); barbell_b200/barbell kit -k SQK-NBD114-96 -i $TMP/reads.fastq -o $TMP/kit -t 8 | tail -3 | tr '\n' ' '; e=$(date +%s%N)
echo "| kit pipeline wall_ms=$(( (e - s) / 1000000 ))" | tee -a $OUT/throughput.txt
rm -rf $TMP
