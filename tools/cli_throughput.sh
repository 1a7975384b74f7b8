#!/bin/bash
# CLI FASTQ -> annotation.tsv throughput on the GPU box (under gpurun): thread sweep, copy forms, parser alone, kit pipeline wall.
set -u
O=gpurun_out/cli; mkdir -p $O
python tools/make_fastq.py /dev/shm/r.fastq 100000 > /dev/null
F=/dev/shm/r.fastq
{
  lscpu | grep -i "model name\|^CPU(s)\|Thread(s) per core\|Socket" | tr -s ' '
  ls -la $F | awk '{print "fastq bytes", $5, "(100000 reads of 10 kb); the file is named 10 times on the command line = 1 M reads, 20 GB of FASTQ text from the page cache"}'
  for t in 4 8 12 16 24; do
    echo "== -t $t"; barbell_b200/barbell annotate --kit SQK-NBD114-96 -i $F $F $F $F $F $F $F $F $F $F -o /dev/shm/x.tsv -t $t --verbose 2>&1 | grep -v "complete\|Auto"
  done
  for t in 8 16; do
    echo "== fastq-stats --count-only -t $t (the parsers alone, no GPU)"; barbell_b200/barbell fastq-stats -i $F $F $F $F $F $F $F $F $F $F -t $t --count-only 2>&1 | grep timing
  done
  echo "== -t 16 --batch-mb 256"; barbell_b200/barbell annotate --kit SQK-NBD114-96 -i $F $F $F $F $F $F $F $F $F $F -o /dev/shm/x.tsv -t 16 --batch-mb 256 --verbose 2>&1 | grep -v "complete\|Auto"
  echo "== -t 16 --no-pack"; barbell_b200/barbell annotate --kit SQK-NBD114-96 -i $F $F $F $F $F $F $F $F $F $F -o /dev/shm/y.tsv -t 16 --no-pack --verbose 2>&1 | grep -v "complete\|Auto"
  md5sum /dev/shm/x.tsv /dev/shm/y.tsv
  barbell_b200/barbell annotate --kit SQK-NBD114-96 -i $F -o /dev/shm/one.tsv -t 16 | tail -2 | head -1; md5sum /dev/shm/one.tsv
  s=$(date +%s%N); barbell_b200/barbell kit -k SQK-NBD114-96 -i $F -o /dev/shm/kit -t 16 | tail -3 | tr '\n' ' '; e=$(date +%s%N); echo "| kit pipeline wall_ms=$(( (e - s) / 1000000 ))"
  s=$(date +%s%N); barbell_b200/barbell annotate --kit SQK-NBD114-96 -i $F -o /dev/shm/one.tsv -t 16 > /dev/null; e=$(date +%s%N); echo "| annotate alone wall_ms=$(( (e - s) / 1000000 ))"
} > $O/cli_throughput.txt 2>&1
cat $O/cli_throughput.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-modes > $O/bench_nbd_e2e_modes.json 2> $O/bench_nbd.err
python - <<'P'
import json
j = json.loads(open('gpurun_out/cli/bench_nbd_e2e_modes.json').read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e'], 'packed', j.get('e2e_packed', {}).get('value'), 'fastq', j.get('e2e_fastq', {}).get('value'))
P
