#!/bin/bash
# Round-2 evidence run on one B200 (under gpurun): stage timings, bench lines of every config, ncu launch list + instruction
# counters of one step per config, ncu --set full of the three hot kernels, CLI FASTQ -> TSV throughput.  Output: gpurun_out/r2/
set -u
O=gpurun_out/r2; mkdir -p $O
python tools/stage_timings.py 100000 > $O/stage_timings_all_configs.txt 2>&1
for c in nbd rbk_k5 ald384 rbk_ext; do
  python bench.py --config $c --steps 10 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
done
for c in nbd rbk_k5 ald384 rbk_ext; do
  ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none --csv \
      --log-file $O/inst_$c.csv python tools/prof_stage.py --config $c --reads 100000 --iters 2 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"k_barcode_rows|k_flank_filter|k_flank_verify|k_read_resolve" -s 8 -c 4 \
    -o $O/ncu_full_nbd python tools/prof_stage.py --config nbd --reads 20000 --iters 3 > $O/ncu_full_nbd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_flank_scan" -s 2 -c 1 \
    -o $O/ncu_full_rbk_ext python tools/prof_stage.py --config rbk_ext --reads 20000 --iters 2 > $O/ncu_full_rbk_ext.log 2>&1
python tools/make_fastq.py /dev/shm/r.fastq 100000 > /dev/null
F=/dev/shm/r.fastq
{
  ls -la $F | awk '{print "fastq bytes", $5, "(100000 reads of 10 kb); the file is named 10 times on the command line = 1 M reads, 20 GB of FASTQ text from the page cache"}'
  nproc | awk '{print "host cores", $1}'
  for t in 4 8 16 24; do
    echo "== -t $t"; barbell_b200/barbell annotate --kit SQK-NBD114-96 -i $F $F $F $F $F $F $F $F $F $F -o /dev/shm/x.tsv -t $t --verbose 2>&1 | grep -v "complete\|Auto"
  done
  echo "== -t 16 --no-pack"; barbell_b200/barbell annotate --kit SQK-NBD114-96 -i $F $F $F $F $F $F $F $F $F $F -o /dev/shm/y.tsv -t 16 --no-pack --verbose 2>&1 | grep -v "complete\|Auto"
  echo "== -t 16 --single-reader"; barbell_b200/barbell annotate --kit SQK-NBD114-96 -i $F $F -o /dev/shm/z.tsv -t 16 --single-reader --verbose 2>&1 | grep -v "complete\|Auto"
  md5sum /dev/shm/x.tsv /dev/shm/y.tsv
  barbell_b200/barbell annotate --kit SQK-NBD114-96 -i $F -o /dev/shm/one.tsv -t 16 | tail -2 | head -1; md5sum /dev/shm/one.tsv
  s=$(date +%s%N); barbell_b200/barbell kit -k SQK-NBD114-96 -i $F -o /dev/shm/kit -t 16 | tail -3 | tr '\n' ' '; e=$(date +%s%N); echo "| kit pipeline wall_ms=$(( (e - s) / 1000000 ))"
  s=$(date +%s%N); barbell_b200/barbell annotate --kit SQK-NBD114-96 -i $F -o /dev/shm/one.tsv -t 16 > /dev/null; e=$(date +%s%N); echo "| annotate alone wall_ms=$(( (e - s) / 1000000 ))"
} > $O/cli_throughput.txt 2>&1
