#!/bin/bash
# One GPU session: parity tests, stage timings, bench line, ncu launch list and full captures of the hot kernels.
# Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh <tag> [what...]   what: tests stages bench launches ncu
set -u
TAG=${1:-run}; shift || true
WHAT=${*:-tests stages bench launches ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
for w in $WHAT; do
  case $w in
    tests) timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log ;;
    stages) timeout 600 python tools/stage_timings.py 100000 > $OUT/stages.log 2>&1; cat $OUT/stages.log ;;
    bench) timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json; tail -3 $OUT/bench.err ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --reads 20000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/launches.log 2>&1 ;;
    ncu) timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_barcode_rows|k_flank_filter|k_flank_verify|k_read_resolve' -s 12 -c 3 -o $OUT/full -f python bench.py --reads 20000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu.log 2>&1; tail -3 $OUT/ncu.log ;;
  esac
done
ls -la $OUT
