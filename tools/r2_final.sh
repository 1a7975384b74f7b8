#!/bin/bash
# Final evidence pass of a round: full GPU suite, then the bench line of every configuration (and the CPU arm of the headline one).
set -u
OUT=gpurun_out/final; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_nbd.json 2> $OUT/bench_nbd.err; tail -c 600 $OUT/bench_nbd.json
for c in rbk_k5 ald384 rbk_ext nbd_ext; do
  timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --config $c > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo "$c exit $?"; head -c 300 $OUT/bench_$c.json; echo
done
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $OUT/bench_nbd_reference_arm.json 2> $OUT/ref.err; cat $OUT/bench_nbd_reference_arm.json | head -c 400
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
