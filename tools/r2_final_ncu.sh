#!/bin/bash
# ncu evidence of the final round-2 build on one B200: instruction counters of one step per config, launch list of the bench command,
# full capture of the hot kernels.  Output: gpurun_out/final_ncu/
set -u
O=gpurun_out/final_ncu; mkdir -p $O
for c in nbd rbk_k5 ald384 rbk_ext; do
  timeout 900 ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none --csv \
      --log-file $O/inst_$c.csv python tools/prof_stage.py --config $c --reads 100000 --iters 2 > /dev/null 2>&1
  echo "inst $c exit $?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launch_list_nbd.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-e2e-fastq > $O/launch_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_barcode_rows|k_flank_filter|k_flank_verify|k_read_resolve" -s 8 -c 4 \
    -o $O/ncu_full_nbd -f python tools/prof_stage.py --config nbd --reads 20000 --iters 3 > $O/ncu_full_nbd.log 2>&1
python tools/ncu_summary.py $O/ncu_full_nbd.ncu-rep > $O/ncu_full_kernels.txt 2>&1
head -12 $O/ncu_full_kernels.txt
