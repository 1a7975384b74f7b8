set -u
TMP=$(mktemp -d); mkdir -p gpurun_out/cli
python tools/make_fastq.py $TMP/reads.fastq 100000 > /dev/null
for i in 1 2 3; do
  s=$(date +%s%N)
  barbell_b200/barbell annotate --kit SQK-NBD114-96 -i $TMP/reads.fastq -o $TMP/out.tsv -t 8 --verbose 2>&1 | grep -a "timing\|Total"
  e=$(date +%s%N); echo "wall_ms=$(( (e - s) / 1000000 ))"
done
s=$(date +%s%N); python -c "
import torch,time
t=time.time(); torch.cuda.init(); torch.zeros(1,device='cuda'); print('torch ctx', time.time()-t)
t=time.time(); x=torch.empty(1<<30,dtype=torch.uint8).pin_memory(); print('pin 1GB', time.time()-t)
"
rm -rf $TMP
