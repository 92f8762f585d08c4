#!/bin/bash
# N-GPU: frame next to / behind the interior.  usage: bash tools/gpu_r01t.sh <tag> <N>
TAG=$1; N=$2
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/xc_nccl_check.py > $OUT/xc_check.log 2>&1; echo "rc=$?" >> $OUT/xc_check.log; tail -2 $OUT/xc_check.log
for mode in "" "--frame-serial" "--no-overlap"; do
timeout 600 $TR bench.py --gpus $N --steps 12 --warmup 3 --no-cpu --no-e2e $mode > $OUT/bench_n${N}${mode}.json 2>> $OUT/bench.err
python -c "import json;d=json.load(open('$OUT/bench_n${N}${mode}.json'));print('N=$N $mode', round(d['ms_per_step'],3),'ms', round(d['value']/1e9,1),'G cells/s, kernel', round(d['roofline']['kernel_ms'],3))" | tee -a $OUT/summary.txt
done
