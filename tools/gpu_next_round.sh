#!/bin/bash
# First GPU call of the next round (N GPUs of one box): where does the multi-GPU step stand with the
# split launch, and does the arctic fold pass under NCCL on the 8-GPU tiling?
# usage: gpurun --gpus 8 -- 'bash tools/gpu_next_round.sh r02a 8'
TAG=${1:-r02a}; N=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
XC_CHECK_CASES=arctic timeout 300 $TR tools/xc_nccl_check.py > $OUT/xc_check_arctic_n$N.log 2>&1; echo "rc=$?" >> $OUT/xc_check_arctic_n$N.log; tail -3 $OUT/xc_check_arctic_n$N.log
for extra in "" "--no-overlap"; do
  timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e $extra > $OUT/bench_n$N$extra.json 2>> $OUT/bench.err
  cut -c1-700 $OUT/bench_n$N$extra.json
done
HYCOM_TSADVC_SPLIT=0 timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e > $OUT/bench_n${N}_nosplit.json 2>> $OUT/bench.err
cut -c1-400 $OUT/bench_n${N}_nosplit.json
tail -3 $OUT/bench.err
