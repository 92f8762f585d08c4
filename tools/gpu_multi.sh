#!/bin/bash
# multi-GPU round: NCCL parity check + bench at N GPUs.  usage: bash tools/gpu_multi.sh <tag> <N> [big]
TAG=$1; N=$2; BIG=$3
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/xc_nccl_check.py > $OUT/xc_check.log 2>&1; echo "rc=$?" >> $OUT/xc_check.log; tail -3 $OUT/xc_check.log
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?" >> $OUT/bench_n$N.err
cat $OUT/bench_n$N.json; tail -3 $OUT/bench_n$N.err
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e --no-overlap > $OUT/bench_n${N}_nooverlap.json 2>> $OUT/bench_n$N.err
cat $OUT/bench_n${N}_nooverlap.json
if [ -n "$BIG" ]; then
  timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-e2e --workload GLBy0.04 > $OUT/bench_n${N}_GLBy.json 2>> $OUT/bench_n$N.err
  cat $OUT/bench_n${N}_GLBy.json
fi
