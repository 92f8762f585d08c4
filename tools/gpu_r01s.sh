#!/bin/bash
# N-GPU round: NCCL parity check + bench at N GPUs.  usage: bash tools/gpu_r01s.sh <tag> <N>
TAG=$1; N=$2
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/xc_nccl_check.py > $OUT/xc_check.log 2>&1; echo "rc=$?" >> $OUT/xc_check.log; tail -3 $OUT/xc_check.log
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?" >> $OUT/bench_n$N.err
cat $OUT/bench_n$N.json | cut -c1-1200; tail -3 $OUT/bench_n$N.err
timeout 600 $TR bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-e2e --temdf2 0.01 > $OUT/bench_n${N}_temdf2.json 2>> $OUT/bench_n$N.err
cat $OUT/bench_n${N}_temdf2.json | cut -c1-400
timeout 300 $TR bench.py --impl reference --gpus $N --steps 1 --warmup 1 > $OUT/bench_ref_n$N.json 2>> $OUT/bench_n$N.err
cat $OUT/bench_ref_n$N.json | cut -c1-600
