#!/bin/bash
# 2-GPU round: NCCL parity check (all cases incl. the arctic fold) and the bench at N=2 with the split launch
TAG=${1:-r01zd}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 400 $TR tools/xc_nccl_check.py > $OUT/xc_check.log 2>&1; echo "rc=$?" >> $OUT/xc_check.log; tail -6 $OUT/xc_check.log
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "rc=$?" >> $OUT/bench_n2.err
cut -c1-1200 $OUT/bench_n2.json; tail -3 $OUT/bench_n2.err
