#!/bin/bash
# 2-GPU round: GPU tests on one GPU, NCCL parity check (advection, diffusion, fct2c) on two
TAG=${1:-r01n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR tools/xc_nccl_check.py > $OUT/xc_check.log 2>&1; echo "rc=$?" >> $OUT/xc_check.log; tail -8 $OUT/xc_check.log
