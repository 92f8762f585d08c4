#!/bin/bash
# 4 GPUs as 4x1 arctic tiles: the fold messages involve a third tile (NW/NE partners differ from the twin),
# which is where NCCL's match-by-order is exercised
TAG=${1:-r01zj}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533"
XC_CHECK_CASES=arctic XC_CHECK_TILES=4x1 timeout 200 $TR tools/xc_nccl_check.py > $OUT/xc_check_arctic_4x1.log 2>&1; echo "rc=$?" >> $OUT/xc_check_arctic_4x1.log; tail -4 $OUT/xc_check_arctic_4x1.log
