#!/bin/bash
# 8-GPU round: NCCL parity check, GLBb0.08 and GLBy0.04 bench lines
TAG=$1; N=$2
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/xc_nccl_check.py > $OUT/xc_check.log 2>&1; echo "rc=$?" >> $OUT/xc_check.log; tail -2 $OUT/xc_check.log
timeout 600 $TR bench.py --gpus $N --steps 12 --warmup 3 --no-cpu > $OUT/bench_n$N.json 2>> $OUT/bench.err
timeout 600 $TR bench.py --gpus $N --steps 12 --warmup 3 --no-cpu --no-e2e --frame-serial > $OUT/bench_n${N}_frame_serial.json 2>> $OUT/bench.err
timeout 600 $TR bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-e2e --workload GLBy0.04 > $OUT/bench_n${N}_GLBy0.04.json 2>> $OUT/bench.err
for f in $OUT/bench_n*.json; do python - $f <<'PY'
import json,sys
for line in open(sys.argv[1]):
    line=line.strip()
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], round(d['ms_per_step'],3),'ms', round(d['value']/1e9,1),'G cells/s, kernel', round(d['roofline']['kernel_ms'],3), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1))
PY
done
