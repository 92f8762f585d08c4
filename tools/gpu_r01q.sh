#!/bin/bash
TAG=${1:-r01q}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python bench.py --e2e-steps 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.err
cat $OUT/bench.json | cut -c1-1400
for v in 0 1 2 3; do
  export HYCOM_TSADVC_TSDFF_VARIANT=$v
  r=$(timeout 300 python bench.py --temdf2 0.01 --steps 6 --warmup 3 --no-cpu --no-e2e 2>>$OUT/err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['clocks']['sm_mhz'])")
  echo "TSDFF_VARIANT=$v: step_ms march_ms clock = $r" | tee -a $OUT/tsdff_variants.txt
done
