#!/bin/bash
# all-sea fast path of the FCT2 march: parity tests, then NC/MINB variants of the bench launch
TAG=${1:-r01y}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
for v in "2 2 1024" "1 3 1024" "2 2 512"; do set -- $v
  export HYCOM_TSADVC_NC=$1 HYCOM_TSADVC_MINB=$2 HYCOM_TSADVC_CHUNK_ROWS=$3
  r=$(timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e 2>>$OUT/err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])")
  echo "NC=$1 MINB=$2 CHUNK=$3: $r" | tee -a $OUT/variants.txt
done
tail -5 $OUT/err.log
