#!/bin/bash
TAG=${1:-r01p}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
for v in "1 3 512" "2 2 512" "1 3 1024" "1 3 256" "2 2 1024" "1 4 512"; do set -- $v
  export HYCOM_TSADVC_NC=$1 HYCOM_TSADVC_MINB=$2 HYCOM_TSADVC_CHUNK_ROWS=$3
  r=$(timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e 2>>$OUT/err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])")
  echo "NC=$1 MINB=$2 CHUNK=$3: $r" | tee -a $OUT/variants.txt
done
unset HYCOM_TSADVC_NC HYCOM_TSADVC_MINB HYCOM_TSADVC_CHUNK_ROWS
timeout 300 python bench.py --advtyp 1 --ntracr 8 --steps 4 --no-cpu --no-e2e > $OUT/bench_mpdata8.json 2>>$OUT/err.log; cat $OUT/bench_mpdata8.json | cut -c1-400
