#!/bin/bash
# bench the marching-kernel variants selected by environment variables; one line per variant
# usage: bash tools/gpu_variants.sh <tag> "TMA NC MINB [CHUNK]" ...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
for v in "$@"; do set -- $v
  export HYCOM_TSADVC_TMA=$1 HYCOM_TSADVC_NC=$2 HYCOM_TSADVC_MINB=$3 HYCOM_TSADVC_CHUNK_ROWS=${4:-512}
  r=$(timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e ${EXTRA} 2>>$OUT/err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])")
  echo "TMA=$1 NC=$2 MINB=$3 CHUNK=${4:-512}: $r" | tee -a $OUT/variants.txt
done
tail -5 $OUT/err.log
