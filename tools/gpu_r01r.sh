#!/bin/bash
TAG=${1:-r01r}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "diffusion or mxlmy or golden or isopyc" > $OUT/pytest_diff.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_diff.log
tail -15 $OUT/pytest_diff.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
run() { r=$(timeout 300 python bench.py --temdf2 0.01 --steps 6 --warmup 3 --no-cpu --no-e2e 2>>$OUT/err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['clocks']['sm_mhz'])"); echo "$1: step_ms march_ms clock = $r" | tee -a $OUT/tsdff_variants.txt; }
run "march (default)"
HYCOM_TSADVC_TSDFF=column run "column MINB=4"
HYCOM_TSADVC_TSDFF=column HYCOM_TSADVC_TSDFF_VARIANT=5 run "column MINB=5"
HYCOM_TSADVC_TSDFF=column HYCOM_TSADVC_TSDFF_VARIANT=6 run "column MINB=6"
timeout 600 ncu --set full --clock-control none -k regex:k_tsdff_march -s 2 -c 1 \
   -o $OUT/prof_tsdff_march -f python bench.py --temdf2 0.01 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/full_tsdff.log 2>&1
ls $OUT
