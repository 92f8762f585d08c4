#!/bin/bash
TAG=${1:-r01m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file $OUT/launches_temdf2.csv python bench.py --temdf2 0.01 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/launches_run.log 2>&1
grep k_tsdff $OUT/launches_temdf2.csv | head -3
timeout 600 ncu --set full --clock-control none -k regex:k_tsdff -s 2 -c 1 \
   -o $OUT/prof_tsdff -f python bench.py --temdf2 0.01 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/full_tsdff.log 2>&1
ls -la $OUT
