#!/bin/bash
# parity tests, bench line (pipelined host-array e2e), diffusion timing + launch list
TAG=${1:-r01l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 900 python bench.py --e2e-steps 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.err
cat $OUT/bench.json
for c in 2 8; do
HYCOM_TSADVC_STEP_CHUNK=$c timeout 600 python bench.py --steps 3 --no-cpu --e2e-steps 3 > $OUT/bench_chunk$c.json 2>> $OUT/bench.err
python -c "import json;d=json.load(open('$OUT/bench_chunk$c.json'));print('chunk $c e2e ms', d['e2e']['ms_per_step'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file $OUT/launches_temdf2.csv python bench.py --temdf2 0.01 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/launches_run.log 2>&1
grep -c k_tsdff $OUT/launches_temdf2.csv
timeout 600 ncu --set full --clock-control none -k regex:k_tsdff -s 2 -c 1 \
   -o $OUT/prof_tsdff -f python bench.py --temdf2 0.01 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/full_tsdff.log 2>&1
ls -la $OUT
