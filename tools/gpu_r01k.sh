#!/bin/bash
# round-1 session-4 GPU run: parity tests (incl. diffusion/EOS), bench line, diffusion timing,
# ncu launch lists, ncu full capture of the marching kernel at the full kdm
TAG=${1:-r01k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
{ nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv; nproc; free -g | head -2;
  which gfortran flang nvfortran ifx 2>&1 | head -3; } > $OUT/env.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.err
cat $OUT/bench.json
timeout 600 python bench.py --temdf2 0.01 --steps 6 --no-e2e --no-cpu > $OUT/bench_temdf2.json 2> $OUT/bench_temdf2.err
cat $OUT/bench_temdf2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file $OUT/launches_temdf2.csv python bench.py --temdf2 0.01 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/launches_run.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_tsdff -s 2 -c 1 \
   -o $OUT/prof_tsdff -f python bench.py --temdf2 0.01 --steps 2 --warmup 3 --no-e2e --no-cpu --kdm 8 > $OUT/full_tsdff.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tsadvc_march -s 3 -c 1 \
   -o $OUT/prof_fct2_k41 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/full_run.log 2>&1
ls -la $OUT
