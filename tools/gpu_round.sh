#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list, ncu full capture of the marching kernel.
# usage (from the repo root, on the GPU box): bash tools/gpu_round.sh <tag> [steps...]
TAG=${1:-r01}; shift
STEPS=${@:-"env tests bench launches full"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for s in $STEPS; do
case $s in
env)
  { nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv; nproc; free -g | head -2;
    which gfortran flang nvfortran ifx 2>&1 | head -3; } > $OUT/env.txt 2>&1 ;;
tests)
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log ;;
smoke)
  timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "rc=$?" >> $OUT/smoke.log ;;
bench)
  timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.err ;;
benchref)
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err ;;
mpdata)
  timeout 900 python bench.py --advtyp 1 --ntracr 8 --steps 5 --no-cpu > $OUT/bench_mpdata.json 2> $OUT/bench_mpdata.err ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
     --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_run.log 2>&1 ;;
full)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tsadvc_march -s 3 -c 1 \
     -o $OUT/prof_fct2 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --kdm 8 > $OUT/full_run.log 2>&1 ;;
fullmp)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tsadvc_march -s 3 -c 1 \
     -o $OUT/prof_mpdata -f python bench.py --advtyp 1 --ntracr 2 --steps 2 --warmup 3 --no-e2e --no-cpu --kdm 8 > $OUT/fullmp_run.log 2>&1 ;;
esac
done
ls -la $OUT
tail -3 $OUT/pytest_gpu.log 2>/dev/null
cat $OUT/bench.json 2>/dev/null
