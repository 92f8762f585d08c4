#!/bin/bash
# MPDATA with the split launch (all-sea segments + the rest, period-3 rings): parity tests, config 3 bench
TAG=${1:-r01ze}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 400 python bench.py --advtyp 1 --ntracr 8 --steps 4 --warmup 3 --no-cpu --no-e2e > $OUT/bench_mpdata8.json 2>$OUT/err.log; cut -c1-1300 $OUT/bench_mpdata8.json
HYCOM_TSADVC_SPLIT=0 timeout 400 python bench.py --advtyp 1 --ntracr 8 --steps 4 --warmup 3 --no-cpu --no-e2e > $OUT/bench_mpdata8_nosplit.json 2>>$OUT/err.log; cut -c1-500 $OUT/bench_mpdata8_nosplit.json
tail -3 $OUT/err.log
