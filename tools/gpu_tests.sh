#!/bin/bash
# GPU parity tests only.  usage: bash tools/gpu_tests.sh <tag>
TAG=${1:-tests}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -40 $OUT/pytest_gpu.log
