#!/bin/bash
TAG=${1:-r01v}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "asselin" > $OUT/pytest_asselin.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_asselin.log
tail -25 $OUT/pytest_asselin.log
timeout 600 python tools/asselin_timing.py > $OUT/asselin_timing.json 2> $OUT/asselin_timing.err; cat $OUT/asselin_timing.json; tail -3 $OUT/asselin_timing.err
