#!/bin/bash
# full ncu capture of the FCT2 march with the per-row all-sea branch (both bodies in the loop)
TAG=${1:-r01z}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tsadvc_march -s 3 -c 1 \
   -o $OUT/prof_fct2_period3 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/full_run.log 2>&1
ls -la $OUT
