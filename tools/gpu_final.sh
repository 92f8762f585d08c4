#!/bin/bash
# end-of-round evidence: parity tests, smoke, bench lines, ncu launch list, ncu full capture of the
# marching kernel (the bench launch itself), compute-sanitizer memcheck of a few small cases
TAG=${1:-r01zf}
OUT=gpurun_out/$TAG; mkdir -p $OUT
{ nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv; nproc; which gfortran flang nvfortran ifx 2>&1 | head -3; } > $OUT/env.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.err
cat $OUT/bench.json | cut -c1-1500
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json | cut -c1-400
timeout 600 python bench.py --temdf2 0.01 --steps 6 --no-e2e --no-cpu > $OUT/bench_temdf2.json 2>> $OUT/bench.err; cat $OUT/bench_temdf2.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_run.log 2>&1
# both marching launches of one call (general segments, then all-sea segments): skip the 3 warm-up calls
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tsadvc_march -s 6 -c 2 \
   -o $OUT/prof_fct2_split_k41 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/full_run.log 2>&1
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x \
   -k "golden_vectors_on_device and (box_fct2 or diffusion_17t or fct2c or arctic or periodic_mpdata)" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?" >> $OUT/memcheck.log
tail -8 $OUT/memcheck.log
ls $OUT
