#!/bin/bash
# One parameterised GPU session (replaces the round-1 one-off scripts).
#   gpurun [--gpus N] --timeout T -- 'bash tools/gpu_run.sh TAG STEP [STEP ...]'
# Steps (each writes under gpurun_out/TAG/):
#   env        toolchain probe (gfortran & co), nvidia-smi, NCCL version
#   tests      python -m pytest tests -m gpu
#   smoke      __graft_entry__.smoke()
#   bench      bench.py at N = $NGPU (default 1), driver flags
#   benchq     bench.py quick (5 steps, no e2e / cpu / extra)
#   ref        bench.py --impl reference
#   launches   ncu launch list of a short bench run (N=1)
#   cnuity     cnuity parity tests + timing at GLBb0.08; cnuity_launches / cnuity_ncu: per-kernel times, full capture
#   sanitize   compute-sanitizer memcheck + racecheck on small cases of the newer kernels
#   traffic    DRAM bytes + L2 hit rate of the marching launches of one full-size step
#   ncu        ncu --set full + source of the marching kernel (N=1, reduced kdm)
#   tma        the tensor-map TMA probe, every variant in its own process
#   nccl       tools/xc_nccl_check.py under torchrun (N = $NGPU >= 2)
#   scale      bench.py at N=1,2,4,8 as far as $NGPU allows (no e2e/cpu)
#   fortran    fortran/build_ref.sh if a Fortran compiler exists
#   variants   kernel variant timings (env knobs listed in $VARIANTS, ';'-separated)
TAG=${1:?tag}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
NGPU=${NGPU:-1}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29533"
for STEP in "$@"; do
  echo "=== $STEP"
  case $STEP in
    env)
      { for c in gfortran flang flang-new nvfortran ifort ifx lfortran f2c mpif90 mpirun; do printf "%s: " $c; which $c || echo none; done
        nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,memory.total --format=csv
        nproc; free -g | head -2
        python -c "import importlib,ctypes as C; p=importlib.import_module('hycom-src_b200'); l=p.load_library(); v=C.c_int32(0); print('nccl rc', l.hycom_tsadvc_comm_version(C.byref(v)), 'version', v.value)"
      } > $OUT/env.txt 2>&1; cat $OUT/env.txt ;;
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log ;;
    bench)
      if [ $NGPU -gt 1 ]; then L="$TR --nproc-per-node $NGPU"; else L=python; fi
      timeout 900 $L bench.py --gpus $NGPU --steps 20 --warmup 5 $BENCH_ARGS > $OUT/bench_n$NGPU.json 2> $OUT/bench_n$NGPU.err; echo "rc=$?"; cut -c1-1800 $OUT/bench_n$NGPU.json; tail -3 $OUT/bench_n$NGPU.err ;;
    benchq)
      timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extra $BENCH_ARGS > $OUT/benchq.json 2> $OUT/benchq.err; echo "rc=$?"; cut -c1-900 $OUT/benchq.json; tail -3 $OUT/benchq.err ;;
    ref)
      timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/ref.json 2> $OUT/ref.err; echo "rc=$?"; cut -c1-1500 $OUT/ref.json; tail -3 $OUT/ref.err ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
        python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > $OUT/launches_run.log 2>&1; echo "rc=$?"; tail -2 $OUT/launches_run.log | cut -c1-300 ;;
    ncu)
      # general launch first, all-sea second in every step: skip the 3 warm-up steps
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-k_tsadvc_march} -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-2} -f -o $OUT/prof \
        python bench.py --kdm ${NCU_KDM:-6} --steps 1 --warmup 3 --no-e2e --no-cpu --no-extra $BENCH_ARGS > $OUT/ncu_run.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_run.log | cut -c1-300
      ls -la $OUT/prof.ncu-rep ;;
    traffic)
      # DRAM bytes of the two marching launches of one full-size step (few counters: one replay pass)
      timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
        -k regex:k_tsadvc_march -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-2} --csv --log-file $OUT/traffic${TRAFFIC_TAG}.csv \
        python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extra $BENCH_ARGS > $OUT/traffic_run.log 2>&1; echo "rc=$?"
      python - $OUT/traffic${TRAFFIC_TAG}.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = 0.0
for r in rows[1:]:
    print(r[ik][:60], r[im], r[iv], r[iu])
    if r[im].startswith("dram__bytes"):
        tot += float(r[iv].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[r[iu]]
print("dram total %.3f GB" % (tot / 1e9))
PY
      ;;
    cnuity)
      timeout 900 python -m pytest tests/test_cnuity_gpu.py -x -q > $OUT/pytest_cnuity.log 2>&1; echo "rc=$?" >> $OUT/pytest_cnuity.log; tail -4 $OUT/pytest_cnuity.log
      timeout 600 python tools/cnuity_timing.py 2>&1 | tail -1 | tee -a $OUT/cnuity_timing.txt ;;
    cnuity_launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_cn_\|k_halo -c 400 --csv --log-file $OUT/cnuity_launches.csv \
        python tools/cnuity_timing.py 1 1 > $OUT/cnuity_launches_run.log 2>&1; echo "rc=$?"
      python - $OUT/cnuity_launches.csv <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
t = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}.get(r[iu], 1e-6)
    n, s = t.get(r[ik][:70], (0, 0.0)); t[r[ik][:70]] = (n + 1, s + v)
for k, (n, s) in t.items():
    print(f"{s:9.3f} ms {n:4d}x  {k}")
PY
      ;;
    cnuity_ncu)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:${CN_KERNEL:-k_cn_loop76} -c 1 -f -o $OUT/cnuity_prof \
        python tools/cnuity_timing.py 1 0 > $OUT/cnuity_ncu_run.log 2>&1; echo "rc=$?"; tail -2 $OUT/cnuity_ncu_run.log | cut -c1-200 ;;
    sanitize)
      # memcheck and racecheck (shared-memory hazards) of the newer kernels on small cases
      SEL=${SANITIZE_SEL:-"cnuity_device_matches_oracle or thickness_diffusion_matches_oracle or isopyc_with_tracers"}
      timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_cnuity_gpu.py tests/test_parity_gpu.py -x -q -k "$SEL" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/memcheck.log; grep -c "Invalid\|ERROR SUMMARY" $OUT/memcheck.log; tail -3 $OUT/memcheck.log
      timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_cnuity_gpu.py -x -q -k "cnuity_device_matches_oracle" > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/racecheck.log; tail -3 $OUT/racecheck.log ;;
    tma)
      (cd tools/probe && nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe4 tma_probe4.cu -lcuda 2>/dev/null)
      for v in 0 1 2 3 4; do timeout 60 tools/probe/tma_probe4 $v; echo "exit=$?"; done > $OUT/tma_probe4.txt 2>&1; cat $OUT/tma_probe4.txt ;;
    tma5)
      (cd tools/probe && nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe5 tma_probe5.cu -lcuda 2>/dev/null)
      for args in "0 2 58 0" "0 18 58 0" "0 34 58 0" "0 2 64 0" "0 2 32 0" "0 2 58 1" "0 2 58 2" "0 2 64 2" "0 3 64 1"; do timeout 60 tools/probe/tma_probe5 $args; echo "exit=$?"; done > $OUT/tma_probe5.txt 2>&1; cat $OUT/tma_probe5.txt ;;
    nccl)
      timeout 900 $TR --nproc-per-node $NGPU tools/xc_nccl_check.py > $OUT/xc_nccl_check_n$NGPU.log 2>&1; echo "rc=$?" >> $OUT/xc_nccl_check_n$NGPU.log; grep -v "^W\|^\[W\|warn" $OUT/xc_nccl_check_n$NGPU.log | tail -8
      if [ -n "$XC_TILES" ]; then XC_CHECK_TILES=$XC_TILES XC_CHECK_CASES=arctic timeout 600 $TR --nproc-per-node $NGPU tools/xc_nccl_check.py > $OUT/xc_nccl_check_$XC_TILES.log 2>&1; echo "rc=$?" >> $OUT/xc_nccl_check_$XC_TILES.log; tail -3 $OUT/xc_nccl_check_$XC_TILES.log; fi ;;
    scale)
      for n in 1 2 4 8; do
        [ $n -gt $NGPU ] && break
        if [ $n -gt 1 ]; then L="$TR --nproc-per-node $n"; else L=python; fi
        timeout 600 $L bench.py --gpus $n --steps 20 --warmup 5 --no-cpu $SCALE_ARGS > $OUT/scale_n$n.json 2> $OUT/scale_n$n.err; echo "n=$n rc=$?"; cut -c1-700 $OUT/scale_n$n.json; tail -2 $OUT/scale_n$n.err
      done ;;
    fortran)
      if which gfortran > /dev/null 2>&1; then bash fortran/build_ref.sh > $OUT/fortran_ref.log 2>&1; tail -5 $OUT/fortran_ref.log; else echo "no gfortran on this box" | tee $OUT/fortran_ref.log; fi ;;
    variants)
      IFS=';' read -ra VS <<< "$VARIANTS"
      for v in "${VS[@]}"; do
        name=$(echo "$v" | tr ' =/' '___')
        env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-extra > $OUT/var_$name.json 2> $OUT/var_$name.err
        python - "$v" $OUT/var_$name.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:60s} ms/step {d['ms_per_step']:.3f} kernel {d['roofline']['kernel_ms']:.3f} frac {d['roofline']['frac']:.4f} cks {d['checksum']['value']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
      done | tee $OUT/variants.txt ;;
    *) echo "unknown step $STEP" ;;
  esac
done
