OUT=gpurun_out/r01i; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
for cfg in "1 1 3 8" "0 2 2 8" "1 2 2 8" "1 1 3 0"; do set -- $cfg
  export HYCOM_TSADVC_TMA=$1 HYCOM_TSADVC_NC=$2 HYCOM_TSADVC_MINB=$3
  r=$(timeout 400 python bench.py --advtyp 1 --ntracr $4 --steps 3 --warmup 3 --no-cpu --no-e2e 2>>$OUT/err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])")
  echo "MPDATA TMA=$1 NC=$2 MINB=$3 ntracr=$4: $r" | tee -a $OUT/variants.txt
done
tail -5 $OUT/err.log
