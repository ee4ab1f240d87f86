#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
run() { # name lanes sets steps extra
  timeout 600 python bench.py --gofs-in-flight $2 --scratch-sets $3 --steps $4 --warmup 3 $5 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  echo "$1 rc=$?"; python - <<P
import json
try:
    d=json.load(open('gpurun_out/bench_$1.json')); print(round(d['value'],2), round(d['e2e']['value'],2), d['gpu_mem_used_gb'], d['host_ms_per_gof'], d.get('cpu_baseline')); print(d['stage_ms_per_frame'])
except Exception as e: print('ERR', e)
P
  tail -2 gpurun_out/bench_$1.err
}
run N8 8 24 32 --no-cpu-baseline
PCCB200_SWEEP_CTAS_PER_SM=2 run N8c2 8 24 32 --no-cpu-baseline
run N6 6 24 24 --no-cpu-baseline
PCCB200_SWEEP_MAX_SLEEP_NS=128 run N8s128 8 24 32 --no-cpu-baseline
