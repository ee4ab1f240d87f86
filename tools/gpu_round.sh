#!/bin/bash
mkdir -p gpurun_out
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
nproc; lscpu | grep "Model name"
PCCB200_SWEEP_CTAS_PER_SM=4 run A4 8 24 32 --no-cpu-baseline
PCCB200_SWEEP_CTAS_PER_SM=16 run B16 8 24 32 --no-cpu-baseline
PCCB200_SWEEP_CTAS_PER_SM=4 run A4again 8 24 32 --no-cpu-baseline
