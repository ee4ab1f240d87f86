#!/bin/bash
mkdir -p gpurun_out
run() { # name lanes sets steps extra
  timeout 600 python bench.py --gofs-in-flight $2 --scratch-sets $3 --steps $4 --warmup 3 --no-cpu-baseline $5 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  echo "$1 rc=$?"; python - <<P
import json
try:
    d=json.load(open('gpurun_out/bench_$1.json')); print(round(d['value'],2), round(d['e2e']['value'],2), d['gpu_mem_used_gb'], d['host_ms_per_gof']); print(d['stage_ms_per_frame']); print(d['gof_log_ms'][-6:])
except Exception as e: print('ERR', e)
P
  tail -2 gpurun_out/bench_$1.err
}
run L4block 4 24 16
PCCB200_SPIN_WAIT=1 run L4spin 4 24 16
run L8block 8 24 32
PCCB200_SPIN_WAIT=1 run L8spin 8 24 32
