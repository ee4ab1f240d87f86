#!/bin/bash
mkdir -p gpurun_out
( timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
( PCCB200_SWEEP_STATIC=1 timeout 200 python -m pytest tests/test_gpu_segment.py tests/test_gpu_gof.py -m gpu -x -q -k "not full_size" ) > gpurun_out/pytest_gpu_static.log 2>&1; tail -3 gpurun_out/pytest_gpu_static.log
run() { # name env
  env $2 timeout 300 python bench.py --steps 24 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  echo "$1 rc=$?"; python - <<P
import json
try:
    d=json.load(open('gpurun_out/bench_$1.json')); print(round(d['value'],2), round(d['e2e']['value'],2), d['stage_ms_per_frame'].get('refine'), d['host_ms_per_gof'])
except Exception as e: print('ERR', e)
P
}
run static1 PCCB200_SWEEP_STATIC=1
run static0 PCCB200_SWEEP_STATIC=0
