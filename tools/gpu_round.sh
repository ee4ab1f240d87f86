#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/bench_yuv.json 2> gpurun_out/bench_yuv.err; echo "rc=$?"; python - <<P
import json
try:
    d=json.load(open('gpurun_out/bench_yuv.json')); print(round(d['value'],2), d['e2e'], d['gpu_mem_used_gb'], d['host_ms_per_gof'])
except Exception as e: print('ERR', e)
P
tail -4 gpurun_out/bench_yuv.err
