#!/bin/bash
# One gpurun call (1 GPU): full GPU test-suite + smoke + the default bench line and the single-GOF latency (end-of-round state)
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/pytest_gpu.log 2>&1; tail -10 gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
summ() { python - "$1" <<'P'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1])); print(sys.argv[1], round(d['value'],2), round(d['e2e']['value'],2), d.get('parity_ok'), d['gpu_mem_used_gb'], d['host_ms_per_gof'], {k:v for k,v in list(d['stage_ms_per_frame'].items())[:5]})
except Exception as e: print(sys.argv[1],'ERR', e)
P
}
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; summ default
timeout 300 python bench.py --steps 4 --warmup 1 --gofs-in-flight 1 --no-cpu-baseline > gpurun_out/bench_1gof.json 2> gpurun_out/bench_1gof.err; summ 1gof
timeout 600 python bench.py --condition ra --steps 32 --no-cpu-baseline > gpurun_out/bench_ra_r5.json 2> gpurun_out/bench_ra_r5.err; summ ra_r5
