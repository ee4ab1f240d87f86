#!/bin/bash
# One gpurun call (1 GPU): random-access bench twice (run-to-run spread) + the default line, after giving the context stream top priority
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ra_pack.py tests/test_gpu_segment.py -m gpu -x -q ) > gpurun_out/pytest_gpu_subset.log 2>&1; tail -2 gpurun_out/pytest_gpu_subset.log
summ() { python - "$1" <<'P'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1])); print(sys.argv[1], round(d['value'],2), round(d['e2e']['value'],2), d.get('parity_ok'), d['host_ms_per_gof'], {k:v for k,v in list(d['stage_ms_per_frame'].items())[:5]})
except Exception as e: print(sys.argv[1],'ERR', e)
P
}
timeout 400 python bench.py --condition ra --steps 32 --no-cpu-baseline > gpurun_out/bench_ra_r5.json 2> gpurun_out/bench_ra_r5.err; summ ra_r5
timeout 400 python bench.py --condition ra --steps 32 --no-cpu-baseline > gpurun_out/bench_ra_r5_b.json 2> gpurun_out/bench_ra_r5_b.err; summ ra_r5_b
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_default_prio.json 2> gpurun_out/bench_default_prio.err; summ default_prio
