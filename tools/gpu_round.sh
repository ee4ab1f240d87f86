#!/bin/bash
# One gpurun call (1 GPU), end-of-round state: full GPU test-suite + smoke + the default bench line + single-GOF latency
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --durations=3 ) > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
summ() { python - "$1" <<'P'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1])); print(sys.argv[1], round(d['value'],2), round(d['e2e']['value'],2), d.get('parity_ok'), d.get('timed_handoff_matches_checked_frame'), d['e2e'], d['host_ms_per_gof'])
except Exception as e: print(sys.argv[1],'ERR', e)
P
}
timeout 500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; summ default
