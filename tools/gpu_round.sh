#!/bin/bash
# One gpurun call (1 GPU): configuration A/B of the default bench (scratch sets, sweep-tail grid)
mkdir -p gpurun_out
summ() { python - "$1" <<'P'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1])); print(sys.argv[1], round(d['value'],2), round(d['e2e']['value'],2), d['gpu_mem_used_gb'], d['host_ms_per_gof'], {k:v for k,v in list(d['stage_ms_per_frame'].items())[:5]})
except Exception as e: print(sys.argv[1],'ERR', e)
P
}
timeout 400 python bench.py --no-cpu-baseline --scratch-sets 32 > gpurun_out/bench_sets32.json 2> gpurun_out/bench_sets32.err; summ sets32
timeout 400 python bench.py --no-cpu-baseline --scratch-sets 40 > gpurun_out/bench_sets40.json 2> gpurun_out/bench_sets40.err; summ sets40
PCCB200_SWEEP_TAIL_CTAS_PER_SM=2 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_tail2.json 2> gpurun_out/bench_tail2.err; summ tail2
timeout 400 python bench.py --no-cpu-baseline --scratch-sets 16 > gpurun_out/bench_sets16.json 2> gpurun_out/bench_sets16.err; summ sets16
