#!/bin/bash
# One gpurun call (1 GPU): sanity subset of the GPU tests, walk prefetch A/B, sweep-tail grid A/B, the other BASELINE configurations
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_segment.py tests/test_gpu_gof.py tests/test_gpu_knn_normals.py -m gpu -x -q ) > gpurun_out/pytest_gpu_subset.log 2>&1; tail -3 gpurun_out/pytest_gpu_subset.log
summ() { python - "$1" <<'P'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1])); print(sys.argv[1], round(d['value'],2), round(d['e2e']['value'],2), d.get('parity_ok'), d['host_ms_per_gof'], {k:v for k,v in list(d['stage_ms_per_frame'].items())[:7]})
except Exception as e: print(sys.argv[1],'ERR', e)
P
}
PCCB200_WALK_EARLY_PREFETCH=0 timeout 300 python bench.py --steps 4 --warmup 1 --gofs-in-flight 1 --no-cpu-baseline > gpurun_out/bench_1gof_late.json 2> gpurun_out/bench_1gof_late.err; summ 1gof_late
timeout 300 python bench.py --steps 4 --warmup 1 --gofs-in-flight 1 --no-cpu-baseline > gpurun_out/bench_1gof.json 2> gpurun_out/bench_1gof.err; summ 1gof
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; summ default
PCCB200_SWEEP_TAIL_CTAS_PER_SM=8 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_default_tail8.json 2> gpurun_out/bench_default_tail8.err; summ default_tail8
PCCB200_WALK_EARLY_PREFETCH=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_default_late.json 2> gpurun_out/bench_default_late.err; summ default_late
timeout 600 python bench.py --condition ra --steps 16 > gpurun_out/bench_ra_r5.json 2> gpurun_out/bench_ra_r5.err; summ ra_r5
timeout 900 python bench.py --bits 11 --frames 8 --gofs-in-flight 4 --scratch-sets 8 --steps 12 --warmup 2 > gpurun_out/bench_vox11.json 2> gpurun_out/bench_vox11.err; summ vox11
