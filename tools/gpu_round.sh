#!/bin/bash
# One gpurun call (1 GPU): GPU test-suite (incl. the application-level drop-in), default bench + single-GOF latency, launch list.
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
summ() { python - "$1" <<'P'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1])); print(sys.argv[1], round(d['value'],2), round(d['e2e']['value'],2), d.get('parity_ok'), d['host_ms_per_gof'], {k:v for k,v in list(d['stage_ms_per_frame'].items())[:8]})
except Exception as e: print(sys.argv[1],'ERR', e)
P
}
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; summ default
timeout 300 python bench.py --steps 4 --warmup 1 --gofs-in-flight 1 --no-cpu-baseline > gpurun_out/bench_1gof.json 2> gpurun_out/bench_1gof.err; summ 1gof
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_traffic.csv \
  python bench.py --frames 1 --steps 1 --warmup 0 --gofs-in-flight 1 --no-cpu-baseline > gpurun_out/ncu_list.out 2> gpurun_out/ncu_list.err
echo "ncu list rc=$?"; wc -l gpurun_out/launches_traffic.csv
