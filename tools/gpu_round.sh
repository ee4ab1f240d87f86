#!/bin/bash
# One gpurun call of round 2 (1 GPU): GPU test-suite incl. the full-size digests, the parked post-reconstruction tests, the default
# bench line, host-core sensitivity (the share of the 16-core box one of 8 ranks gets) and the ncu launch list with DRAM bytes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
( timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
( timeout 300 python -m pytest tests/gpu_pending_postrecon.py -m gpu -q ) > gpurun_out/pytest_pending.log 2>&1; tail -6 gpurun_out/pytest_pending.log
summ() { python - "$1" <<'P'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1])); print(sys.argv[1], round(d['value'],2), round(d['e2e']['value'],2), d.get('parity_ok'), d['host_ms_per_gof'], d.get('canvas_exchange'))
except Exception as e: print(sys.argv[1],'ERR', e)
P
}
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; summ default
taskset -c 0,1 timeout 300 python bench.py --steps 16 --no-cpu-baseline > gpurun_out/bench_2cores.json 2> gpurun_out/bench_2cores.err; summ 2cores
taskset -c 0-3 timeout 300 python bench.py --steps 16 --no-cpu-baseline > gpurun_out/bench_4cores.json 2> gpurun_out/bench_4cores.err; summ 4cores
timeout 300 python bench.py --steps 4 --warmup 1 --gofs-in-flight 1 --no-cpu-baseline > gpurun_out/bench_1gof.json 2> gpurun_out/bench_1gof.err; summ 1gof
# launch list of one full-size frame through a1-a26 with DRAM traffic per launch
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_traffic.csv \
  python bench.py --frames 1 --steps 1 --warmup 1 --gofs-in-flight 1 --no-cpu-baseline > gpurun_out/ncu_list.out 2> gpurun_out/ncu_list.err
echo "ncu rc=$?"; wc -l gpurun_out/launches_traffic.csv
