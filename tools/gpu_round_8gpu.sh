#!/bin/bash
# gpurun --gpus 8: the frame-sharded all-intra bench on all eight GPUs of the box (what the driver's scaling run does at N=8)
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
python - <<'P'
import json
try:
    d=json.load(open('gpurun_out/bench_8gpu.json')); print('8gpu', round(d['value'],2), round(d['e2e']['value'],2), d['host_ms_per_gof'], d.get('exchange'))
except Exception as e: print('8gpu ERR', e)
P
tail -3 gpurun_out/bench_8gpu.err
