#!/bin/bash
# gpurun --gpus 2: the frame-sharded RANDOM-ACCESS bench over NCCL (one all-gather of patch records per GOF, packing replicated)
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --condition ra --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_ra.json 2> gpurun_out/bench_2gpu_ra.err
python - <<'P'
import json
try:
    d=json.load(open('gpurun_out/bench_2gpu_ra.json')); print('2gpu_ra', round(d['value'],2), round(d['e2e']['value'],2), d['host_ms_per_gof'], d.get('exchange'))
except Exception as e: print('2gpu_ra ERR', e)
P
tail -5 gpurun_out/bench_2gpu_ra.err
