#!/bin/bash
# gpurun --gpus 2: the frame-sharded bench over NCCL (canvas exchange off the critical path), normally and with the host share an
# 8-rank run leaves each rank (16 cores / 8 ranks = 2 cores per rank)
mkdir -p gpurun_out
summ() { python - "$1" <<'P'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1])); print(sys.argv[1], round(d['value'],2), round(d['e2e']['value'],2), d['host_ms_per_gof'], d.get('exchange'))
except Exception as e: print(sys.argv[1],'ERR', e)
P
}
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; summ 2gpu
taskset -c 0-3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_4cores.json 2> gpurun_out/bench_2gpu_4cores.err; summ 2gpu_4cores
tail -3 gpurun_out/bench_2gpu.err
