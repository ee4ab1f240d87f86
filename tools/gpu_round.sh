#!/bin/bash
# One gpurun call (1 GPU): the bench configurations, launch list with DRAM traffic, full ncu captures of one launch of each heavier kernel
mkdir -p gpurun_out
summ() { python - "$1" <<'P'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1])); print(sys.argv[1], round(d['value'],2), round(d['e2e']['value'],2), d.get('parity_ok'), d['host_ms_per_gof'], {k:v for k,v in list(d['stage_ms_per_frame'].items())[:7]})
except Exception as e: print(sys.argv[1],'ERR', e)
P
}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; summ default
kill $SMI
timeout 300 python bench.py --steps 4 --warmup 1 --gofs-in-flight 1 --no-cpu-baseline > gpurun_out/bench_1gof.json 2> gpurun_out/bench_1gof.err; summ 1gof
timeout 600 python bench.py --condition ra --steps 32 > gpurun_out/bench_ra_r5.json 2> gpurun_out/bench_ra_r5.err; summ ra_r5
timeout 900 python bench.py --bits 11 --frames 8 --gofs-in-flight 8 --scratch-sets 8 --steps 16 --warmup 2 > gpurun_out/bench_vox11.json 2> gpurun_out/bench_vox11.err; summ vox11
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_traffic.csv \
  python bench.py --frames 1 --steps 1 --warmup 0 --gofs-in-flight 1 --no-cpu-baseline > gpurun_out/ncu_list.out 2> gpurun_out/ncu_list.err
echo "ncu list rc=$?"; wc -l gpurun_out/launches_traffic.csv
# the first launch of each of the heavier kernels, full metric set (no source import: the report has to stay small)
timeout 600 ncu --set full --clock-control none --kernel-id '::regex:kSweepTail|kSweepStatic|kKnn|kAdjacency|kLocalSubtrees|kRecountAndActivate|kHook|kEdgeKeys|DeviceRadixSortOnesweep|kPropagateEdges|kCrossEdges|kApply2Assign|kScanSweep1|kNormals:1' -c 16 -o gpurun_out/prof_r02h -f \
  python bench.py --frames 1 --steps 1 --warmup 0 --gofs-in-flight 1 --iterations 2 --no-cpu-baseline > gpurun_out/ncu_full.out 2> gpurun_out/ncu_full.err
echo "ncu full rc=$?"; ls -la gpurun_out/prof_r02h.ncu-rep
find gpurun_out -size +45M -name '*.ncu-rep' -delete
du -sh gpurun_out
