#!/bin/bash
# One gpurun call (1 GPU): compute-sanitizer racecheck (shared-memory hazards) over the tests that run the shared-memory kernels of this
# round: in-CTA kd-subtrees, block scans, push-pull pyramid tail, adjacency sort, packing
mkdir -p gpurun_out
( timeout 800 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python -m pytest tests/test_gpu_knn_normals.py "tests/test_gpu_gof.py::test_encode_gof_all_products" -m gpu -x -q ) > gpurun_out/racecheck_r02.log 2>&1
tail -12 gpurun_out/racecheck_r02.log; grep -c "hazard" gpurun_out/racecheck_r02.log; true
