#!/bin/bash
# One gpurun call (1 GPU): compute-sanitizer memcheck over the GPU tests that exercise the kernels written this round (kd-tree build
# with look-back scans + in-CTA subtrees, refine sweeps, cross-edge propagation, push-pull tail, sharded random-access packing, smoothing)
mkdir -p gpurun_out
( timeout 700 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_knn_normals.py tests/test_gpu_segment.py tests/test_gpu_gof.py tests/test_gpu_postrecon.py tests/test_ra_pack.py -m gpu -x -q -k "not full_size and not vs_reference and not decoder_side_binding" ) > gpurun_out/memcheck_r02.log 2>&1
tail -8 gpurun_out/memcheck_r02.log; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck_r02.log
