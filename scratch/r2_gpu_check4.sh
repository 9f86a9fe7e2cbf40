#!/bin/bash
mkdir -p gpurun_out
python scratch/debug_mid3.py > gpurun_out/r2_debug_mid3.log 2>&1
tail -40 gpurun_out/r2_debug_mid3.log
CUDA_LAUNCH_BLOCKING=1 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_hub_gpu.py::test_hub_tiles_match_pull_kernel_and_fp64" -q -x -k "4099" > gpurun_out/r2_sanit.log 2>&1
grep -v "^$" gpurun_out/r2_sanit.log | head -60
