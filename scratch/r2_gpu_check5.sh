#!/bin/bash
mkdir -p gpurun_out
python scratch/debug_mid4.py > gpurun_out/r2_debug_mid4.log 2>&1
CUDA_LAUNCH_BLOCKING=1 python scratch/debug_mid4.py >> gpurun_out/r2_debug_mid4.log 2>&1
grep "MULTI_STREAM\|Error" gpurun_out/r2_debug_mid4.log
compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_hub_gpu.py -q -x > gpurun_out/r2_sanit.log 2>&1
grep -v "Host Frame" gpurun_out/r2_sanit.log | grep -v "^$" | tail -30
