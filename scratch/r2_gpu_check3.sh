#!/bin/bash
mkdir -p gpurun_out
python scratch/debug_mid2.py > gpurun_out/r2_debug_mid2.log 2>&1
tail -12 gpurun_out/r2_debug_mid2.log
python -m pytest tests/test_hub_gpu.py tests/test_oracle_microcases.py tests/test_kernels_gpu.py tests/test_sage_gpu.py -m gpu -q -x > gpurun_out/r2_tests3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests3.log
tail -40 gpurun_out/r2_tests3.log
