#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gemm_tc_gpu.py tests/test_mlp_gpu.py -m gpu -q > gpurun_out/r2_tests9.log 2>&1
grep -n "^FAILED\|passed\|failed\|AssertionError: " gpurun_out/r2_tests9.log | head -20
