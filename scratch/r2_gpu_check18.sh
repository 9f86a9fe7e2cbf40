#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gemm_tc_gpu.py tests/test_mlp_gpu.py tests/test_sage_gpu.py tests/test_oracle_golden.py -m gpu -q > gpurun_out/r2_tests8.log 2>&1
tail -6 gpurun_out/r2_tests8.log
python scratch/timeline.py SAGE 128 2 > gpurun_out/r2_timeline_sage2.log 2>&1
grep -v Warn gpurun_out/r2_timeline_sage2.log | grep "ms/step\|span\|rror"
KGB_GEMM_ROWS_MIN_M=65536 python scratch/timeline.py SAGE 128 2 2>&1 | grep "ms/step"
