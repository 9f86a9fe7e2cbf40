#!/bin/bash
echo "--- 8 producer warps"; python scratch/bench_gemm.py --big
echo "--- 12 producer warps"; KGB_GEMM_PROD_WARPS=12 python scratch/bench_gemm.py --big
KGB_GEMM_PROD_WARPS=12 python -m pytest tests/test_gemm_tc_gpu.py -q -k "row_streaming" 2>&1 | tail -2
