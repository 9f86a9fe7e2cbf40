#!/bin/bash
echo "--- default dispatch"; python scratch/bench_gemm.py
echo "--- KGB_GEMM_ROWS_MIN_M=4096"; KGB_GEMM_ROWS_MIN_M=4096 python scratch/bench_gemm.py
