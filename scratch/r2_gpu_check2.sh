#!/bin/bash
mkdir -p gpurun_out
python scratch/debug_mid_grads.py > gpurun_out/r2_debug_mid.log 2>&1
tail -50 gpurun_out/r2_debug_mid.log
python -m pytest tests -m gpu -q --durations=8 --deselect tests/test_oracle_golden.py::test_cuda_matches_reference_at_bench_widths > gpurun_out/r2_tests2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests2.log
tail -40 gpurun_out/r2_tests2.log
