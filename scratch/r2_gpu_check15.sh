#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gat_gpu.py tests/test_kernels_gpu.py tests/test_oracle_golden.py tests/test_oracle_microcases.py tests/test_zz_sharded_gat_gpu.py tests/test_kgwas_gpu.py -m gpu -q -x > gpurun_out/r2_tests6.log 2>&1
tail -5 gpurun_out/r2_tests6.log
python scratch/timeline.py GAT 128 2 > gpurun_out/r2_timeline_gat3.log 2>&1
grep -v Warn gpurun_out/r2_timeline_gat3.log | grep "ms/step\|span\|rror"
python scratch/timeline.py GAT 256 3 > gpurun_out/r2_timeline_gat4.log 2>&1
grep -v Warn gpurun_out/r2_timeline_gat4.log | grep "ms/step\|span\|rror"
