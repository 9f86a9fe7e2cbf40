#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_sampler_gpu.py tests/test_mlp_gpu.py tests/test_kgwas_gpu.py tests/test_oracle_golden.py -m gpu -q -x > gpurun_out/r2_tests7.log 2>&1
tail -8 gpurun_out/r2_tests7.log
timeout 600 python scratch/bench_minibatch.py GAT 8 > gpurun_out/r2_minibatch_gat.log 2>&1
tail -3 gpurun_out/r2_minibatch_gat.log
