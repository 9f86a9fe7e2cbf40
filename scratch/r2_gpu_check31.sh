#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_dist_gpu.py tests/test_sage_gpu.py -q -x 2>&1 | tail -2
bash scratch/r2_scale.sh 2
