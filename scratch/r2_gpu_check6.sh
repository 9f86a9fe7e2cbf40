#!/bin/bash
mkdir -p gpurun_out
python scratch/debug_mid5.py > gpurun_out/r2_debug_mid5.log 2>&1
grep -v Warning gpurun_out/r2_debug_mid5.log | tail -30
python -m pytest tests/test_hub_gpu.py -q -x > gpurun_out/r2_hub.log 2>&1
tail -5 gpurun_out/r2_hub.log
