#!/bin/bash
mkdir -p gpurun_out
python scratch/debug_mid6.py > gpurun_out/r2_debug_mid6.log 2>&1
grep -v Warning gpurun_out/r2_debug_mid6.log | tail -50
