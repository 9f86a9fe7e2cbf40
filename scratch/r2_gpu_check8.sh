#!/bin/bash
mkdir -p gpurun_out
python scratch/debug_mid7.py > gpurun_out/r2_debug_mid7.log 2>&1
grep -v Warning gpurun_out/r2_debug_mid7.log | sed -n '/=====/,$p' | tail -40
