#!/bin/bash
mkdir -p gpurun_out
python scratch/timeline.py SAGE 128 2 > gpurun_out/r2_timeline_sage.log 2>&1
grep -v Warn gpurun_out/r2_timeline_sage.log | tail -40
