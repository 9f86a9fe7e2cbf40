#!/bin/bash
mkdir -p gpurun_out
KGB_CHAIN_PRIORITY=1 python scratch/timeline.py SAGE 128 2 > gpurun_out/r2_timeline_sage_cp.log 2>&1
mv gpurun_out/timeline_SAGE_h128_L2.csv gpurun_out/timeline_SAGE_h128_L2_chainprio.csv
grep -v Warn gpurun_out/r2_timeline_sage_cp.log | grep "ms/step\|span"
python scratch/timeline.py GAT 128 2 > gpurun_out/r2_timeline_gat.log 2>&1
grep -v Warn gpurun_out/r2_timeline_gat.log | grep "ms/step\|span"
