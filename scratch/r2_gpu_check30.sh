#!/bin/bash
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 scratch/timeline.py SAGE 128 2 > gpurun_out/r2_timeline_n8.log 2>&1
echo rc=$?
grep -v Warn gpurun_out/r2_timeline_n8.log | grep "ms/step\|span" | head -3 | cut -c1-200
