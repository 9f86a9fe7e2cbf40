#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scratch/timeline.py SAGE 128 2 > gpurun_out/r2_timeline_n2.log 2>&1
grep -v Warn gpurun_out/r2_timeline_n2.log | grep "ms/step\|span\|gap" | cut -c1-400
