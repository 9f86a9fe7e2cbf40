#!/bin/bash
# round 2, first GPU call: the whole -m gpu suite, then the N=1 bench line (SAGE headline + GAT block + parity)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests.log
tail -30 gpurun_out/r2_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.log 2> gpurun_out/r2_bench1.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench1.log
tail -5 gpurun_out/r2_bench1.err
