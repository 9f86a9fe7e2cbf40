#!/bin/bash
# round-2 evidence: launch list of one bench step (graph off, so every launch is visible), ncu --set full of the
# gather-reduce launches (DRAM traffic), tensor-pipe / DRAM metrics of the tcgen05 GEMMs, smoke()
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_sage.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-gat --no-parity --no-cuda-graph --profile-range > gpurun_out/r2_ncu_launches.log 2>&1; tail -c 300 gpurun_out/r2_ncu_launches.log | head -3
timeout 500 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:k_spmm -o gpurun_out/r2_spmm_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-gat --no-parity --no-cuda-graph --profile-range > gpurun_out/r2_ncu_full.log 2>&1; tail -2 gpurun_out/r2_ncu_full.log
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --profile-from-start off -k regex:k_gemm_tc --csv --log-file gpurun_out/r2_gemm_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-gat --no-parity --no-cuda-graph --profile-range > gpurun_out/r2_ncu_gemm.log 2>&1; tail -c 200 gpurun_out/r2_ncu_gemm.log | head -2
python __graft_entry__.py smoke 2>&1 | tail -3
