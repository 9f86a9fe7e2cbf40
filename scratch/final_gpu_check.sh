mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log); tail -c 800 gpurun_out/tests.log
timeout 400 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; tail -c 400 gpurun_out/bench_final.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_final.log').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['roofline']['frac'],d['roofline']['measured_in'],d['roofline_step']['frac'],d['e2e']['ms_per_step'],d['cuda_graph'][:40],d['cpu_baseline'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile-range > gpurun_out/b_ncu.log 2>&1; tail -c 200 gpurun_out/b_ncu.log | head -3
timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:k_spmm -o gpurun_out/spmm_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-cuda-graph --profile-range > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
python __graft_entry__.py smoke 2>&1 | tail -3
