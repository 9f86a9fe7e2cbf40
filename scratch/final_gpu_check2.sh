mkdir -p gpurun_out

timeout 400 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_final.log').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['roofline']['frac'],d['roofline']['kernel_share_of_step'],d['roofline_step']['frac'],d['e2e']['ms_per_step'],d['cuda_graph'][:40],d['gpu_launches_per_step'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile-range > gpurun_out/b_ncu.log 2>&1; tail -c 100 gpurun_out/b_ncu.log
