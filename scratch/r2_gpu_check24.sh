#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_sage_gpu.py tests/test_oracle_golden.py tests/test_zz_fullsize_oracle_gpu.py tests/test_kgwas_gpu.py tests/test_dist_gpu.py tests/test_mlp_gpu.py -m gpu -q -x > gpurun_out/r2_tests10.log 2>&1
tail -4 gpurun_out/r2_tests10.log
for m in 1 0; do
KGB_MERGE_XF=$m python bench.py --steps 20 --warmup 5 --no-gat --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2_bench_merge$m.log 2> gpurun_out/r2_bench_merge$m.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_merge$m.log").read().strip().splitlines()[-1])
print("merge=$m ms/step", d["ms_per_step"], "launches/step", d["gpu_launches_per_step"], "roofline", d["roofline"]["frac"])
PY
done
