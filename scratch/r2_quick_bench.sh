#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --no-gat --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2_bench_quick.log 2> gpurun_out/r2_bench_quick.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_quick.log").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "launches/step", d["gpu_launches_per_step"], "roofline", d["roofline"]["frac"])
PY
timeout 200 python -m pytest tests/test_sage_gpu.py tests/test_oracle_golden.py -m gpu -q 2>&1 | tail -2
