#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2_tests4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests4.log
tail -15 gpurun_out/r2_tests4.log
python bench.py --steps 20 --warmup 5 --no-gat --no-cpu-baseline > gpurun_out/r2_bench2.log 2> gpurun_out/r2_bench2.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench2.log").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "roofline", d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], "step frac", d["roofline_step"]["frac"], "e2e", d.get("e2e", {}).get("ms_per_step"))
PY
KGB_SPMM_HUB=0 python bench.py --steps 20 --warmup 5 --no-gat --no-cpu-baseline --no-e2e > gpurun_out/r2_bench2_nohub.log 2> gpurun_out/r2_bench2_nohub.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench2_nohub.log").read().strip().splitlines()[-1])
print("NO HUB ms/step", d["ms_per_step"], "roofline", d["roofline"]["frac"], d["roofline"]["avg_launch_ms"])
PY
python scratch/bench_hub.py 128 > gpurun_out/r2_hub_ab.log 2>&1; tail -12 gpurun_out/r2_hub_ab.log
