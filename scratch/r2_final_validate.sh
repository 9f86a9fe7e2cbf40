#!/bin/bash
# final validation of the round: whole -m gpu suite, default N=1 bench line, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_tests_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_final.log
tail -9 gpurun_out/r2_tests_final.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.log 2> gpurun_out/r2_bench_final.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_final.log").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "roofline", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "step frac", d["roofline_step"]["frac"], "e2e", d["e2e"]["ms_per_step"], "launches", d["gpu_launches_per_step"])
print("gat", [(g.get("layers"), g.get("hidden"), g.get("ms_per_step"), g.get("roofline_step", {}).get("frac"), g.get("error")) for g in d.get("gat", [])])
print("parity", d.get("parity"), "cpu", d.get("cpu_baseline", {}).get("value"))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-200
