#!/bin/bash
# full validation: whole -m gpu suite, the default N=1 bench line, the reference arm, an ncu launch list of one replayed step
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2_tests_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_full.log
tail -12 gpurun_out/r2_tests_full.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_full.log 2> gpurun_out/r2_bench_full.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_full.log").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "roofline", d["roofline"]["frac"], "step frac", d["roofline_step"]["frac"], "e2e", d["e2e"]["ms_per_step"])
print("gat", [(g.get("layers"), g.get("hidden"), g.get("ms_per_step"), g.get("roofline_step", {}).get("frac"), g.get("error")) for g in d.get("gat", [])])
print("parity", d.get("parity"), "cpu", d.get("cpu_baseline", {}).get("value"))
PY
python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-300
