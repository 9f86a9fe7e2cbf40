#!/bin/bash
mkdir -p gpurun_out
for w in 4 16 8; do
KGB_SPMM_WAVES=$w timeout 100 python bench.py --steps 20 --warmup 5 --no-gat --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2_waves_$w.log 2>/dev/null
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_waves_$w.log").read().strip().splitlines()[-1])
print("waves=$w ms/step", round(d["ms_per_step"], 4), "roofline", round(d["roofline"]["frac"], 4))
PY
done
