#!/bin/bash
mkdir -p gpurun_out
for m in 1 0; do
KGB_MERGE_XF=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$m bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong --no-e2e --no-parity > gpurun_out/r2_bench_n2_m$m.log 2> gpurun_out/r2_bench_n2_m$m.err
echo "merge=$m rc=$?"
python - <<PY
import json
ls = [l for l in open("gpurun_out/r2_bench_n2_m$m.log").read().strip().splitlines() if l.startswith("{")]
if ls:
    d = json.loads(ls[-1]); print("  ms/step", d["ms_per_step"], "value", d["value"])
PY
done
