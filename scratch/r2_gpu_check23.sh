#!/bin/bash
mkdir -p gpurun_out
for cp in 0 1; do
KGB_CHAIN_PRIORITY=$cp timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$cp bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong --no-e2e --no-parity > gpurun_out/r2_bench_n2_cp$cp.log 2> gpurun_out/r2_bench_n2_cp$cp.err
echo "chain_priority=$cp rc=$?"
python - <<PY
import json
ls = [l for l in open("gpurun_out/r2_bench_n2_cp$cp.log").read().strip().splitlines() if l.startswith("{")]
if ls:
    d = json.loads(ls[-1]); print("  ms/step", d["ms_per_step"], "value", d["value"])
PY
done
