#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.log 2> gpurun_out/r2_bench_n2.err
echo "n2 rc=$?"
grep -v "^$" gpurun_out/r2_bench_n2.err | grep -iv "warn\|\*\*\*\|OMP_NUM" | tail -8
python - <<'PY'
import json
ls = [l for l in open("gpurun_out/r2_bench_n2.log").read().strip().splitlines() if l.startswith("{")]
if ls:
    d = json.loads(ls[-1])
    print("2 GPU", d["scaling"], "ms/step", d["ms_per_step"], "value", d["value"], "parity", d.get("parity"), "e2e", d.get("e2e", {}).get("ms_per_step"))
    print("weak", d.get("weak"))
    print(d["cuda_graph"])
PY
