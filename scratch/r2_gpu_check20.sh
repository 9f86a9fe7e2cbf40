#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_dist_gpu.py tests/test_zz_sharded_gat_gpu.py -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2b.log 2> gpurun_out/r2_bench_n2b.err
echo "n2 rc=$?"
grep -v "^$" gpurun_out/r2_bench_n2b.err | grep -iv "warn\|\*\*\*\|OMP_NUM" | tail -6
python - <<'PY'
import json
ls = [l for l in open("gpurun_out/r2_bench_n2b.log").read().strip().splitlines() if l.startswith("{")]
if ls:
    d = json.loads(ls[-1])
    print("2 GPU", d["scaling"], "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d.get("e2e", {}).get("ms_per_step"))
    print("parity", d.get("parity"))
    w = d.get("weak") or {}
    print("weak ms", w.get("ms_per_step"), "e2e", (w.get("e2e") or {}).get("ms_per_step"))
PY
