#!/bin/bash
# N-GPU bench line (strong headline + weak block + parity), tight timeout
N=$1
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.log 2> gpurun_out/r2_bench_n$N.err
echo "n=$N rc=$?"
grep -v "^$" gpurun_out/r2_bench_n$N.err | grep -iv "warn\|\*\*\*\|OMP_NUM" | tail -5
python - <<PY
import json
ls = [l for l in open("gpurun_out/r2_bench_n$N.log").read().strip().splitlines() if l.startswith("{")]
if ls:
    d = json.loads(ls[-1])
    print("N=$N", d["scaling"], "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d.get("e2e", {}).get("ms_per_step"), d.get("e2e", {}).get("value"))
    print("parity", {k: d.get("parity", {}).get(k) for k in ("err", "grad_err", "n_grads_none_vs_zero")})
    w = d.get("weak") or {}
    print("weak ms", w.get("ms_per_step"), "value", w.get("value"), "e2e", (w.get("e2e") or {}).get("value"))
PY
