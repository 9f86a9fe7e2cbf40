#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-gat --no-cpu-baseline --no-e2e > gpurun_out/r2_bench3.log 2> gpurun_out/r2_bench3.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench3.log").read().strip().splitlines()[-1])
print("1 GPU ms/step", d["ms_per_step"], "roofline", d["roofline"]["frac"], "step frac", d["roofline_step"]["frac"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.log 2> gpurun_out/r2_bench_n2.err
echo "n2 rc=$?"
tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n2.log").read().strip().splitlines() if l.startswith("{")][-1])
print("2 GPU strong ms/step", d["ms_per_step"], "value", d["value"], "parity", d.get("parity"), "e2e", d.get("e2e", {}).get("ms_per_step"))
print("weak", d.get("weak"))
print(d["cuda_graph"])
PY
python -m pytest tests/test_dist_gpu.py tests/test_zz_sharded_gat_gpu.py -q 2>&1 | tail -3
