#!/bin/bash
mkdir -p gpurun_out
python scratch/bench_hub.py 128 > gpurun_out/r2_hub_ab.log 2>&1; tail -12 gpurun_out/r2_hub_ab.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"k_hub|k_spmm_lean" -c 60 --csv --log-file gpurun_out/r2_hub_launches.csv python scratch/bench_hub.py 128 > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2_hub_launches.csv", errors="ignore")))
hdr = None
out = {}
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    key = (d["ID"], d["Kernel Name"][:40], d["Grid Size"])
    out.setdefault(key, {})[d["Metric Name"]] = d["Metric Value"]
for k, v in list(out.items())[:40]:
    print(k, {a.split("__")[-1][:24]: b for a, b in v.items()})
PY
