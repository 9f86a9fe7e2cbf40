#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gat_gpu.py tests/test_kernels_gpu.py tests/test_oracle_golden.py tests/test_oracle_microcases.py -m gpu -q -x > gpurun_out/r2_tests5.log 2>&1
tail -5 gpurun_out/r2_tests5.log
python scratch/timeline.py GAT 128 2 > gpurun_out/r2_timeline_gat2.log 2>&1
grep -v Warn gpurun_out/r2_timeline_gat2.log | grep "ms/step\|span"
python - <<'PY'
import csv, collections
rows=list(csv.DictReader(open('gpurun_out/timeline_GAT_h128_L2.csv')))
tot=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    n=r['name'].replace('void ','').replace('kgb::','').split('(')[0][:50]
    tot[n][0]+=1; tot[n][1]+=float(r['dur_us'])
for k,v in sorted(tot.items(), key=lambda kv:-kv[1][1])[:14]:
    print(f"{v[1]:9.0f} us n={v[0]:4d} {k}")
PY
