"""ncu raw page (csv) of every kgb_spmm launch of ONE bench step -> profiles/<out>.json (+ a table on stdout).
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python scratch/ncu_traffic.py raw.csv "<source note>" [out.json] """
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def val(d, name):
    v = float(d[ix[name]].replace(",", ""))
    u = units[ix[name]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)


tot = 0.0
print("| # | kernel | grid | us | DRAM read MB | DRAM write MB | L2->L1 MB | L1 hit % | L2 hit % | lts % | warp-instr M | issue % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for k, d in enumerate(data):
    rd, wr = val(d, "dram__bytes_read.sum"), val(d, "dram__bytes_write.sum")
    tot += rd + wr
    print(f"| {k} | {d[ix['Kernel Name']][:34]} | {d[ix['Grid Size']]} | {val(d, 'gpu__time_duration.sum'):.1f} | {rd / 1e6:.0f} | "
          f"{wr / 1e6:.0f} | {val(d, 'l1tex__m_xbar2l1tex_read_bytes.sum') / 1e6:.0f} | {float(d[ix['l1tex__t_sector_hit_rate.pct']]):.1f} | "
          f"{float(d[ix['lts__t_sector_hit_rate.pct']]):.1f} | {float(d[ix['lts__throughput.avg.pct_of_peak_sustained_elapsed']]):.1f} | "
          f"{float(d[ix['smsp__inst_executed.sum']].replace(',', '')) / 1e6:.1f} | "
          f"{float(d[ix['smsp__issue_active.avg.pct_of_peak_sustained_active']]):.1f} |")
out = {"hidden": 128, "backbone": "SAGE", "launches": len(data), "dram_bytes_per_step": tot,
       "source": sys.argv[2] if len(sys.argv) > 2 else "ncu --set full, one bench step"}
json.dump(out, open(os.path.join(ROOT, "profiles", sys.argv[3] if len(sys.argv) > 3 else "r02_spmm_traffic.json"), "w"), indent=1)
print(f"\ntotal DRAM bytes of {len(data)} launches: {tot / 1e9:.3f} GB")
