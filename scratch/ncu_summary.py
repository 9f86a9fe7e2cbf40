"""Summarise an ncu --set full report (raw page) per launch: time, DRAM bytes, L2/L1 hit, stalls."""
import csv, subprocess, sys, json
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
def col(k): return hdr.index(k) if k in hdr else None
keys = {"time_us": "gpu__time_duration.sum", "grid": "launch__grid_size", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l2_hit_pct": "lts__t_sector_hit_rate.pct", "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
        "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "regs": "launch__registers_per_thread", "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"}
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
res = []
for r in rows[2:]:
    d = {"kernel": r[col("Kernel Name")][:60]}
    for k, m in keys.items():
        c = col(m)
        if c is None: continue
        d[k] = to_bytes(r[c], units[c]) if k.startswith("dram_r") or k.startswith("dram_w") else float(r[c].replace(",", "")) * (1e-3 if units[c] == "ns" and k == "time_us" else 1)
    res.append(d)
json.dump(res, sys.stdout, indent=1)
