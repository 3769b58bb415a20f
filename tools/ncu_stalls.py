#!/usr/bin/env python
"""Stall-reason / throughput digest of an `ncu --page raw --csv` export: for the largest launch of each
kernel, the warp-state breakdown (cycles per issued instruction by reason) and the headline rates."""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; col = {n: i for i, n in enumerate(hdr)}
kn = col["Kernel Name"]
best = {}
for r in rows[hi + 2:]:
    if len(r) < len(hdr): continue
    name = re.sub(r"\(.*", "", re.sub(r"^void\s+", "", r[kn])).replace("tdg::", "")
    try: t = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
    except Exception: continue
    if name not in best or t > best[name][0]: best[name] = (t, r)
stall = [c for c in hdr if c.startswith("smsp__average_warp") and "issue_stalled" in c and c.endswith("_per_warp_active.pct")]
if not stall:
    stall = [c for c in hdr if "warps_issue_stalled" in c and c.endswith(".ratio")]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
for name, (t, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
    print(f"## {name}")
    print("  " + "  ".join(f"{k.split('__')[-1].split('.')[0]}={r[col[k]]}" for k in keys if k in col))
    ss = []
    for c in stall:
        try: ss.append((float(r[col[c]].replace(",", "")), c))
        except Exception: pass
    for v, c in sorted(ss, reverse=True)[:7]:
        print(f"    {v:8.2f}  {c.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '')}")
