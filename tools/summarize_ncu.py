#!/usr/bin/env python
"""Turn ncu CSV exports into the small summaries committed under profiles/.

    python tools/summarize_ncu.py launches <launches.csv> [--bench bench.json]   -> per-kernel time shares
    python tools/summarize_ncu.py raw <prof_raw.csv>                             -> per-kernel --set full digest
"""
import collections
import csv
import json
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("tdg::", "")


def launches(path, bench=None):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    col = {n: i for i, n in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        a = agg.setdefault(short(r[col["Kernel Name"]]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list: {path}\n")
    print("per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
    print("| kernel | launches | total ms (ncu) | share |")
    print("|---|---|---|---|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ms:.3f} | {ms / tot:.1%} |")
    print(f"| total | {sum(a[0] for a in agg.values())} | {tot:.3f} | |")
    if bench:
        d = json.loads(open(bench).read().strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        t = sum(v["ms_per_step"] for v in pk.values())
        print(f"\nbench.py CUDA-event shares of the same workload ({bench}, {d['ms_per_step']:.2f} ms/step):\n")
        print("| kernel class | ms/step (events) | share | achieved GB/s |")
        print("|---|---|---|---|")
        for k, v in pk.items():
            print(f"| `{k}` | {v['ms_per_step']:.3f} | {v['ms_per_step'] / t:.1%} | {v['achieved_gbs']:.0f} |")


WANT = [
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp inst"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def raw(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    col = {n: i for i, n in enumerate(hdr)}
    seen = collections.OrderedDict()
    for r in data:
        if len(r) < len(hdr):
            continue
        vals = []
        for w, _ in WANT:
            try:
                vals.append((float(r[col[w]].replace(",", "")), units[col[w]]))
            except (KeyError, ValueError):
                vals.append((float("nan"), ""))
        seen.setdefault(short(r[col["Kernel Name"]]), []).append((r[col["Grid Size"]], vals))
    print(f"# ncu --set full digest: {path}\n")
    print("largest launch of each kernel; traffic = dram rd + dram wr of that launch\n")
    print("| kernel | grid | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|---|" + "---|" * len(WANT))
    for k, l in seen.items():
        def secs(t):
            v, u = t[1][0]
            return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}.get(u, 1.0)
        l.sort(key=lambda t: -secs(t))
        g, vals = l[0]
        print(f"| `{k}` | {g} | " + " | ".join(f"{v:.4g} {u}".strip() for v, u in vals) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        b = sys.argv[sys.argv.index("--bench") + 1] if "--bench" in sys.argv else None
        launches(sys.argv[2], b)
    else:
        raw(sys.argv[2])
