#!/usr/bin/env python
"""profiles/ncu_traffic.json: measured DRAM traffic (ncu --set full, dram__bytes_read.sum +
dram__bytes_write.sum) of every pass kernel class over ONE bench step, beside the algorithmic bytes
bench.py counted for the same step, and their ratio.  bench.py scales the ratio to its own launch size
for `roofline.traffic`.

    python tools/make_traffic.py <prof_raw.csv> <log holding the bench JSON line of the same run> <out.json>
"""
import csv
import json
import re
import sys

CLASS = [(r"k_voice<\(?(?:int\))?0", "k_voice<debug_sine>"), (r"k_voice<\(?(?:int\))?1", "k_voice<synth>"),
         (r"k_voice<\(?(?:int\))?2", "k_voice<sampsyn>"), (r"k_fx<\(?(?:int\))?0", "k_fx<sum>"),
         (r"k_fx<\(?(?:int\))?1", "k_fx<adsr>"), (r"k_fx<\(?(?:int\))?2", "k_fx<norm_peak>")]


def classify(name):
    name = name.replace("tdg::", "").replace("void ", "")
    for pat, cls in CLASS:
        if re.search(pat, name):
            return cls
    return re.sub(r"[<(].*", "", name)


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main(raw, log, out):
    rows = list(csv.reader(open(raw)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    col = {n: i for i, n in enumerate(hdr)}
    rd, wr, tm = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"], col["gpu__time_duration.sum"]
    agg = {}
    for r in data:
        if len(r) < len(hdr):
            continue
        a = agg.setdefault(classify(r[col["Kernel Name"]]), {"launches": 0, "dram_bytes": 0.0})
        a["launches"] += 1
        a["dram_bytes"] += to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
    bench = None
    for ln in open(log):
        if ln.startswith("{") and '"roofline"' in ln:
            bench = json.loads(ln)
    steps = bench["steps"]
    res = {"_source": {"ncu_raw": raw, "bench_line_of_same_run": {"projects_per_gpu": bench["config"]["projects_per_gpu"], "steps": steps}},
           "_note": "ratio = measured DRAM bytes / algorithmic bytes of the same step; < 1 where gathers hit L2 or a pass is fused"}
    for k, v in bench["roofline"]["per_kernel"].items():
        if k not in agg:
            continue
        algo = v["achieved_gbs"] * 1e9 * v["ms_per_step"] * 1e-3
        res[k] = {"launches_per_step": agg[k]["launches"], "dram_bytes_per_step": agg[k]["dram_bytes"],
                  "algorithmic_bytes_per_step": algo, "ratio": agg[k]["dram_bytes"] / algo if algo else None}
    json.dump(res, open(out, "w"), indent=1)
    for k, v in res.items():
        if not k.startswith("_"):
            print(f"{k:22s} dram {v['dram_bytes_per_step'] / 1e9:8.3f} GB  algo {v['algorithmic_bytes_per_step'] / 1e9:8.3f} GB  ratio {v['ratio'] if v['ratio'] is None else round(v['ratio'], 3)}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
