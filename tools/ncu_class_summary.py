#!/usr/bin/env python
"""Per kernel class summary of `ncu --set full` raw CSV exports (ncu -i X.ncu-rep --page raw --csv):

    python tools/ncu_class_summary.py <out.json> <out.md> <label> <raw1.csv> [<raw2.csv> ...] [--bench <log with the JSON line of the SAME run>]

With --bench (the capture must be `bench.py --warmup 0 --steps 1 ...` under ncu, every library kernel captured): only the
launches of the first (timed) step are used, and each class also gets `ratio` = measured DRAM bytes / the algorithmic
bytes bench.py counted for the same launches — what bench.py scales to its own launch size for `roofline.traffic`.

For every class (k_voice<synth>, k_fx<adsr>, ...) over all captured launches: measured DRAM bytes (dram__bytes_read.sum
+ dram__bytes_write.sum), kernel time, DRAM GB/s = bytes / time and its fraction of the measured HBM peak, issue-slot
utilisation and pipe utilisations (time-weighted).  bench.py reads the JSON for `roofline.per_kernel[*].dram_frac_ncu`
and `issue_slot_utilisation_pct_ncu`: fractions from DRAM traffic, not from algorithmic bytes (a gather that hits L2
must not read as > 1.0 of HBM)."""
import csv, json, re, sys, os

CLASS = [(r"k_voice<\(?(?:int\))?0", "k_voice<debug_sine>"), (r"k_voice<\(?(?:int\))?1", "k_voice<synth>"),
         (r"k_voice<\(?(?:int\))?2", "k_voice<sampsyn>"), (r"k_fx<\(?(?:int\))?0", "k_fx<sum>"),
         (r"k_fx<\(?(?:int\))?1", "k_fx<adsr>"), (r"k_fx<\(?(?:int\))?2", "k_fx<norm_peak>")]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
        "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def classify(name):
    name = name.replace("tdg::", "").replace("void ", "")
    for pat, cls in CLASS:
        if re.search(pat, name):
            return cls
    return re.sub(r"[<(].*", "", name)


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return 0.0


def main(out_json, out_md, label, raws, bench_log=None):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bench = None
    if bench_log:
        for ln in open(bench_log, errors="replace"):
            if ln.startswith("{") and '"roofline"' in ln:
                bench = json.loads(ln)
    first_n = (bench["gpu_launches"] // bench["steps"]) if bench else None
    peak = 6532.9
    try:
        peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError, ValueError):
        pass
    agg = {}
    pct = {"issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "fma": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fp64": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lsu": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l2_hit": "lts__t_sector_hit_rate.pct", "l1_hit": "l1tex__t_sector_hit_rate.pct"}
    for raw in raws:
        rows = list(csv.reader(open(raw, errors="replace")))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        if "Metric Name" in rows[hi]:
            # long format (`ncu --metrics ... --csv --log-file`): one row per (launch, metric) -> pivot to the wide raw-page form
            c0 = {n: i for i, n in enumerate(rows[hi])}
            names, wide, unit_of = [], {}, {}
            for r in rows[hi + 1:]:
                if len(r) < len(rows[hi]):
                    continue
                lid = r[c0["ID"]]
                w = wide.setdefault(lid, {"Kernel Name": r[c0["Kernel Name"]]})
                m = r[c0["Metric Name"]]
                if m not in unit_of:
                    unit_of[m] = r[c0["Metric Unit"]]; names.append(m)
                w[m] = r[c0["Metric Value"]]
            hdr = ["ID", "Kernel Name"] + names
            units = ["", ""] + [unit_of[m] for m in names]
            data = [[lid, w["Kernel Name"]] + [w.get(m, "0") for m in names] for lid, w in sorted(wide.items(), key=lambda kv: int(kv[0]))]
        else:
            hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
        col = {n: i for i, n in enumerate(hdr)}
        rd, wr, tm = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"], col["gpu__time_duration.sum"]
        if first_n is not None:
            data = [r for r in data if len(r) >= len(hdr)][:first_n]
        for r in data:
            if len(r) < len(hdr):
                continue
            a = agg.setdefault(classify(r[col["Kernel Name"]]), {"launches": 0, "dram_bytes": 0.0, "seconds": 0.0, "warp_inst": 0.0,
                                                                 "regs": 0, **{k: 0.0 for k in pct}})
            t = num(r[tm]) * UNIT.get(units[tm], 1e-9)
            a["launches"] += 1
            a["dram_bytes"] += num(r[rd]) * UNIT.get(units[rd], 1) + num(r[wr]) * UNIT.get(units[wr], 1)
            a["seconds"] += t
            if "smsp__inst_executed.sum" in col:
                a["warp_inst"] += num(r[col["smsp__inst_executed.sum"]])
            if "launch__registers_per_thread" in col:
                a["regs"] = max(a["regs"], int(num(r[col["launch__registers_per_thread"]])))
            for k, m in pct.items():
                if m in col:
                    a[k] += num(r[col[m]]) * t
    res = {"_label": label, "_peak_gbs": peak,
           "_note": "ncu --set full --clock-control none; times are cold-cache and serialised (profiler), so use the FRACTIONS; "
                    "dram_frac = (dram read + write bytes) / kernel time / measured HBM peak"}
    for cls, a in sorted(agg.items(), key=lambda kv: -kv[1]["seconds"]):
        s = a["seconds"] or 1e-30
        res[cls] = {"launches": a["launches"], "ms": a["seconds"] * 1e3, "dram_gb": a["dram_bytes"] / 1e9,
                    "dram_gbs": a["dram_bytes"] / s / 1e9, "dram_frac": a["dram_bytes"] / s / 1e9 / peak,
                    "warp_inst_per_launch": a["warp_inst"] / max(a["launches"], 1), "regs": a["regs"],
                    **{k + "_pct": a[k] / s for k in pct}}
        if bench and cls in bench["roofline"]["per_kernel"]:
            pk = bench["roofline"]["per_kernel"][cls]
            algo = pk["achieved_gbs"] * 1e9 * pk["ms_per_step"] / 1e3          # algorithmic bytes of the class in one step
            res[cls]["algorithmic_gb_same_launches"] = algo / 1e9
            res[cls]["ratio"] = (a["dram_bytes"] / algo) if algo > 0 else None
    if bench:
        res["_bench_of_same_run"] = {"projects_per_gpu": bench["config"]["projects_per_gpu"], "launches_in_step": first_n}
    json.dump(res, open(out_json, "w"), indent=1)
    with open(out_md, "w") as f:
        f.write(f"# ncu --set full, per kernel class: {label}\n\nsources: {', '.join(os.path.basename(r) for r in raws)}; measured HBM peak {peak:.0f} GB/s\n\n")
        f.write("| kernel class | launches | DRAM GB/s | fraction of HBM peak (DRAM bytes) | issue slots % | ALU % | FMA % | FP64 % | XU % | LSU % | L2 hit % | regs | warp inst / launch |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for cls, v in res.items():
            if cls.startswith("_"):
                continue
            f.write(f"| `{cls}` | {v['launches']} | {v['dram_gbs']:.0f} | {v['dram_frac']:.2f} | {v['issue_pct']:.1f} | {v['alu_pct']:.1f} | {v['fma_pct']:.1f} | "
                    f"{v['fp64_pct']:.1f} | {v['xu_pct']:.1f} | {v['lsu_pct']:.1f} | {v['l2_hit_pct']:.1f} | {v['regs']} | {v['warp_inst_per_launch']:.3g} |\n")


if __name__ == "__main__":
    args = sys.argv[1:]
    blog = None
    if "--bench" in args:
        i = args.index("--bench"); blog = args[i + 1]; del args[i:i + 2]
    main(args[0], args[1], args[2], args[3:], blog)
