"""Print the per-kernel table of one or more bench JSON lines side by side (ms per step)."""
import json, sys
rows = [json.load(open(p)) for p in sys.argv[1:]]
names = []
for d in rows:
    for k in d["roofline"]["per_kernel"]:
        if k not in names:
            names.append(k)
print(f"{'kernel':20s}" + "".join(f"{p[-22:]:>24s}" for p in sys.argv[1:]))
for k in names:
    print(f"{k:20s}" + "".join(f"{d['roofline']['per_kernel'].get(k, {}).get('ms_per_step', 0):12.2f} {d['roofline']['per_kernel'].get(k, {}).get('frac', 0):10.3f} " for d in rows))
print(f"{'sum':20s}" + "".join(f"{sum(v['ms_per_step'] for v in d['roofline']['per_kernel'].values()):12.2f} {'':10s} " for d in rows))
print(f"{'ms_per_step':20s}" + "".join(f"{d['ms_per_step']:12.2f} {'':10s} " for d in rows))
print(f"{'e2e ms':20s}" + "".join(f"{d['e2e']['ms_per_step']:12.2f} {'':10s} " for d in rows))
