"""Pinned host<->device copy bandwidth of the box (what bounds bench.py's e2e: the PCM bytes of a step
have to cross PCIe once).  Prints one JSON line."""
import json, torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for name, (dst, src) in {"d2h": (h, d), "h2d": (d, h)}.items():
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dst.copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize()
        best = max(best, n / (e0.elapsed_time(e1) / 1e3) / 1e9)
    out[name + "_gbs"] = round(best, 2)
print(json.dumps(out))
