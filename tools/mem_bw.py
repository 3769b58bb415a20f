"""Read-only, write-only and copy bandwidth of the GPU (torch kernels, CUDA events, best of 5 over 4 GiB): the
ceilings write-dominated kernels (sample gathers: 8 B written per frame, the source sample read from L2) and
read-dominated ones are judged against next to the copy figure of MEASURED_PEAKS.json.  One JSON line."""
import json, torch
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
b = torch.empty_like(a)
def t(fn, nbytes):
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = max(best, nbytes / (e0.elapsed_time(e1) / 1e3) / 1e9)
    return round(best, 1)
out = {"write_only_gbs": t(lambda: b.fill_(1.5), 4 * n), "read_only_gbs": t(lambda: a.sum(), 4 * n),
       "copy_gbs_read_plus_write": t(lambda: b.copy_(a), 8 * n)}
print(json.dumps(out))
