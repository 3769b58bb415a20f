"""Aggregate pinned device->host bandwidth of the box with every rank copying AT THE SAME TIME (torchrun): the
ceiling of bench.py's e2e at N GPUs, where every step each rank hands its PCM (2.95 GB at 512 projects) to the host.
Prints one JSON line on rank 0: per-rank and aggregate GB/s, alone (rank 0 only) and concurrently."""
import json, os, time, torch
import torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
def bar():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
def run(reps, active):
    bar()
    t0 = time.perf_counter()
    if active:
        for _ in range(reps): h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    bar()
    return n * reps / dt / 1e9 if active else 0.0
run(2, True)
alone = run(6, rank == 0)
together = run(6, True)
t = torch.tensor([alone, together], dtype=torch.float64, device="cuda")
if world > 1:
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
else:
    g = [t]
if rank == 0:
    tog = [float(x[1]) for x in g]
    print(json.dumps({"n_gpus": world, "d2h_alone_gbs_rank0": round(float(g[0][0]), 2), "d2h_concurrent_gbs_per_rank": [round(x, 2) for x in tog],
                      "d2h_concurrent_gbs_aggregate": round(sum(tog), 2), "numa_nodes": len([x for x in os.listdir("/sys/devices/system/node") if x.startswith("node")]) if os.path.isdir("/sys/devices/system/node") else None,
                      "cpus": os.cpu_count()}))
if world > 1: dist.destroy_process_group()
