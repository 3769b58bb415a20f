#!/usr/bin/env python
"""bench.py — rendered frames/s of termdaw's audio-graph render path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c3|c4] [--impl reference]

A step = one pass of the hot path over one batch of synthetic input: for every project of the
rank's shard the reference's canonical sequence refresh -> normalize (scan) -> render
(graph.rs:222-237, state.rs:477-577), i.e. two whole-timeline evaluations of the graph plus
quantisation.  Default workload: BASELINE.json configs[4] — the algorithmic-composition batch,
4096 projects x 30 s over 8 GPUs = 512 projects per GPU (weak scaling; the same 512-project shard at
N=1).  `value` = project-rate frames of rendered audio per second summed over ranks, timed on the
device (CUDA events around each pass's kernel sequence, max over ranks); `e2e` = the same through
the C ABI with host buffers (floww events re-uploaded, plan tables H2D, PCM D2H into pinned memory)
on the host clock.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PSR = 48000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5", "c2", "c3", "c4"])
    ap.add_argument("--projects", type=int, default=512, help="projects per GPU (c5)")
    ap.add_argument("--scale", type=float, default=1.0, help="duration scale of the workload")
    ap.add_argument("--cpu-sample", type=int, default=12, help="projects in the cpu_baseline sample (c5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------ workloads
def build_projects(args, rank: int):
    """[(project, assets, settings)] for this rank's shard."""
    from termdaw_b200 import synthetic as S
    if args.workload == "c5":
        base = rank * args.projects
        return [S.config_5_project(base + i, args.scale) for i in range(args.projects)]
    if args.workload == "c2":
        return [S.config_2(args.scale)]
    if args.workload == "c3":
        return [S.config_3(args.scale)]
    return [S.config_4(args.scale)]


WORKLOAD_NAMES = {
    "c5": "configs[4]: algorithmic-composition batch, 4096 projects x 30 s over 8 GPUs (512 projects per GPU shard)",
    "c2": "configs[1]: floww synth + ADSR -> band pass -> normalize, 60 s, 48 kHz, 24-bit",
    "c3": "configs[2]: wavetable synth 512 voices + 2 sample_lerp (44.1/96 kHz sources), 600 s",
    "c4": "configs[3]: 64-vertex mixed graph, 3600 s, 24-bit",
}


def make_states(lib, cfgs, device):
    from termdaw_b200.state import State
    states = []
    for p, a, s in cfgs:
        st = State(lib, s["psr"], s["bl"], device=device, assets=a)
        assert st.refresh(p), st.messages
        states.append((st, p, a))
    return states


def rewind(states, reupload_flowws: bool):
    """What the 'stop' + 'refresh' commands do between two renders of a project (ui_workflow.rs:153-160,
    state.rs:240-250,467): rewind transport and floww cursor, reload flowws, reset normalize maxima."""
    nbytes = 0
    for st, p, a in states:
        c = st.ctx
        if reupload_flowws:
            c.floww_reset()
            for (_, path) in p.midis:
                t, n, v = a.flowws[path]
                c.floww_add(t, n, v)
                nbytes += 12 * len(t)
        c.set_time(0)
        c.floww_set_time(0)
        c.reset_normalize()
    return nbytes


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU arms
def oracle_render_projects(cfgs, threads: int):
    """scan + render of each project through the CPU oracle (restatement of the reference, not the Rust
    build); returns (frames, seconds).  Projects are independent, so `threads` of them run at once."""
    from helpers import load_oracle
    from termdaw_b200.state import State
    lib = load_oracle()
    states = []
    for p, a, s in cfgs:
        st = State(lib, s["psr"], s["bl"], assets=a)
        assert st.refresh(p), st.messages
        states.append(st)

    def work(st):
        st.scan_exact()
        st.render()
        return st.cs * st.bl

    t0 = time.perf_counter()
    if threads <= 1:
        frames = sum(work(st) for st in states)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=threads) as ex:          # ctypes releases the GIL
            frames = sum(ex.map(work, states))
    dt = time.perf_counter() - t0
    for st in states:
        st.close()
    return frames, dt


def cpu_sample_cfgs(args):
    from termdaw_b200 import synthetic as S
    if args.workload == "c5":
        return [S.config_5_project(i, args.scale) for i in range(args.cpu_sample)], \
            f"{args.cpu_sample} of the shard's projects, full 30 s each, scan + render"
    frac = {"c2": 1.0, "c3": 0.004, "c4": 0.004}[args.workload]
    cfg = {"c2": S.config_2, "c3": S.config_3, "c4": S.config_4}[args.workload](args.scale * frac)
    return [cfg], f"first {frac:.3%} of the timeline, scan + render"


def run_reference(args, rank: int):
    """Reference arm: the reference's own CPU implementation of the path.  The Rust build is impossible
    here (no cargo, crates not vendored), so this is the oracle port on all host threads."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    from termdaw_b200 import synthetic as S
    if args.workload == "c5":
        per_step = max(cores, 8)
        cfgs = [S.config_5_project(i, args.scale) for i in range(per_step)]
        sample = f"{per_step} projects x 30 s per step on {cores} threads (projects are independent)"
        threads = cores
    else:
        cfgs, sample = cpu_sample_cfgs(args)
        threads = 1
    times, frames = [], 0
    for i in range(args.warmup + args.steps):
        fr, dt = oracle_render_projects(cfgs, threads)
        if i >= args.warmup:
            times.append(dt); frames = fr
    dt = sum(times) / len(times)
    v = frames / dt
    print(json.dumps({
        "impl": "reference", "metric": "rendered frames/sec", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[args.workload], "realtime_factor": v / PSR},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import termdaw_b200
    from termdaw_b200 import normalize_scan_many, render_many

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the renderer has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = termdaw_b200.load_library()

    cfgs = build_projects(args, rank)
    states = make_states(lib, cfgs, local_rank)
    ctxs = [st.ctx for st, _, _ in states]
    cs = [st.cs for st, _, _ in states]
    rsr = [st.render_sr for st, _, _ in states]
    bds = [st.bd for st, _, _ in states]
    frames_per_step = sum(c * st.bl for c, (st, _, _) in zip(cs, states))
    ctxs[0].set_profiling(True)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step(e2e: bool, outs=None):
        """returns (device_ms, launches, h2d_bytes, d2h_bytes, kernel stats)"""
        ev_bytes = rewind(states, reupload_flowws=e2e)
        normalize_scan_many(lib, ctxs, cs)
        s1, k1 = ctxs[0].last_pass_stats(), ctxs[0].last_pass_kernel_stats()
        if e2e:
            render_many(lib, ctxs, cs, rsr, bds, want_pcm=True, outs=outs)
        else:
            render_many(lib, ctxs, cs, rsr, bds, want_pcm=False)
        s2, k2 = ctxs[0].last_pass_stats(), ctxs[0].last_pass_kernel_stats()
        ks = {}
        for k in (k1, k2):
            for name, v in k.items():
                d = ks.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0})
                for f in d:
                    d[f] += v[f]
        return (s1["device_ms"] + s2["device_ms"], s1["kernel_launches"] + s2["kernel_launches"],
                ev_bytes + s1["h2d_bytes"] + s2["h2d_bytes"], s2["d2h_bytes"], ks,
                s1["algorithmic_bytes"] + s2["algorithmic_bytes"], s1["plan_ms"] + s2["plan_ms"])

    # ---- device-resident timing (value)
    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    dev_ms, launches, kstats, algo, plan_ms = 0.0, 0, {}, 0, 0.0
    for _ in range(args.steps):
        d, l, _, _, ks, ab, pm = step(False)
        dev_ms += d; launches += l; algo += ab; plan_ms += pm
        for name, v in ks.items():
            dd = kstats.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0})
            for f in dd:
                dd[f] += v[f]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()

    # ---- end to end through the C ABI with host buffers
    outs = [torch.empty(int(lib.render_bound_bytes(c.h, n, r, b)), dtype=torch.uint8, pin_memory=True).numpy()
            for c, n, r, b in zip(ctxs, cs, rsr, bds)]
    step(True, outs)
    barrier()
    t1 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        _, _, hb, db, _, _, _ = step(True, outs)
        h2d += hb; d2h += db
    barrier()
    e2e_wall = time.perf_counter() - t1

    # ---- max over ranks
    t = torch.tensor([dev_ms, wall, e2e_wall], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_max, e2e_max = t.tolist()
    total_frames = frames_per_step * world          # weak scaling: every rank renders an equal shard
    value = total_frames * args.steps / (dev_ms_max / 1e3)
    e2e_value = total_frames * args.steps / e2e_max

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else \
            (6650.0, "fallback (B200_PROFILING.md)")
        dom = max(kstats, key=lambda k: kstats[k]["ms"])
        ach = kstats[dom]["bytes"] / (kstats[dom]["ms"] / 1e3) / 1e9 if kstats[dom]["ms"] > 0 else 0.0
        line = {
            "metric": "rendered frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": WORKLOAD_NAMES[args.workload], "projects_per_gpu": len(states),
                "seconds_per_project": cs[0] * states[0][0].bl / PSR, "flow": "refresh -> normalize scan -> render",
                "project_sr": PSR, "buffer_length": states[0][0].bl,
                "realtime_factor": value / PSR, "l2": "per-step working set >> 126 MB L2 (no flush needed)"
                if frames_per_step * 8 * 4 > 4e8 else "working set fits L2: numbers are L2-assisted",
                "value_timing": "CUDA events around each pass's kernel sequence (plan tables already uploaded)",
                "host_wall_ms_per_step": wall_max / args.steps * 1e3, "host_plan_ms_per_step": plan_ms / args.steps,
                "algorithmic_gb_per_step": algo / args.steps / 1e9,
                "whole_step_algorithmic_gbs": algo / (dev_ms / 1e3) / 1e9,
            },
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps, "realtime_factor": e2e_value / PSR},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                         "share_of_step": kstats[dom]["ms"] / dev_ms,
                         "per_kernel": {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] // args.steps,
                                            "achieved_gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["ms"] > 0 else 0.0}
                                        for k, v in sorted(kstats.items(), key=lambda kv: -kv[1]["ms"])}},
        }
        if world == 1 and not args.no_cpu_baseline:
            ccfgs, sample = cpu_sample_cfgs(args)
            fr, dt = oracle_render_projects(ccfgs, 1)
            line["cpu_baseline"] = {"value": fr / dt, "unit": "frames/s", "cores": 1, "kind": "port",
                                    "sample": sample + f" ({dt:.1f} s of CPU work; restatement of the reference, not the Rust build)",
                                    "host_cores_available": os.cpu_count()}
        print(json.dumps(line))
    for st, _, _ in states:
        st.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
