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
on the host clock, with the library's default software pipelining of project groups.

`--workload c1|c2|c3` run BASELINE.json configs[0..2] on one GPU (c1 and c2 move 12-200 MB: launch-bound by
construction); `--workload c4` is BASELINE.json configs[3], the single long 64-vertex graph: at N>1 it is sharded by
TIME over the ranks (strong scaling; segment hand-off through NCCL inside the library).
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
    ap.add_argument("--workload", default="c5", choices=["c5", "c1", "c2", "c3", "c4"])
    ap.add_argument("--projects", type=int, default=512, help="projects per GPU (c5)")
    ap.add_argument("--scale", type=float, default=1.0, help="duration scale of the workload")
    ap.add_argument("--cpu-sample", type=int, default=12, help="projects in the cpu_baseline sample (c5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the `secondary` block of the default line (configs[3] time-sharded over the ranks, configs[2] on rank 0)")
    ap.add_argument("--secondary-steps", type=int, default=3)
    ap.add_argument("--secondary-scale", type=float, default=1.0, help="duration scale of the secondary workloads")
    ap.add_argument("--value-group", type=int, default=0,
                    help="projects per pipelined group for the device-resident `value` passes (0 = one launch sequence)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------ workloads
def build_projects(args, rank: int):
    """[(project, assets, settings)] for this rank's shard."""
    from termdaw_b200 import synthetic as S
    if args.workload == "c5":
        base = rank * args.projects
        return [S.config_5_project(base + i, args.scale) for i in range(args.projects)]
    if args.workload == "c1":
        return [S.config_1(args.scale)]
    if args.workload == "c2":
        return [S.config_2(args.scale)]
    if args.workload == "c3":
        return [S.config_3(args.scale)]
    return [S.config_4(args.scale)]


WORKLOAD_NAMES = {
    "c5": "configs[4]: algorithmic-composition batch, 4096 projects x 30 s over 8 GPUs (512 projects per GPU shard)",
    "c1": "configs[0]: README example, 2 sampleloops -> normalize, 3 s, rendered to 44.1 kHz 16-bit (sinc down-sampler)",
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


def _rewind_args(st, p, a):
    """Per project: the float32 event arrays of its flowws in load order, as ready ctypes arguments (the
    conversion is done once; the calls below still hand the library host pointers every step)."""
    import ctypes as C
    import numpy as np
    from termdaw_b200._capi import c_f32p
    out = []
    for (_, path) in p.midis:
        t, n, v = (np.ascontiguousarray(x, dtype=np.float32) for x in a.flowws[path])
        out.append((t.ctypes.data_as(c_f32p), n.ctypes.data_as(c_f32p), v.ctypes.data_as(c_f32p), C.c_uint64(t.size),
                    12 * t.size, (t, n, v)))
    return out


def rewind(states, reupload_flowws: bool, _cache={}):
    """What the 'stop' + 'refresh' commands do between two renders of an unchanged project
    (ui_workflow.rs:153-160, state.rs:240-250,328-467): reload the flowws, rebuild every vertex (all
    per-vertex state back to its constructor value), rewind the transport, reset the normalize maxima.
    Without the rebuild the reference's vertex state would leak from one step into the next and the steps
    would not be the same work."""
    import ctypes as C
    from termdaw_b200._capi import c_f32p
    key = (id(states), len(states))
    packed = _cache.get(key)
    if packed is None:
        # the batched form of the refresh (tdg_rewind_many): the host pointers of every project's event arrays in
        # one flattened table, built once; every step still hands the library those host buffers to reload
        per, ts, ns, vs, counts, keep, nbytes = [], [], [], [], [], [], 0
        for st, p, a in states:
            args = _rewind_args(st, p, a)
            per.append(len(args))
            for tp, npx, vp, n, nb, k in args:
                ts.append(tp); ns.append(npx); vs.append(vp); counts.append(n.value); keep.append(k); nbytes += nb
        m = len(ts)
        packed = _cache[key] = {
            "hs": (C.c_void_p * len(states))(*[st.ctx.h for st, _, _ in states]),
            "t": (c_f32p * max(m, 1))(*ts), "n": (c_f32p * max(m, 1))(*ns), "v": (c_f32p * max(m, 1))(*vs),
            "counts": (C.c_uint64 * max(m, 1))(*counts), "per": (C.c_uint32 * len(states))(*per), "keep": keep, "nbytes": nbytes}
    lib = states[0][0].ctx.lib
    k = packed
    if reupload_flowws == "tables":          # the caller hands them to tdg_refresh_normalize_render_many itself
        return k
    if reupload_flowws:
        rc = lib.rewind_many(k["hs"], len(states), k["t"], k["n"], k["v"], k["counts"], k["per"])
    else:
        rc = lib.rewind_many(k["hs"], len(states), None, None, None, None, None)
    if rc < 0:
        msgs = [(lib.last_error(st.ctx.h) or b"").decode() for st, _, _ in states]
        raise RuntimeError(next((m_ for m_ in msgs if m_), "tdg_rewind_many failed"))
    return k["nbytes"] if reupload_flowws else 0


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
    frac = {"c1": 1.0, "c2": 1.0, "c3": 0.004, "c4": 0.004}[args.workload]
    cfg = {"c1": S.config_1, "c2": S.config_2, "c3": S.config_3, "c4": S.config_4}[args.workload](args.scale * frac)
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
    print_line(json.dumps({
        "impl": "reference", "metric": "rendered frames/sec", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[args.workload], "realtime_factor": v / PSR},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------ GPU arm
def _merge_kstats(dst, src):
    for name, v in src.items():
        d = dst.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0})
        for f in d:
            d[f] += v[f]


class BatchRunner:
    """Project sharding: this rank's shard of independent projects, scan + render per step."""
    scaling = "weak"

    def __init__(self, lib, args, rank, local_rank, world, dist):
        import torch
        self.lib, self.torch = lib, torch
        cfgs = build_projects(args, rank)
        self.states = make_states(lib, cfgs, local_rank)
        self.ctxs = [st.ctx for st, _, _ in self.states]
        self.cs = [st.cs for st, _, _ in self.states]
        self.rsr = [st.render_sr for st, _, _ in self.states]
        self.bds = [st.bd for st, _, _ in self.states]
        self.bl = self.states[0][0].bl
        self.frames_per_step_rank = sum(c * st.bl for c, (st, _, _) in zip(self.cs, self.states))
        self.total_frames_per_step = self.frames_per_step_rank * world      # every rank renders an equal shard
        self.ctxs[0].set_profiling(True)
        self.value_group = args.value_group
        self.outs = None
        self.parallelism = f"project-sharded x{world}" if world > 1 else "single GPU"

    def alloc_outs(self):
        torch = self.torch
        self.outs = [torch.empty(int(self.lib.render_bound_bytes(c.h, n, r, b)), dtype=torch.uint8, pin_memory=True).numpy()
                     for c, n, r, b in zip(self.ctxs, self.cs, self.rsr, self.bds)]

    def step(self, e2e: bool):
        from termdaw_b200 import normalize_render_many, normalize_scan_many, render_many
        c0 = self.ctxs[0]
        t_rw = time.perf_counter()
        fused_refresh = e2e and os.environ.get("TDG_E2E_FUSED_REFRESH", "1") != "0"
        if fused_refresh:
            k = rewind(self.states, reupload_flowws="tables")
            ev_bytes = k["nbytes"]
        else:
            ev_bytes = rewind(self.states, reupload_flowws=e2e)
        self.rewind_ms = (time.perf_counter() - t_rw) * 1e3
        if e2e:
            # the public batched call: refresh + scan + render of every project, groups of 32 projects (the library's
            # default) pipelined over two streams (refresh + planning / kernels / PCM D2H of different groups overlap)
            c0.set_pipeline_group(int(os.environ.get("TDG_E2E_GROUP", "32")))
            if fused_refresh:
                from termdaw_b200 import refresh_normalize_render_many
                refresh_normalize_render_many(self.lib, self.ctxs, (k["t"], k["n"], k["v"], k["counts"], k["per"]), self.cs, self.rsr,
                                              self.bds, want_pcm=True, outs=self.outs)
            else:
                normalize_render_many(self.lib, self.ctxs, self.cs, self.rsr, self.bds, want_pcm=True, outs=self.outs)
            s2, k2 = c0.last_pass_stats(), c0.last_pass_kernel_stats()
            s1, k1 = {k: 0 for k in s2}, {}
        else:
            # value: one launch sequence per pass for the whole shard (clean per-kernel event timing)
            c0.set_pipeline_group(self.value_group)
            normalize_scan_many(self.lib, self.ctxs, self.cs)
            s1, k1 = c0.last_pass_stats(), c0.last_pass_kernel_stats()
            render_many(self.lib, self.ctxs, self.cs, self.rsr, self.bds, want_pcm=False)
            s2, k2 = c0.last_pass_stats(), c0.last_pass_kernel_stats()
        ks = {}
        _merge_kstats(ks, k1); _merge_kstats(ks, k2)
        return {"device_ms": s1["device_ms"] + s2["device_ms"], "launches": s1["kernel_launches"] + s2["kernel_launches"],
                "h2d": ev_bytes + s1["h2d_bytes"] + s2["h2d_bytes"], "d2h": s2["d2h_bytes"], "kstats": ks,
                "algo": s1["algorithmic_bytes"] + s2["algorithmic_bytes"], "plan_ms": s1["plan_ms"] + s2["plan_ms"],
                "reused_v": s2.get("reused_vertices", 0), "reused_b": s2.get("reused_bytes", 0),
                "inlined_v": s1.get("inlined_vertices", 0) + s2.get("inlined_vertices", 0),
                "inlined_b": s1.get("inlined_bytes", 0) + s2.get("inlined_bytes", 0),
                "voice_frames": s1.get("voice_frames", 0) + s2.get("voice_frames", 0), "collectives": 0}

    def close(self):
        for st, _, _ in self.states:
            st.close()


class TimeShardRunner:
    """Time sharding: ONE long project, this rank renders its consecutive range of chunks."""
    scaling = "strong"

    def __init__(self, lib, args, rank, local_rank, world, dist):
        import torch
        from termdaw_b200.timeshard import TimeShardedRender
        self.lib, self.torch, self.rank, self.world = lib, torch, rank, world
        cfgs = build_projects(args, 0)                       # the same project on every rank
        self.states = make_states(lib, cfgs, local_rank)
        st = self.states[0][0]
        self.ts = TimeShardedRender(st, rank, world, dist)
        self.bl = st.bl
        self.cs = [st.cs]
        self.total_frames_per_step = st.cs * st.bl
        self.frames_per_step_rank = self.ts.bounds[rank][1] * st.bl
        st.ctx.set_profiling(True)
        self.out = None
        self.parallelism = f"time-sharded x{world} (segment hand-off over NCCL)" if world > 1 else "single GPU"

    def alloc_outs(self):
        n = self.ts.bounds[self.rank][1] * self.bl * 2 * (self.states[0][0].bd // 8)
        self.out = self.torch.empty(max(n, 1), dtype=self.torch.uint8, pin_memory=True).numpy()

    def step(self, e2e: bool):
        import ctypes as C
        st = self.states[0][0]
        c = st.ctx
        ev_bytes = rewind(self.states, reupload_flowws=e2e)
        col0 = c.collectives()
        self.ts.scan_exact()
        s1, k1 = c.last_pass_stats(), c.last_pass_kernel_stats()
        c0, n = self.ts.bounds[self.rank]
        if e2e:
            frames = C.c_uint64(0)
            rc = self.lib.render_segment(c.h, st.cs, c0, n, 0, 0, st.bd, self.out.ctypes.data_as(C.POINTER(C.c_uint8)),
                                         self.out.size, C.byref(frames))
            assert rc >= 0, self.lib.last_error(c.h)
        else:
            self.ts.render(want_pcm=False)
        s2, k2 = c.last_pass_stats(), c.last_pass_kernel_stats()
        ks = {}
        _merge_kstats(ks, k1); _merge_kstats(ks, k2)
        return {"device_ms": s1["device_ms"] + s2["device_ms"], "launches": s1["kernel_launches"] + s2["kernel_launches"],
                "h2d": ev_bytes + s1["h2d_bytes"] + s2["h2d_bytes"], "d2h": s2["d2h_bytes"], "kstats": ks,
                "algo": s1["algorithmic_bytes"] + s2["algorithmic_bytes"], "plan_ms": s1["plan_ms"] + s2["plan_ms"],
                "reused_v": s2.get("reused_vertices", 0), "reused_b": s2.get("reused_bytes", 0),
                "inlined_v": s1.get("inlined_vertices", 0) + s2.get("inlined_vertices", 0),
                "inlined_b": s1.get("inlined_bytes", 0) + s2.get("inlined_bytes", 0),
                "voice_frames": s1.get("voice_frames", 0) + s2.get("voice_frames", 0), "collectives": c.collectives() - col0}

    def close(self):
        self.ts.close()
        for st, _, _ in self.states:
            st.close()


def run_secondary(lib, args, rank, local_rank, world, dist, torch, barrier, peak):
    """The other two configurations BASELINE.json's north_star asks numbers for, measured inside the default run
    (fewer steps): configs[3] — ONE 64-vertex, 3600 s graph sharded by TIME over the N ranks (strong scaling;
    normalize peaks / band-pass state / resampler halo cross the segment boundaries through NCCL inside the
    library) — and configs[2] — 512 wavetable voices for 600 s — on rank 0.  Same timing rules as the primary."""
    import copy
    out = {}
    for wl in ("c4", "c3"):
        a = copy.copy(args)
        a.workload, a.scale, a.projects = wl, args.secondary_scale, 1
        solo = wl == "c3"                                  # one GPU's worth of work: rank 0 runs it, the others wait
        R = None
        res = None
        if not solo or rank == 0:
            R = (TimeShardRunner(lib, a, rank, local_rank, world, dist) if wl == "c4"
                 else BatchRunner(lib, a, 0, local_rank, 1, None))
        sync = (lambda: torch.cuda.synchronize()) if solo else barrier
        steps = max(1, args.secondary_steps)
        if R is not None:
            R.step(False)
            sync()
            dev_ms = plan_ms = 0.0
            launches = algo = collectives = vframes = 0
            kstats = {}
            for _ in range(steps):
                r = R.step(False)
                dev_ms += r["device_ms"]; launches += r["launches"]; algo += r["algo"]; plan_ms += r["plan_ms"]
                collectives += r["collectives"]; vframes += r.get("voice_frames", 0)
                _merge_kstats(kstats, r["kstats"])
            sync()
            (R.states[0][0].ctx).set_profiling(False)
            R.alloc_outs()
            R.step(True)
            sync()
            t1 = time.perf_counter()
            h2d = d2h = 0
            for _ in range(steps):
                r = R.step(True)
                h2d += r["h2d"]; d2h += r["d2h"]
            sync()
            e2e_wall = time.perf_counter() - t1
            t = torch.tensor([dev_ms, e2e_wall], dtype=torch.float64, device="cuda")
            if dist is not None and not solo:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev_ms_max, e2e_max = t.tolist()
            frames = R.total_frames_per_step
            x = kstats.get("nccl_exchange", {"launches": 0, "ms": 0.0})
            res = {
                "workload": WORKLOAD_NAMES[wl], "n_gpus": 1 if solo else world, "scaling": "strong" if wl == "c4" else "single GPU",
                "parallelism": R.parallelism, "steps": steps, "seconds": R.cs[0] * R.bl / PSR,
                "value": frames * steps / (dev_ms_max / 1e3), "unit": "frames/s", "ms_per_step": dev_ms_max / steps,
                "realtime_factor": frames * steps / (dev_ms_max / 1e3) / PSR,
                "e2e": {"value": frames * steps / e2e_max, "unit": "frames/s", "ms_per_step": e2e_max / steps * 1e3,
                        "h2d_bytes_per_step": h2d // steps, "d2h_bytes_per_step": d2h // steps},
                "host_plan_ms_per_step": plan_ms / steps, "gpu_launches": launches,
                "algorithmic_gb_per_step_this_rank": algo / steps / 1e9,
                "whole_step_algorithmic_gbs_this_rank": algo / (dev_ms / 1e3) / 1e9 if dev_ms > 0 else 0.0,
                "per_kernel_rank0": {k: {"ms_per_step": v["ms"] / steps, "launches_per_step": v["launches"] // steps,
                                         "frac_of_hbm_peak": (v["bytes"] / (v["ms"] / 1e3) / 1e9 / peak) if v["ms"] > 0 else 0.0}
                                     for k, v in sorted(kstats.items(), key=lambda kv: -kv[1]["ms"])[:8]},
            }
            if wl == "c4":
                res["nccl_exchange_rounds_per_step"] = collectives // steps
                res["nccl_exchange_ms_per_round_rank0"] = (x["ms"] / x["launches"]) if x["launches"] else None
                res["nccl_exchange_ms_per_step_rank0"] = x["ms"] / steps
            else:
                res["voice_frames_per_step"] = vframes // steps
                res["voice_frames_per_s"] = vframes / (dev_ms_max / 1e3)
            R.close()
        if not solo or dist is None:
            pass
        barrier()
        if rank == 0:
            out[wl] = res
    return out


def _claim_stdout():
    """stdout carries exactly ONE JSON line.  Native libraries print there too (NCCL announces its version on
    the first communicator), so fd 1 is pointed at stderr for the run and the line goes to the saved fd."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    args = parse()
    out = _claim_stdout()
    global print_line
    print_line = lambda s: (out.write(s + "\n"), out.flush())      # noqa: E731
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import termdaw_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the renderer has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = termdaw_b200.load_library()
    R = (TimeShardRunner if args.workload == "c4" else BatchRunner)(lib, args, rank, local_rank, world, dist)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    for _ in range(args.warmup):
        R.step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    dev_ms, launches, kstats, algo, plan_ms, collectives = 0.0, 0, {}, 0, 0.0, 0
    reused_v = reused_b = inlined_v = inlined_b = 0
    for _ in range(args.steps):
        r = R.step(False)
        dev_ms += r["device_ms"]; launches += r["launches"]; algo += r["algo"]; plan_ms += r["plan_ms"]
        reused_v += r.get("reused_v", 0); reused_b += r.get("reused_b", 0)
        inlined_v += r.get("inlined_v", 0); inlined_b += r.get("inlined_b", 0)
        collectives += r["collectives"]
        _merge_kstats(kstats, r["kstats"])
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()

    # ---- the same steps with pass reuse off (every vertex evaluated in both passes, as the reference does):
    # reported beside `value` so that the saving is visible, never used for `value` itself
    dev_ms_off, off_steps = 0.0, min(args.steps, 3)
    ctx0 = R.states[0][0].ctx
    try:
        ctx0.set_pass_reuse(False)
        R.step(False)
        for _ in range(off_steps):
            dev_ms_off += R.step(False)["device_ms"]
    finally:
        ctx0.set_pass_reuse(True)

    # ---- end to end through the C ABI with host buffers (no per-launch events: that is a measurement aid
    # of the device-resident loop above, not part of the call a user makes)
    R.states[0][0].ctx.set_profiling(False)
    R.alloc_outs()
    for _ in range(2):
        R.step(True)
    barrier()
    t1 = time.perf_counter()
    h2d = d2h = 0
    e2e_steps = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        r = R.step(True)
        e2e_steps.append((time.perf_counter() - ts) * 1e3)
        if hasattr(R, "rewind_ms"):
            e2e_steps[-1] = (e2e_steps[-1], R.rewind_ms)
        h2d += r["h2d"]; d2h += r["d2h"]
    barrier()
    e2e_wall = time.perf_counter() - t1

    # ---- max over ranks
    t = torch.tensor([dev_ms, wall, e2e_wall, dev_ms_off], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_max, e2e_max, dev_ms_off_max = t.tolist()
    total_frames = R.total_frames_per_step
    value = total_frames * args.steps / (dev_ms_max / 1e3)
    e2e_value = total_frames * args.steps / e2e_max

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        # MEASURED_PEAKS.json is driver-written and git-ignored; the figure it held at round start (6532.9 GB/s
        # copy bandwidth on this pool) is recorded in BASELINE.md / SURVEY.md and used when the file is absent
        peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else \
            (6532.9, "measured (MEASURED_PEAKS.json value recorded in BASELINE.md; file absent on this box)")
        dom = max(kstats, key=lambda k: kstats[k]["ms"])
        ach = kstats[dom]["bytes"] / (kstats[dom]["ms"] / 1e3) / 1e9 if kstats[dom]["ms"] > 0 else 0.0
        # measured DRAM bytes per launch of the dominant kernel: the committed `ncu --set full` capture
        # (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum over one step at 64 projects,
        # beside the algorithmic bytes of the same launches) gives the traffic / algorithmic ratio of the
        # kernel; scaled to this run's algorithmic bytes per launch, like `achieved`
        # profiles/ncu_classes.json (tools/ncu_class_summary.py over an `ncu --set full` capture of bench.py at 64
        # projects; its `_label` names the commit it was taken on): per kernel class the measured DRAM bytes / the
        # algorithmic bytes of the same launches (`ratio`), the DRAM-traffic fraction of the HBM peak and the
        # issue-slot / pipe utilisations.  Scaled to this run's algorithmic bytes per launch, like `achieved`.
        traffic = traffic_detail = None
        ncu = {}
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_classes.json")))
            traffic_detail = ncu.get(dom)
            if traffic_detail and traffic_detail.get("ratio") is not None and kstats[dom]["launches"]:
                traffic = traffic_detail["ratio"] * kstats[dom]["bytes"] / kstats[dom]["launches"]
        except (OSError, ValueError, KeyError):
            pass
        issue_pct = (ncu.get(dom) or {}).get("issue_pct")

        def per_kernel_entry(k, v):
            e = {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] // args.steps,
                 "achieved_gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["ms"] > 0 else 0.0,
                 "frac": (v["bytes"] / (v["ms"] / 1e3) / 1e9 / peak) if v["ms"] > 0 else 0.0}
            n = ncu.get(k)
            if n and v["ms"] > 0:
                # the fraction of the HBM peak from DRAM traffic (a gather whose source sits in L2 must not read as
                # > 1.0 of HBM): algorithmic fraction x measured DRAM / algorithmic ratio of the class
                if n.get("ratio") is not None:
                    e["dram_frac"] = e["frac"] * n["ratio"]
                e["issue_slots_pct_ncu"] = n.get("issue_pct")
            return e
        line = {
            "metric": "rendered frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": R.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": WORKLOAD_NAMES[args.workload], "projects_per_gpu": len(R.states), "parallelism": R.parallelism,
                "seconds_per_project": R.cs[0] * R.bl / PSR, "flow": "refresh -> normalize scan -> render",
                "project_sr": PSR, "buffer_length": R.bl,
                "realtime_factor": value / PSR, "l2": "per-step working set >> 126 MB L2 (no flush needed)"
                if R.frames_per_step_rank * 8 * 4 > 4e8 else "working set fits L2: numbers are L2-assisted",
                "value_timing": "CUDA events around each pass's kernel sequence (plan tables already uploaded)",
                "host_wall_ms_per_step": wall_max / args.steps * 1e3, "host_plan_ms_per_step": plan_ms / args.steps,
                "algorithmic_gb_per_step": algo / args.steps / 1e9,
                "whole_step_algorithmic_gbs": algo / (dev_ms / 1e3) / 1e9,
                "nccl_exchange_rounds_per_step": collectives // args.steps,
                # the render reads vertices whose whole cone Graph::set_time resets from the scan before it instead
                # of evaluating them twice (tdg_set_pass_reuse): algorithmic_gb_per_step is what the
                # launched kernels move; SURVEY 8(d)'s both-passes-evaluate-everything figure is that plus the skipped
                "pass_reuse": {"vertices_read_from_scan_per_step": reused_v // args.steps,
                               "algorithmic_gb_not_moved_per_step": reused_b / args.steps / 1e9,
                               "survey_8d_algorithmic_gb_per_step": (algo + reused_b + inlined_b) / args.steps / 1e9,
                               "ms_per_step_with_reuse_off": dev_ms_off_max / off_steps,
                               "value_with_reuse_off": total_frames * off_steps / (dev_ms_off_max / 1e3)},
                # Sum vertices evaluated inside their consumer's kernel (tdg_set_inline_sums): no buffer, no
                # launch; the write + read of the intermediate buffer is not in algorithmic_gb_per_step either
                "inlined_sums": {"vertices_per_step": inlined_v // args.steps, "gb_not_moved_per_step": inlined_b / args.steps / 1e9},
            },
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps, "realtime_factor": e2e_value / PSR,
                    "ms_per_step": e2e_max / args.steps * 1e3, "ms_each_step_rank0": [round(x[0], 2) if isinstance(x, tuple) else round(x, 2) for x in e2e_steps],
                    "host_refresh_ms_each_step": [round(x[1], 2) for x in e2e_steps if isinstance(x, tuple)],
                    "refresh": ("inside the pipelined call (tdg_refresh_normalize_render_many): a group's refresh runs while the GPU renders the groups in front"
                                if os.environ.get("TDG_E2E_FUSED_REFRESH", "1") != "0" else "tdg_rewind_many before the call")},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "frac_of_nominal_8000_gbs": ach / 8000.0, "traffic": traffic,
                         "traffic_source": traffic_detail,
                         "issue_slot_utilisation_pct_ncu": issue_pct, "peak_source": peak_src,
                         "share_of_step": kstats[dom]["ms"] / dev_ms,
                         "note": "k_voice<*> kernels are instruction-issue bound, not HBM bound (DESIGN.md §5): their HBM "
                                 "fraction is reported as the contract asks; every streaming kernel is listed in per_kernel "
                                 "(frac = algorithmic bytes / time / peak; dram_frac = the same from measured DRAM bytes)",
                         "ncu_capture": ncu.get("_label"),
                         "per_kernel": {k: per_kernel_entry(k, v) for k, v in sorted(kstats.items(), key=lambda kv: -kv[1]["ms"])}},
        }
        if world == 1 and not args.no_cpu_baseline:
            ccfgs, sample = cpu_sample_cfgs(args)
            fr, dt = oracle_render_projects(ccfgs, 1)
            cpu_model = None
            try:
                with open("/proc/cpuinfo") as f:
                    cpu_model = next((ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")), None)
            except OSError:
                pass
            line["cpu_baseline"] = {"value": fr / dt, "unit": "frames/s", "cores": 1, "kind": "port", "cpu_model": cpu_model,
                                    "sample": sample + f" ({dt:.1f} s of CPU work; restatement of the reference, not the Rust build)",
                                    "host_cores_available": os.cpu_count()}
    R.close()
    if args.workload == "c5" and not args.no_secondary:
        peak_gbs = 6532.9
        try:
            peak_gbs = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak_gbs)
        except (OSError, ValueError):
            pass
        secondary = run_secondary(lib, args, rank, local_rank, world, dist, torch, barrier, peak_gbs)
        if rank == 0:
            line["secondary"] = secondary
    if rank == 0:
        print_line(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
