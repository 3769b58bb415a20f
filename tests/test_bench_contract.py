"""bench.py's control flow and its ONE JSON line, exercised on CPU with a fake runner in place of the GPU
runner (no kernels run, so none of the numbers mean anything): every key the bench contract names must
be there with the right type, the extra `roofline` / `cpu_baseline` / `e2e` objects included, and the
reference arm (which really is CPU work: the oracle) must print its own line."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT

sys.path.insert(0, ROOT)


class _FakeCtx:
    def set_profiling(self, on): pass
    def set_pass_reuse(self, on): self.reuse = on


class _FakeState:
    def __init__(self):
        self.ctx, self.cs, self.bl = _FakeCtx(), 1407, 1024


class _FakeRunner:
    scaling = "weak"

    def __init__(self, lib, args, rank, local_rank, world, dist):
        self.states = [(_FakeState(), None, None) for _ in range(4)]
        self.cs = [st.cs for st, _, _ in self.states]
        self.bl = 1024
        self.frames_per_step_rank = sum(self.cs) * self.bl
        self.total_frames_per_step = self.frames_per_step_rank * world
        self.parallelism = "single GPU"
        self.rewind_ms = 0.5

    def alloc_outs(self): pass

    def step(self, e2e):
        ks = {"k_voice<sampsyn>": {"launches": 2, "ms": 3.0, "bytes": 3_000_000_000},
              "k_fx<sum>": {"launches": 8, "ms": 1.0, "bytes": 6_000_000_000},
              "k_bp_carry": {"launches": 8, "ms": 0.1, "bytes": 0}}
        return {"device_ms": 5.0, "launches": 18, "h2d": 1000, "d2h": 2000, "kstats": ks, "algo": 9_000_000_000,
                "plan_ms": 1.0, "collectives": 0, "reused_v": 7, "reused_b": 1_000_000}

    def close(self): pass


def test_bench_line_has_every_contract_key(monkeypatch, capfd):
    import torch
    import bench
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda data, **kw: real_tensor(data, **{k: v for k, v in kw.items() if k != "device"}))
    monkeypatch.setattr(bench, "BatchRunner", _FakeRunner)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "2", "--warmup", "3", "--cpu-sample", "1", "--scale", "0.02", "--no-secondary"])
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    bench.main()
    out = capfd.readouterr().out.strip().splitlines()
    assert len(out) == 1, out                                 # stdout is the one JSON line
    d = json.loads(out[0])
    for key, typ in (("metric", str), ("value", float), ("unit", str), ("n_gpus", int), ("steps", int), ("warmup", int),
                     ("ms_per_step", float), ("higher_is_better", bool), ("scaling", str), ("dtype", str), ("data", str),
                     ("config", dict), ("e2e", dict), ("gpu_launches", int), ("clocks", dict), ("roofline", dict),
                     ("cpu_baseline", dict)):
        assert isinstance(d[key], typ), key
    assert d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["metric"] == "rendered frames/sec" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] == pytest.approx(4 * 1407 * 1024 * 2 / (10.0 / 1e3))
    assert d["gpu_launches"] == 36
    e = d["e2e"]
    assert e["unit"] == "frames/s" and e["value"] > 0 and e["h2d_bytes_per_step"] == 1000 and e["d2h_bytes_per_step"] == 2000
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and r["kernel"] == "k_voice<sampsyn>"
    assert r["achieved"] == pytest.approx(1000.0) and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert r["traffic"] is None or isinstance(r["traffic"], float)
    assert set(r["per_kernel"]) == {"k_voice<sampsyn>", "k_fx<sum>", "k_bp_carry"}
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] == 1 and c["value"] > 0 and c["unit"] == "frames/s" and c["sample"]
    pr = d["config"]["pass_reuse"]
    assert pr["vertices_read_from_scan_per_step"] == 7 and pr["ms_per_step_with_reuse_off"] == pytest.approx(5.0)


def test_reference_arm_prints_its_line():
    env = {**os.environ, "RANK": "0", "WORLD_SIZE": "1"}
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--scale", "0.05"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rendered frames/sec" and d["unit"] == "frames/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # under torchrun only rank 0 works
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=60, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"}, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_real_runner_on_the_emulated_library(monkeypatch, capfd, tmp_path):
    """The REAL BatchRunner (rewind, scan, render, pipelined e2e call, kernel statistics, pass-reuse A/B) through
    bench.main(), with the emulated library (tests/emu) standing in for the CUDA one: two tiny projects, one step.
    The numbers are meaningless (CPU emulation); what is checked is that every stage runs and the line is complete."""
    import torch
    import bench
    import termdaw_b200
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from termdaw_b200._capi import Library
    emu = Library(build_emu.build(str(tmp_path / "emu")), "tdg_")
    monkeypatch.setattr(termdaw_b200, "load_library", lambda: emu)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    real_tensor, real_empty = torch.tensor, torch.empty
    monkeypatch.setattr(torch, "tensor", lambda data, **kw: real_tensor(data, **{k: v for k, v in kw.items() if k != "device"}))
    monkeypatch.setattr(torch, "empty", lambda *a, **kw: real_empty(*a, **{k: v for k, v in kw.items() if k != "pin_memory"}))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "1", "--warmup", "1", "--projects", "2", "--scale", "0.02",
                                      "--cpu-sample", "1", "--secondary-scale", "0.001", "--secondary-steps", "1"])
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    bench.main()
    out = capfd.readouterr().out.strip().splitlines()
    assert len(out) == 1, out
    d = json.loads(out[0])
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 0
    # the secondary block: configs[3] (time-sharded runner, one rank here) and configs[2] (512 voices) really ran
    sec = d["secondary"]
    assert set(sec) == {"c3", "c4"}
    for wl in ("c3", "c4"):
        assert sec[wl]["value"] > 0 and sec[wl]["e2e"]["value"] > 0 and sec[wl]["gpu_launches"] > 0, wl
        assert sec[wl]["e2e"]["d2h_bytes_per_step"] > 0 and sec[wl]["host_plan_ms_per_step"] > 0
    assert sec["c4"]["scaling"] == "strong" and sec["c4"]["nccl_exchange_rounds_per_step"] == 0       # one rank: nothing to exchange
    assert sec["c3"]["voice_frames_per_s"] > 0 and sec["c3"]["voice_frames_per_step"] > 2 * 500 * 0.5 * 48000 * 0.6
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] == 2 * 29 * 1024 * 4
    assert d["config"]["projects_per_gpu"] == 2
    assert d["config"]["pass_reuse"]["vertices_read_from_scan_per_step"] > 0
    assert d["config"]["pass_reuse"]["ms_per_step_with_reuse_off"] > 0
    assert d["roofline"]["kernel"] in d["roofline"]["per_kernel"] and d["roofline"]["achieved"] > 0
    assert d["cpu_baseline"]["value"] > 0
