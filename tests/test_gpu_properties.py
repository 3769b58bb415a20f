"""Full-size parity and size-independent properties of the CUDA path (through the C ABI).

Configs 1, 2 and the projects of config 5 are small enough for the CPU oracle at their FULL
BASELINE.json size, so they are compared directly.  Configs 3 and 4 are not (1.5e10 voice-frames,
1.1e10 vertex-frames); they are checked at larger slices than test_gpu_parity.py through properties
that do not need the oracle:
  * time sharding is invisible: N segments == one pass (bit-exact without an IIR);
  * power-of-two gain invariance: doubling every generator's gain leaves the normalized PCM
    bit-identical (every vertex is linear, and scaling by 2 commutes with f32 rounding);
  * sample_loop periodicity; normalize idempotence (flow B peak is exactly the scan's maximum);
  * batch == single: a project rendered inside a pipelined batch equals the same project alone.
Edge cases: empty flowws, zero-length renders, ragged buffer lengths (not a multiple of 4), hits on
chunk boundaries, fan-in wider than the inline-pointer limit, samples shorter than a chunk.
"""
import copy

import numpy as np
import pytest

from helpers import max_abs_diff, pcm_to_int, run_project
from termdaw_b200 import synthetic as S
from termdaw_b200.state import Assets, DecodedWav, Project, State

pytestmark = pytest.mark.gpu

TOL_IIR = 2e-5


def _pcm_close(a, b, bd, lsb):
    x, y = pcm_to_int(a, bd), pcm_to_int(b, bd)
    assert x.shape == y.shape, (x.shape, y.shape)
    d = int(np.abs(x - y).max()) if x.size else 0
    assert d <= lsb, d
    return d


# ------------------------------------------------------------------------------------ full-size parity
def test_config2_full_size(gpulib, oracle):
    """configs[1] at its full 60 s, 24 bit: every tap and the PCM against the oracle."""
    cfg = S.config_2(1.0)
    got, want = run_project(gpulib, *cfg), run_project(oracle, *cfg)
    assert got["cs"] == want["cs"] == 2813
    for name, (l, r) in want["taps"].items():
        peak = max(1.0, float(np.nanmax(np.abs(np.concatenate([l, r])))))
        d = max(max_abs_diff(got["taps"][name][0], l), max_abs_diff(got["taps"][name][1], r))
        assert d / peak <= TOL_IIR, (name, d, peak)
    _pcm_close(got["pcm"], want["pcm"], 24, 16)


@pytest.mark.parametrize("index", [0, 5, 11, 300, 4095])
def test_config5_full_size_projects(gpulib, oracle, index):
    """projects of configs[4] at their full 30 s (first, last and a few in between)."""
    cfg = S.config_5_project(index, 1.0)
    got, want = run_project(gpulib, *cfg, taps=False), run_project(oracle, *cfg, taps=False)
    assert got["cs"] == want["cs"] == 1407
    _pcm_close(got["pcm"], want["pcm"], 16, 1)
    for k, v in want["norm"].items():
        assert abs(got["norm"][k] - v) <= TOL_IIR * max(1.0, abs(v))


# ------------------------------------------------------------------------------------ properties
def _double_generator_gains(p: Project) -> Project:
    """gain -> 2 * effective gain, where the effective gain follows apply_gain's skip rule (sample.rs:109:
    |gain - 1| < 0.001 is not applied at all, i.e. acts as exactly 1)."""
    def eff(g):
        return 1.0 if abs(np.float32(g) - np.float32(1.0)) < np.float32(0.001) else g
    q = copy.deepcopy(p)
    for lst in ("sampleloops", "samplemultis", "samplelerps", "debugsines", "synths", "sampsyns"):
        setattr(q, lst, [(t[0], 2.0 * eff(t[1])) + tuple(t[2:]) for t in getattr(q, lst)])
    return q


@pytest.mark.parametrize("cfg_name", ["c3", "c4", "c5"])
def test_power_of_two_gain_invariance(gpulib, cfg_name):
    """x2 on every generator: all vertices are linear, so every buffer doubles exactly and the normalized
    PCM does not change by a single bit (no oracle involved: holds at any size)."""
    cfg = {"c3": lambda: S.config_3(0.05, voices=128), "c4": lambda: S.config_4(0.01, 64),
           "c5": lambda: S.config_5_project(7, 1.0)}[cfg_name]()
    p, a, s = cfg
    base = run_project(gpulib, p, a, s, taps=True)
    dbl = run_project(gpulib, _double_generator_gains(p), a, s, taps=True)
    assert np.array_equal(base["pcm"], dbl["pcm"])
    gens = [t[0] for lst in (p.sampleloops, p.samplemultis, p.samplelerps, p.synths, p.sampsyns) for t in lst]
    for name in gens:
        assert np.array_equal(base["taps"][name][0] * np.float32(2), dbl["taps"][name][0]), name


def test_time_sharding_is_invisible_config4(gpulib):
    """configs[3] slice (108 s, 64 vertices, 24 bit): 8 segments in one launch sequence against one pass.
    Everything but the band-pass hand-off is bit-exact, so the PCM may differ by the IIR tolerance only."""
    from termdaw_b200 import render_segments_local, scan_segments_local
    from termdaw_b200.timeshard import shard_chunks
    p, a, s = S.config_4(0.03, 64)
    whole = run_project(gpulib, p, a, s, taps=False)
    states = []
    for _ in range(8):
        st = State(gpulib, s["psr"], s["bl"], assets=a)
        assert st.refresh(p), st.messages
        states.append(st)
    cs = states[0].cs
    bounds = shard_chunks(cs, 8)
    ctxs = [st.ctx for st in states]
    scan_segments_local(gpulib, ctxs, cs, bounds)
    pcm = np.concatenate(render_segments_local(gpulib, ctxs, cs, bounds, states[0].bd))
    for st in states:
        for k, v in whole["norm"].items():
            assert abs(st.normalization_values()[k] - v) <= TOL_IIR * max(1.0, abs(v))
        st.close()
    _pcm_close(pcm, whole["pcm"], 24, 16)


def test_sample_loop_periodicity_and_normalize_peak(gpulib):
    """sample_loop output repeats with the sample's length for the whole render, and after a scan the
    normalized output's peak is exactly max * (1/max) (1.0 up to one rounding)."""
    p, a, s = S.config_1(4.0)
    p.set_render_samplerate(48000)
    got = run_project(gpulib, p, a, s)
    for name, L in (("one", 12000), ("two", 19200)):
        l, r = got["taps"][name]
        assert np.array_equal(l[:-L], l[L:]) and np.array_equal(r[:-L], r[L:]), name
    l, r = got["taps"]["sum"]
    peak = max(float(np.abs(l).max()), float(np.abs(r).max()))
    assert abs(peak - 1.0) <= 2 ** -23
    m = got["norm"]["sum"]
    raw = np.float32(m) * (np.float32(1.0) / np.float32(m))
    assert peak == float(raw)
    assert int(np.abs(pcm_to_int(got["pcm"], 16)).max()) in (32766, 32767)


def test_project_in_batch_equals_project_alone(gpulib):
    """Pipelined batch of 40 projects (tapered groups over two lanes) == each project rendered alone."""
    from termdaw_b200 import normalize_render_many
    cfgs = [S.config_5_project(100 + i, 0.05) for i in range(40)]
    states = []
    for p, a, s in cfgs:
        st = State(gpulib, s["psr"], s["bl"], assets=a)
        assert st.refresh(p), st.messages
        states.append(st)
    ctxs = [st.ctx for st in states]
    ctxs[0].set_pipeline_group(8)
    pcms = normalize_render_many(gpulib, ctxs, [st.cs for st in states], [st.render_sr for st in states],
                                 [st.bd for st in states])
    ctxs[0].set_pipeline_group(64)
    for (p, a, s), st, pcm in zip(cfgs, states, pcms):
        alone = run_project(gpulib, p, a, s, taps=False)
        assert np.array_equal(pcm, alone["pcm"])
        st.close()


# ------------------------------------------------------------------------------------ edge cases
def _tone(n, sr=48000):
    t = np.arange(n) / sr
    x = (np.sin(2 * np.pi * 220 * t) * 12000).astype(np.float32)
    return DecodedWav(np.repeat(x, 2), 2, sr, 16)


def _edge_assets():
    a = Assets()
    a.wavs["short"] = _tone(100)            # shorter than a chunk: wraps ~10 times per chunk
    a.wavs["tone"] = _tone(5000)
    a.blobs["table"] = S.wavetable_blob(9, 8, 64, 4.0)
    a.flowws["empty"] = (np.zeros(0, np.float32),) * 3
    # hits exactly on chunk boundaries (bl = 1000 at 48 kHz: t = k * 1000 / 48000) and one frame before
    fr = np.array([0, 999, 1000, 1999, 2000, 2001, 5000], np.float64)
    a.flowws["edges"] = ((fr / 48000.0).astype(np.float32), np.full(fr.size, 36, np.float32),
                         np.linspace(0.4, 1.0, fr.size).astype(np.float32))
    ev = []
    for t0, note, vel in ((0.0, 60, 1.0), (0.02, 64, 0.8), (0.05, 60, 0.9), (0.0625, 64, 0.7)):
        ev += [(t0, note, vel), (t0 + 0.015, note, 0.0)]         # on, then off 15 ms later (voices overlap)
    ev.sort(key=lambda e: e[0])
    arr = np.array(ev, np.float32)
    a.flowws["notes"] = (arr[:, 0].copy(), arr[:, 1].copy(), arr[:, 2].copy())
    return a


def _edge_project(seconds, wide_fan_in=False):
    p = Project()
    p.set_length(seconds)
    p.set_render_samplerate(48000)
    p.set_render_bitdepth(16)
    for nm in ("short", "tone"):
        p.load_sample(nm, nm, "")
    p.load_resource("table", "table")
    for nm in ("empty", "edges", "notes"):
        p.load_midi_floww(nm, nm)
    p.add_sampleloop("loop_short", 0.5, 10.0, "short")
    p.add_sample_multi("multi_edges", 0.7, -20.0, "tone", "edges", -1)
    p.add_sample_multi("multi_empty", 1.0, 0.0, "tone", "empty", -1)
    p.add_sample_lerp("lerp_edges", 0.9, 0.0, "tone", "edges", -1, 40)
    p.add_sample_lerp("lerp_empty", 1.0, 0.0, "short", "empty", 36, 0)
    p.add_synth("syn", 1.0, 0.0, "notes", 0.4, 0.3, S.HIT_ADSR, 1.0, 0.8, S.NOTE_ADSR, 0.5, S.NOTE_ADSR)
    p.add_synth("syn_empty", 1.0, 0.0, "empty", 0.4, 0.3, S.HIT_ADSR, 1.0, 0.8, S.NOTE_ADSR, 0.5, S.STD_ADSR)
    p.add_sampsyn("pad", 0.8, 0.0, "notes", S.STD_ADSR, "table")
    p.add_adsr("env", 1.0, 0.0, 1.0, "edges", False, True, -1, [0.01, 0.1, 0.8, 0.1, 0.2, 0.05])
    p.add_adsr("env_empty", 1.0, 0.0, 0.7, "empty", True, False, -1, [0.01, 0.1, 0.8, 0.1, 0.2, 0.05])
    p.add_bandpass("bp", 1.0, 0.0, 1.0, 1000.0, 50.0, True)
    p.add_sum("mix", 1.0, 0.0)
    p.add_normalize("out", 1.0, 0.0)
    gens = ["loop_short", "multi_edges", "multi_empty", "lerp_edges", "lerp_empty", "syn", "syn_empty", "pad"]
    for g in gens[:3]:
        p.connect(g, "env")
    for g in gens[3:5]:
        p.connect(g, "env_empty")
    for g in gens[5:]:
        p.connect(g, "bp")
    for f in ("env", "env_empty", "bp"):
        p.connect(f, "mix")
    if wide_fan_in:                               # 3 + 8 = 11 inputs: beyond the 8 inline pointer pairs
        for g in gens:
            p.connect(g, "mix")
    p.connect("mix", "out")
    p.set_output("out")
    return p


@pytest.mark.parametrize("bl", [1000, 1022, 1024, 37])
@pytest.mark.parametrize("wide", [False, True])
def test_edge_cases_against_oracle(gpulib, oracle, bl, wide):
    """ragged buffer lengths, empty flowws, boundary hits, tiny samples, wide fan-in: taps and PCM."""
    a = _edge_assets()
    p = _edge_project(0.15, wide)
    s = {"psr": 48000, "bl": bl, "flow": "B"}
    got, want = run_project(gpulib, p, a, s), run_project(oracle, p, a, s)
    assert got["cs"] == want["cs"]
    for name, (l, r) in want["taps"].items():
        peak = max(1.0, float(np.nanmax(np.abs(np.concatenate([l, r])))))
        d = max(max_abs_diff(got["taps"][name][0], l), max_abs_diff(got["taps"][name][1], r))
        if name in ("loop_short", "multi_edges", "multi_empty", "lerp_edges", "lerp_empty", "syn_empty"):
            assert d == 0.0, (name, d)
        else:
            assert d / peak <= TOL_IIR, (name, d, peak)
    _pcm_close(got["pcm"], want["pcm"], 16, 1)


def test_zero_length_render(gpulib, oracle):
    """set_length(0): no chunks, no frames, no error (state.rs:104 gives cs = 0)."""
    a = _edge_assets()
    p = _edge_project(0.0)
    s = {"psr": 48000, "bl": 1024, "flow": "B"}
    got, want = run_project(gpulib, p, a, s, taps=False), run_project(oracle, p, a, s, taps=False)
    assert got["cs"] == want["cs"] == 0
    assert got["pcm"].size == want["pcm"].size == 0
