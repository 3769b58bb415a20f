"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star):
  * sample indexing / note timing: bit-exact -> graphs without transcendentals must match bit for bit
    before the IIR (sample_loop / sample_multi / sample_lerp / sampsyn / sum / adsr / normalize);
  * oscillators: the device evaluates glibc's sinf operation for operation (tdg_math.cuh m_sinf_glibc, pinned
    against the host libm over all 2^32 arguments by tests/test_sinf_restatement.py), every other f32
    operation keeps the reference's order with one rounding each -> EVERY vertex whose input cone holds no
    band pass must match the oracle bit for bit (`exact_outside_iir`), synth / sampsyn / debug_sine / adsr /
    normalize included, at any bit depth;
  * band pass: the carried-in state of each 8-frame run is a reassociated scan, everything else is
    the reference's operation order -> stated tolerance 2e-5 * max(1, peak) on the f32 buffers of the filter
    and of everything downstream of it (measured worst: 7.5e-7);
  * PCM: <= 1 LSB at 16 bit; at 24 bit bit-exact graphs are 0 LSB, graphs with a band pass <= 16 LSB
    (1.9e-6 of full scale; measured worst 6: one LSB there is 1.2e-7, below one f32 ulp of a full-scale sample).
"""
import json
import os

import numpy as np
import pytest

from helpers import ROOT, max_abs_diff, pcm_to_int, run_project
from termdaw_b200 import synthetic as S
from termdaw_b200.state import Assets, Project

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5
TOL_IIR = 2e-5


def iir_tainted(project):
    """Names of the band-pass vertices and of everything downstream of one (the only reassociated arithmetic)."""
    tainted = {t[0] for t in project.bandpasses}
    grew = True
    while grew:
        grew = False
        for a, b in project.edges:
            if a in tainted and b not in tainted:
                tainted.add(b); grew = True
    return tainted


def compare(got, want, label, exact_kinds=(), tol=TOL_F32, pcm_lsb=1, report=None, exact_outside_iir=None):
    assert got["cs"] == want["cs"]
    worst = {}
    if exact_outside_iir is not None:
        tainted = iir_tainted(exact_outside_iir)
        exact_kinds = tuple(exact_kinds) + tuple(n for n in want["taps"] if n not in tainted)
        if not tainted:
            pcm_lsb = 0
    for name, (l, r) in want["taps"].items():
        gl, gr = got["taps"][name]
        peak = max(1.0, float(np.nanmax(np.abs(np.concatenate([l, r])))))
        d = max(max_abs_diff(gl, l), max_abs_diff(gr, r))
        worst[name] = d / peak
        if any(k == name for k in exact_kinds) or any(k in name for k in exact_kinds if exact_outside_iir is None):
            assert np.array_equal(gl, l, equal_nan=True) and np.array_equal(gr, r, equal_nan=True), \
                (label, name, "expected bit-exact", d)
        else:
            assert d / peak <= tol, (label, name, d, peak)
    a, b = pcm_to_int(got["pcm"], got["bd"]), pcm_to_int(want["pcm"], want["bd"])
    assert a.shape == b.shape, (label, a.shape, b.shape)
    lsb = int(np.abs(a - b).max()) if a.size else 0
    if report is not None:
        report[label] = {"worst_rel_fp32": max(worst.values()) if worst else 0.0, "pcm_lsb": lsb,
                         "per_vertex": worst}
    assert lsb <= pcm_lsb, (label, lsb)
    for k, v in want["norm"].items():
        assert abs(got["norm"][k] - v) <= tol * max(1.0, abs(v)), (label, k, got["norm"][k], v)
    return worst, lsb


@pytest.fixture(scope="module")
def report():
    rep = {}
    yield rep
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(rep, f, indent=1, sort_keys=True)


def test_config1_readme_example(gpulib, oracle, report):
    """2 x sampleloop -> normalize, 3 s, sinc down-sampled to 44.1 kHz, 16 bit (configs[0])."""
    cfg = S.config_1(1.0)
    got, want = run_project(gpulib, *cfg), run_project(oracle, *cfg)
    compare(got, want, "config1", report=report, exact_outside_iir=cfg[0])      # resampler positions are the reference's sequential f64 sums


def test_config1_flow_a_running_max(gpulib, oracle, report):
    """refresh -> render without a scan: normalize uses a running max (extensions.rs:326-328)."""
    cfg = S.config_1(0.5)
    got, want = run_project(gpulib, *cfg, flow="A"), run_project(oracle, *cfg, flow="A")
    compare(got, want, "config1_flowA", report=report, exact_outside_iir=cfg[0])


def test_config2_synth_adsr_bandpass(gpulib, oracle, report):
    """floww-driven synth + adsr -> band pass -> normalize, 24 bit (configs[1]), 6 s slice."""
    cfg = S.config_2(0.1)
    got, want = run_project(gpulib, *cfg), run_project(oracle, *cfg)
    compare(got, want, "config2", tol=TOL_IIR, pcm_lsb=16, report=report, exact_outside_iir=cfg[0])


def test_config3_wavetable_lerp(gpulib, oracle, report):
    """wavetable voices + 2 sample_lerp from 44.1/96 kHz sources (configs[2]) scaled to 64 voices, 6 s."""
    cfg = S.config_3(0.01, voices=64)
    got, want = run_project(gpulib, *cfg), run_project(oracle, *cfg)
    compare(got, want, "config3", report=report, exact_outside_iir=cfg[0])


def test_config4_mixed_graph(gpulib, oracle, report):
    """64-vertex mixed graph (configs[3]), 7 s slice, 24 bit."""
    cfg = S.config_4(0.002, 64)
    got, want = run_project(gpulib, *cfg), run_project(oracle, *cfg)
    compare(got, want, "config4", tol=TOL_IIR, pcm_lsb=16, report=report, exact_outside_iir=cfg[0])


@pytest.mark.parametrize("index", [0, 1, 2, 3, 4, 5, 6, 7])
def test_config5_batch_projects(gpulib, oracle, report, index):
    """projects of the algorithmic-composition batch (configs[4]), 3 s each, 16 bit."""
    cfg = S.config_5_project(index, 0.1)
    got, want = run_project(gpulib, *cfg), run_project(oracle, *cfg)
    compare(got, want, f"config5_{index}", tol=TOL_IIR, report=report, exact_outside_iir=cfg[0])


def test_render_many_equals_single(gpulib, oracle):
    """tdg_render_many over 6 projects == the oracle's per-project renders (scan + render)."""
    from termdaw_b200 import normalize_scan_many, render_many
    from termdaw_b200.state import State
    cfgs = [S.config_5_project(i, 0.05) for i in range(6)]
    states = []
    for p, a, s in cfgs:
        st = State(gpulib, s["psr"], s["bl"], assets=a)
        assert st.refresh(p), st.messages
        states.append(st)
    ctxs = [st.ctx for st in states]
    normalize_scan_many(gpulib, ctxs, [st.cs for st in states])
    pcms = render_many(gpulib, ctxs, [st.cs for st in states], [st.render_sr for st in states],
                       [st.bd for st in states])
    for (p, a, s), st, pcm in zip(cfgs, states, pcms):
        want = run_project(oracle, p, a, s, taps=False)
        x, y = pcm_to_int(pcm, st.bd), pcm_to_int(want["pcm"], want["bd"])
        assert x.shape == y.shape and np.abs(x - y).max() <= 1
        st.close()


@pytest.mark.parametrize("group", [0, 2, 3])
def test_pipelined_batches_equal_oracle(gpulib, oracle, group):
    """tdg_normalize_render_many / tdg_*_many with software pipelining (groups of `group` projects over two
    streams, two plan slots each; 0 = one launch sequence) == the oracle's per-project scan + render."""
    from termdaw_b200 import normalize_render_many, normalize_scan_many, render_many
    from termdaw_b200.state import State
    cfgs = [S.config_5_project(i, 0.04) for i in range(11)]
    wants = [run_project(oracle, p, a, s, taps=False) for p, a, s in cfgs]
    for combined in (True, False):
        states = []
        for p, a, s in cfgs:
            st = State(gpulib, s["psr"], s["bl"], assets=a)
            assert st.refresh(p), st.messages
            states.append(st)
        ctxs = [st.ctx for st in states]
        ctxs[0].set_pipeline_group(group)
        args = ([st.cs for st in states], [st.render_sr for st in states], [st.bd for st in states])
        if combined:
            pcms = normalize_render_many(gpulib, ctxs, *args)
        else:
            normalize_scan_many(gpulib, ctxs, args[0])
            pcms = render_many(gpulib, ctxs, *args)
        ctxs[0].set_pipeline_group(64)
        for st, pcm, want in zip(states, pcms, wants):
            x, y = pcm_to_int(pcm, st.bd), pcm_to_int(want["pcm"], want["bd"])
            assert x.shape == y.shape and np.abs(x - y).max() <= 1, (group, combined)
            for k, v in want["norm"].items():
                assert abs(st.normalization_values()[k] - v) <= TOL_IIR * max(1.0, abs(v))
            st.close()


def test_pipelined_ramp_and_taper_groups(gpulib, oracle):
    """41 projects in groups of 16: the call is cut into a ramp group of 8, full groups and a tapered tail
    (run_scan_render_pipelined); every project must still equal the oracle, and the batch with one launch
    sequence, bit for bit."""
    from termdaw_b200 import normalize_render_many
    from termdaw_b200.state import State
    cfgs = [S.config_5_project(i, 0.02) for i in range(41)]
    got = {}
    for group in (16, 0):
        states = []
        for p, a, s in cfgs:
            st = State(gpulib, s["psr"], s["bl"], assets=a)
            assert st.refresh(p), st.messages
            states.append(st)
        ctxs = [st.ctx for st in states]
        ctxs[0].set_pipeline_group(group)
        got[group] = [x.copy() for x in normalize_render_many(gpulib, ctxs, [st.cs for st in states],
                                                              [st.render_sr for st in states], [st.bd for st in states])]
        ctxs[0].set_pipeline_group(64)
        for st in states:
            st.close()
    for i, (p, a, s) in enumerate(cfgs):
        assert np.array_equal(got[16][i], got[0][i]), i
        want = run_project(oracle, p, a, s, taps=False)
        x, y = pcm_to_int(got[16][i], want["bd"]), pcm_to_int(want["pcm"], want["bd"])
        assert x.shape == y.shape and np.abs(x - y).max() <= 1, i


@pytest.mark.parametrize("bd", [8, 16, 24, 32])
def test_bitdepths(gpulib, oracle, bd):
    """quantisation to every supported depth (state.rs:515-532) on a bit-exact graph."""
    p, a, s = S.config_1(0.3)
    p.set_render_samplerate(48000)
    p.set_render_bitdepth(bd)
    got, want = run_project(gpulib, p, a, s), run_project(oracle, p, a, s)
    assert np.array_equal(got["pcm"], want["pcm"])


def test_unsupported_and_errors(gpulib):
    from termdaw_b200 import Context, TermdawError
    c = Context(gpulib, 48000, 1024)
    with pytest.raises(TermdawError) as e:
        c.add_lv2fx(1.0, 0.0, 1.0, 0)
    assert e.value.code == -3
    with pytest.raises(TermdawError) as e:
        c.add_adsr(1.0, 0.0, 1.0, c.floww_add([], [], []), False, True, -1, [1, 2, 3])
    assert e.value.code == -4
    c.close()


def test_refresh_cycle_is_idempotent(gpulib, oracle):
    """refresh -> scan -> render twice on one context: with tdg_reset_vertex_state in between (what
    State::refresh does to an unchanged project, state.rs:328-467) the second render is byte-identical to
    the first and to the oracle's fresh render; without it vertex state leaks and the two
    differ for a graph with stateful vertices."""
    from termdaw_b200.state import State
    p, a, s = S.config_5_project(3, 0.1)
    want = run_project(oracle, p, a, s, taps=False)
    st = State(gpulib, s["psr"], s["bl"], assets=a)
    assert st.refresh(p), st.messages
    st.scan_exact()
    first = st.render().copy()
    st.ctx.reset_vertex_state()
    st.ctx.floww_set_time(0)
    st.ctx.reset_normalize()
    st.scan_exact()
    second = st.render().copy()
    assert np.array_equal(first, second)
    x, y = pcm_to_int(first, st.bd), pcm_to_int(want["pcm"], want["bd"])
    assert x.shape == y.shape and np.abs(x - y).max() <= 1
    st.close()
