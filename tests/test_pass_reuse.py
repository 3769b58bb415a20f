"""Pass-invariant cones (tdg_set_pass_reuse): a render that directly follows a normalize scan
reads the buffers of vertices whose whole input cone Graph::set_time resets (extensions.rs:196-204) from the
scan instead of evaluating them again.  The values are the scan's own, so every test here demands
BIT-identical PCM with the switch on and off, the oracle's result within the usual tolerance, and that
anything that could make the two passes differ (an edit, a seek, a different length, another scan on the
same lane) turns the reuse off for that render."""
import numpy as np
import pytest

from helpers import pcm_to_int, run_project
from termdaw_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _cfg(name):
    if name == "c1":
        return S.config_1(0.5)
    if name == "c2":
        return S.config_2(0.05)
    if name == "c4":
        return S.config_4(0.002, 64)
    return S.config_5_project(int(name[1:]), 0.1)


def _state(lib, cfg):
    from termdaw_b200.state import State
    p, a, s = cfg
    st = State(lib, s["psr"], s["bl"], assets=a)
    assert st.refresh(p), st.messages
    return st


def _scan_render(gpulib, cfg, reuse):
    st = _state(gpulib, cfg)
    try:
        st.ctx.set_pass_reuse(reuse)
        st.scan_exact()
        pcm = st.render().copy()
        stats = st.ctx.last_pass_stats()
        return pcm, stats, st.normalization_values(), st.bd
    finally:
        st.ctx.set_pass_reuse(True)
        st.close()


def _n_invariant_generators(cfg):
    p = cfg[0]
    return len(p.sampleloops) + len(p.synths) + len(p.debugsines)


@pytest.mark.parametrize("name", ["c1", "c2", "c4", "p0", "p1", "p2", "p3", "p5", "p7", "p11", "p40"])
def test_reuse_is_bit_identical(gpulib, oracle, name):
    cfg = _cfg(name)
    on, s_on, n_on, bd = _scan_render(gpulib, cfg, True)
    off, s_off, n_off, _ = _scan_render(gpulib, cfg, False)
    assert np.array_equal(on, off), name
    assert n_on == n_off
    assert s_off["reused_vertices"] == 0 and s_off["reused_bytes"] == 0
    want = run_project(oracle, *cfg, taps=False)
    x, y = pcm_to_int(on, bd), pcm_to_int(want["pcm"], want["bd"])
    assert x.shape == y.shape and int(np.abs(x - y).max()) <= (1 if bd == 16 else 16)
    if name in ("c1", "c2"):
        # c1: both sample loops; c2: the synth (adsr, band pass and normalize carry state / depend on the scan)
        assert s_on["reused_vertices"] == (2 if name == "c1" else 1)
        assert s_on["algorithmic_bytes"] < s_off["algorithmic_bytes"]
    if name == "c4":
        assert s_on["reused_vertices"] >= 16 + 8           # at least the sample loops and the synths


def test_reuse_is_used_by_the_batched_calls(gpulib, oracle):
    from termdaw_b200 import normalize_render_many, normalize_scan_many, render_many
    cfgs = [S.config_5_project(i, 0.04) for i in range(9)]
    gens = sum(_n_invariant_generators(c) for c in cfgs)
    wants = [run_project(oracle, *c, taps=False) for c in cfgs]
    ref = None
    for mode in ("one-sequence", "pipelined-pairs", "pipelined-separate", "off"):
        states = [_state(gpulib, c) for c in cfgs]
        ctxs = [st.ctx for st in states]
        args = ([st.cs for st in states], [st.render_sr for st in states], [st.bd for st in states])
        try:
            ctxs[0].set_pass_reuse(mode != "off")
            if mode == "one-sequence" or mode == "off":
                ctxs[0].set_pipeline_group(0)
                normalize_scan_many(gpulib, ctxs, args[0])
                pcms = render_many(gpulib, ctxs, *args)
            elif mode == "pipelined-pairs":
                ctxs[0].set_pipeline_group(2)
                pcms = normalize_render_many(gpulib, ctxs, *args)
            else:
                # separately pipelined scans overwrite each other's kept buffers lane by lane: only the last
                # group of each lane can still be read, the others must notice and evaluate everything
                ctxs[0].set_pipeline_group(2)
                normalize_scan_many(gpulib, ctxs, args[0])
                pcms = render_many(gpulib, ctxs, *args)
            reused = ctxs[0].last_pass_stats()["reused_vertices"]
            if mode == "off":
                assert reused == 0
            elif mode == "pipelined-separate":
                assert 0 <= reused < gens
            else:
                assert reused >= gens, (mode, reused, gens)
            pcms = [p.copy() for p in pcms]
            if ref is None:
                ref = pcms
            for i, (st, pcm, want) in enumerate(zip(states, pcms, wants)):
                assert np.array_equal(pcm, ref[i]), (mode, i)
                x, y = pcm_to_int(pcm, st.bd), pcm_to_int(want["pcm"], want["bd"])
                assert x.shape == y.shape and np.abs(x - y).max() <= 1, (mode, i)
        finally:
            ctxs[0].set_pass_reuse(True)
            ctxs[0].set_pipeline_group(64)
            for st in states:
                st.close()


def _sequence(lib, cfg, ops, is_gpu):
    """run a list of transport operations; returns (pcms of the renders, reused vertex counts on the GPU)"""
    st = _state(lib, cfg)
    pcms, reused = [], []
    try:
        for op in ops:
            if op == "scan":
                st.scan_exact()
            elif op == "render":
                pcms.append(st.render().copy())
                if is_gpu:
                    reused.append(st.ctx.last_pass_stats()["reused_vertices"])
            elif op == "shorter":
                st.cs -= 1
            elif op[0] == "seek":
                st.ctx.set_time(op[1] * st.bl)
            elif op == "refresh":
                assert st.refresh(cfg[0])
        return pcms, reused, st.bd
    finally:
        st.close()


@pytest.mark.parametrize("ops,expect_reuse", [
    (["scan", "render"], [True]),
    (["render"], [False]),                                   # flow A: nothing to read
    (["scan", "render", "render"], [True, False]),           # the record is consumed by the first render
    (["scan", "scan", "render"], [True]),
    (["scan", ("seek", 3), "render"], [False]),              # loops now start at frame 3*bl
    (["scan", "shorter", "render"], [False]),                # another length
    (["scan", "refresh", "render"], [False]),
    (["scan", "render", "scan", "render"], [True, True]),
    (["render", "scan", "render"], [False, True]),
])
def test_reuse_only_when_the_passes_cannot_differ(gpulib, oracle, ops, expect_reuse):
    for cfg in (S.config_1(0.3), S.config_5_project(2, 0.05), S.config_5_project(7, 0.05)):
        if _n_invariant_generators(cfg) == 0:
            continue
        got, reused, bd = _sequence(gpulib, cfg, ops, True)
        want, _, _ = _sequence(oracle, cfg, ops, False)
        assert [r > 0 for r in reused] == expect_reuse, (ops, reused)
        for g, w in zip(got, want):
            x, y = pcm_to_int(g, bd), pcm_to_int(w, bd)
            assert x.shape == y.shape and int(np.abs(x - y).max()) <= 1, ops


def test_two_projects_interleaved_on_one_lane(gpulib, oracle):
    """scan A, scan B, render A, render B: B's scan lays its kept buffers over A's, so A evaluates everything and
    B reads its own — neither may ever read the other's audio."""
    ca, cb = S.config_1(0.3), S.config_5_project(2, 0.05)
    a, b = _state(gpulib, ca), _state(gpulib, cb)
    try:
        a.scan_exact(); b.scan_exact()
        pa = a.render().copy(); ra = a.ctx.last_pass_stats()["reused_vertices"]
        pb = b.render().copy(); rb = b.ctx.last_pass_stats()["reused_vertices"]
        assert ra == 0 and rb > 0
        for cfg, st, pcm in ((ca, a, pa), (cb, b, pb)):
            want = run_project(oracle, *cfg, taps=False)
            x, y = pcm_to_int(pcm, st.bd), pcm_to_int(want["pcm"], want["bd"])
            assert x.shape == y.shape and int(np.abs(x - y).max()) <= 1
    finally:
        a.close(); b.close()


def test_heads_only_reuse_of_sample_multi_and_lerp(gpulib, oracle):
    """sample_multi / sample_lerp are kept whole by the scan and the render redoes only the chunks in which state
    leaked from the pass before is audible: the render's algorithmic bytes for them shrink to a few chunks' worth,
    the result stays bit-identical to evaluate-everything-twice, also when the scan itself started from leaked state."""
    from termdaw_b200.state import Project, State
    base, a, s = S.config_5_project(0, 0.2)
    p = Project()
    p.set_length(6.0); p.set_render_samplerate(48000); p.set_render_bitdepth(16)
    for nm in ("s_noise", "s_tone", "s_long"):
        p.load_sample(nm, nm, "")
    p.load_midi_floww("hits0", "hits0"); p.load_midi_floww("hits1", "hits1")
    p.add_sample_multi("m", 0.9, 10.0, "s_long", "hits0", -1)
    p.add_sample_lerp("l", 0.8, -30.0, "s_tone", "hits1", -1, 200)
    p.add_normalize("out", 1.0, 0.0)
    p.connect("m", "out"); p.connect("l", "out"); p.set_output("out")
    cfg = (p, a, s)
    res = {}
    for reuse in (True, False):
        st = State(gpulib, s["psr"], s["bl"], assets=a)
        assert st.refresh(p), st.messages
        st.ctx.set_pass_reuse(reuse)
        out = []
        for _ in range(2):                                   # the second scan starts from what the first render leaked
            st.scan_exact()
            out.append(st.render().copy())
            stats = st.ctx.last_pass_stats()
        res[reuse] = (out, stats, st.cs)
        st.ctx.set_pass_reuse(True)
        st.close()
    (on, s_on, cs), (off, s_off, _) = res[True], res[False]
    for x, y in zip(on, off):
        assert np.array_equal(x, y)
    assert s_on["reused_vertices"] == 2 and s_off["reused_vertices"] == 0
    frames = cs * s["bl"]
    # off: multi 8 B out + gathers, lerp 16 B per frame, normalize 24, quantise 12; on: the two generators shrink to their heads
    assert s_off["algorithmic_bytes"] - s_on["algorithmic_bytes"] > 0.8 * (8 + 16) * frames
    want = run_project(oracle, *cfg, taps=False)
    x, y = pcm_to_int(on[0], 16), pcm_to_int(want["pcm"], 16)
    assert x.shape == y.shape and int(np.abs(x - y).max()) <= 1
