"""Playback / stream feed: tdg_play renders `chunks` more chunks from the current
transport position as one device pass and rewinds nothing, tdg_floww_append feeds a declared stream while
the transport runs.  The reference's loop is one Graph::render per chunk (ui_workflow.rs:195-210,
stream_workflow.rs:63-100); the oracle restates exactly that, so every comparison here is "any batching of
the device passes == the oracle's chunk-at-a-time stream", all vertex state carried across the calls."""
import numpy as np
import pytest

from helpers import max_abs_diff
from termdaw_b200 import synthetic as S
from termdaw_b200.state import Project, State

pytestmark = pytest.mark.gpu
TOL = 2e-5


def _stream(lib, cfg, plan, scan=False, ops=()):
    p, a, s = cfg
    st = State(lib, s["psr"], s["bl"], assets=a)
    try:
        assert st.refresh(p), st.messages
        if scan:
            st.scan_exact()
        out = []
        for k, n in enumerate(plan):
            for at, op in ops:
                if at == k:
                    op(st)
            out.append(st.ctx.play(n))
        return np.concatenate(out), st.ctx.get_time()
    finally:
        st.close()


def _close(got, want):
    assert got.shape == want.shape
    peak = max(1.0, float(np.nanmax(np.abs(want))) if want.size else 1.0)
    assert max_abs_diff(got, want) <= TOL * peak


def check_play_equals_chunkwise_reference(lib, oracle, name):
    cfg = {"c1": lambda: S.config_1(0.2), "c2": lambda: S.config_2(0.03), "c4": lambda: S.config_4(0.0008, 64),
           "p0": lambda: S.config_5_project(0, 0.05), "p3": lambda: S.config_5_project(3, 0.05)}[name]()
    for scan in (False, True):
        want, tw = _stream(oracle, cfg, [1] * 24, scan)
        for plan in ([24], [1] * 24, [5, 1, 7, 11]):
            got, tg = _stream(lib, cfg, plan, scan)
            assert tg == tw == 24 * cfg[2]["bl"]
            _close(got, want)
            if name == "c1":                                  # no transcendental, no IIR: bit for bit
                assert np.array_equal(got, want)


def check_play_then_seek_and_render(lib, oracle):
    """transport commands between plays, and a normal render after playing (state.rs:575 rewinds; play does not)"""
    cfg = S.config_5_project(3, 0.05)
    ops = ((2, lambda st: st.ctx.set_time(3 * st.bl)), (2, lambda st: st.ctx.floww_set_time(3 * st.bl)))
    want, _ = _stream(oracle, cfg, [1] * 6, ops=((4, ops[0][1]), (4, ops[1][1])))
    got, _ = _stream(lib, cfg, [3, 1, 2], ops=ops)
    _close(got, want)
    for L, res in ((lib, []), (oracle, [])):
        p, a, s = cfg
        st = State(L, s["psr"], s["bl"], assets=a)
        assert st.refresh(p)
        st.ctx.play(5)
        st.ctx.set_time(0); st.ctx.floww_set_time(0)
        res.append(st.render())
        st.close()
        if L is lib:
            pcm_lib = res[0]
        else:
            from helpers import pcm_to_int
            x, y = pcm_to_int(pcm_lib, 16), pcm_to_int(res[0], 16)
            assert x.shape == y.shape and int(np.abs(x - y).max()) <= 1


def _stream_project():
    """examples/stream.lua in miniature: a synth and a sample_multi driven by declared streams"""
    base, a, s = S.config_5_project(0, 0.05)
    p = Project()
    p.set_length(2.0); p.set_render_samplerate(48000); p.set_render_bitdepth(16)
    p.load_sample("s_noise", "s_noise", "")
    p.declare_stream("keys"); p.declare_stream("drums")
    p.add_synth("syn", 0.8, 10.0, "keys", 0.4, 0.3, S.HIT_ADSR, 1.0, 0.6, S.NOTE_ADSR, 0.5, S.NOTE_ADSR)
    p.add_sample_multi("hit", 1.0, -20.0, "s_noise", "drums", -1)
    p.add_adsr("env", 1.0, 0.0, 1.0, "drums", False, True, -1, [0.01, 0.1, 0.8, 0.1, 0.2, 0.05])
    p.add_normalize("out", 1.0, 0.0)
    p.connect("syn", "env"); p.connect("env", "out"); p.connect("hit", "out")
    p.set_output("out")
    return p, a, s


def check_stream_feed(lib, oracle):
    cfg = _stream_project()
    rng = np.random.default_rng(9)
    feeds = []                                               # (before play call k, stream, events)
    t = 0.0
    for k in range(6):
        n = int(rng.integers(1, 5))
        ts = np.sort(rng.uniform(t, t + 0.3, n)).astype(np.float32)
        notes = rng.integers(40, 80, n).astype(np.float32)
        vel = np.where(rng.random(n) < 0.7, rng.uniform(0.3, 1.0, n), 0.0).astype(np.float32)
        feeds.append((k, "keys" if k % 2 == 0 else "drums", (ts, notes, vel)))
        t += 0.15

    def run(L, plan, when):
        p, a, s = cfg
        st = State(L, s["psr"], s["bl"], assets=a)
        assert st.refresh(p), st.messages
        out = []
        for k, n in enumerate(plan):
            for (at, name, ev) in feeds:
                if when(at) == k:
                    st.ctx.floww_append(st.floww_index[name], *ev)
            out.append(st.ctx.play(n))
        st.close()
        return np.concatenate(out)
    want = run(oracle, [1] * 48, lambda at: at * 8)          # the reference: a chunk at a time, fed every 8 chunks
    got = run(lib, [8] * 6, lambda at: at)                   # the device: 8 chunks per pass
    _close(got, want)
    assert float(np.abs(want).max()) > 0.01                  # it does make sound


@pytest.mark.parametrize("name", ["c1", "c2", "c4", "p0", "p3"])
def test_play_equals_chunkwise_reference(gpulib, oracle, name):
    check_play_equals_chunkwise_reference(gpulib, oracle, name)


def test_play_then_seek_and_render(gpulib, oracle):
    check_play_then_seek_and_render(gpulib, oracle)


def test_stream_feed(gpulib, oracle):
    check_stream_feed(gpulib, oracle)


def test_play_errors(gpulib):
    from termdaw_b200 import TermdawError
    p, a, s = S.config_1(0.05)
    st = State(gpulib, s["psr"], s["bl"], assets=a)
    assert st.refresh(p)
    with pytest.raises(TermdawError):                        # events before the stream's last event
        fid = st.ctx.floww_add([1.0], [60.0], [1.0])
        st.ctx.floww_append(fid, [0.5], [60.0], [1.0])
    with pytest.raises(TermdawError):
        st.ctx.floww_append(99, [0.5], [60.0], [1.0])
    assert st.ctx.play(0).size == 0
    st.close()


def check_async_play_ring(lib, oracle):
    """tdg_play_begin / tdg_play_wait: blocks queued ahead of the player (tickets), collected in order from the pinned
    ring == the synchronous tdg_play stream == the oracle's chunk-at-a-time loop; a render issued while playback work
    is still queued waits for it; the ring refuses a seventh uncollected ticket."""
    from termdaw_b200._capi import TermdawError
    cfg = S.config_5_project(3, 0.05)
    want, _ = _stream(oracle, cfg, [1] * 30)
    p, a, s = cfg
    st = State(lib, s["psr"], s["bl"], assets=a)
    try:
        assert st.refresh(p), st.messages
        plan = [4, 1, 7, 2, 11, 5]
        tickets = [st.ctx.play_begin(n) for n in plan]               # six blocks in flight, nothing collected yet
        assert st.ctx.get_time() == 30 * st.bl                        # the transport has already advanced
        with pytest.raises(TermdawError):
            st.ctx.play_begin(1)                                      # ring full
        got = np.concatenate([st.ctx.play_wait(t).reshape(-1) for t in tickets])
        _close(got, want)
        with pytest.raises(TermdawError):
            st.ctx.play_wait(tickets[0])                              # a ticket is collected once
        t = st.ctx.play_begin(2)                                      # ... and the ring turns over
        more = st.ctx.play_wait(t).reshape(-1)
        assert more.size == 2 * 2 * st.bl
        # a synchronous call after queued asynchronous ones continues the same stream
        t = st.ctx.play_begin(3)
        tail = st.ctx.play(2)
        head = st.ctx.play_wait(t).reshape(-1)
        want2, _ = _stream(oracle, cfg, [1] * 37)
        _close(np.concatenate([got, more, head, tail]), want2)
    finally:
        st.close()


def test_async_play_ring(gpulib, oracle):
    check_async_play_ring(gpulib, oracle)
