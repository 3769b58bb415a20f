"""Parity holes named by the round-1 review, closed on the device (every test calls through the C ABI and
compares with the CPU oracle; reference lines cited per test):

  (a) debug_sine: same-note re-strike, note-off, set_time clear           extensions.rs:423-457, 196-204
  (b) every SampleLoadMethod on stereo / odd-length / mono / silent input sample.rs:38-77,125-147,262-310
  (c) the five ADSR known-answer tests driven through the CUDA envelopes  adsr.rs:121-204
  (d) hound TooWide at 8/24 bit, NaN -> 0, saturation at 16/32 bit        state.rs:515-532
  (e) angle / wet clamps of Vertex::new                                   graph.rs:255-256
  (f) late timeline (t > 2^24 frames, full 3600 s) and 512-voice polyphony extensions.rs:450-452,501,525-527

Tolerances as in test_gpu_parity.py: bit-exact where no transcendental and no IIR is involved; oscillator
graphs within 1e-5 of max(1, peak) on the f32 taps and <= 1 LSB at 16 bit.
"""
import os

import numpy as np
import pytest

from helpers import max_abs_diff, pcm_to_int, run_project
from termdaw_b200 import synthetic as S
from termdaw_b200._capi import Context, TermdawError
from termdaw_b200.state import Assets, DecodedWav, Project, State

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5
EMULATED = bool(os.environ.get("TDG_TEST_LIB"))          # tests/emu: same bodies, shorter timelines
PSR, BL = 48000, 1024


def _f32(*xs):
    return np.asarray(xs, np.float32)


def _events(ev):
    a = np.array(sorted(ev, key=lambda e: e[0]), np.float32).reshape(-1, 3)
    return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy()


def _ones_wav(frames=4800, value=12000.0):
    """A sample whose every value is one constant: SampleBank::add normalises it to exactly 1.0."""
    lin = np.full(frames * 2, value, np.float32)
    return DecodedWav(lin, 2, PSR, 16)


# ------------------------------------------------------------------------------------ (a) debug_sine
def _debug_sine_project(seconds=1.5):
    a = Assets()
    a.flowws["n"] = _events([
        (0.05, 60.0, 0.5),            # voice A
        (0.10, 64.0, 0.7),            # voice B
        (0.25, 60.0004, 0.9),         # |dnote| < 0.001: A's velocity is replaced, no new voice (extensions.rs:433-438)
        (0.25, 67.0, 0.3),            # same frame as the re-strike: voice C
        (0.40, 64.0, 0.0),            # note-off removes B (retain, extensions.rs:444)
        (0.40, 64.0, 0.6),            # ... and a note-on in the same frame pushes it again, behind C
        (0.70, 60.002, 0.4),          # |dnote| > 0.001: a fourth voice
        (0.90, 60.0, 0.0),            # removes A only
        (1.20, 72.0, 0.8),            # still sounding when the pass ends: Graph::set_time clears it (extensions.rs:199)
    ])
    p = Project()
    p.set_length(seconds)
    p.set_render_samplerate(PSR)
    p.set_render_bitdepth(16)
    p.load_midi_floww("n", "n")
    p.add_debug_sine("ds", 0.7, 25.0, "n")
    p.add_normalize("out", 1.0, 0.0)
    p.connect("ds", "out")
    p.set_output("out")
    return p, a, {"psr": PSR, "bl": BL, "flow": "B"}


def _assert_close_taps(got, want, tol=TOL_F32):
    worst = (0.0, None, -1)
    for name, (l, r) in want["taps"].items():
        gl, gr = got["taps"][name]
        peak = max(1.0, float(np.nanmax(np.abs(np.concatenate([l, r])))))
        for g, w in ((gl, l), (gr, r)):
            d = np.abs(np.asarray(g, np.float64) - np.asarray(w, np.float64))
            d[np.isnan(g) & np.isnan(w)] = 0.0
            i = int(np.nanargmax(d)) if d.size else 0
            if d.size and d[i] / peak > worst[0]:
                worst = (float(d[i] / peak), name, i)
    assert worst[0] <= tol, worst
    return worst


def test_debug_sine_restrike_noteoff_and_clear(gpulib, oracle):
    """add_debug_sine through scan + render: dedupe by |dnote| < 0.001 replaces the velocity, note-off removes,
    the voice left over by the scan pass is cleared by set_time before the render (extensions.rs:423-457)."""
    cfg = _debug_sine_project()
    got, want = run_project(gpulib, *cfg), run_project(oracle, *cfg)
    _assert_close_taps(got, want)
    a, b = pcm_to_int(got["pcm"], 16), pcm_to_int(want["pcm"], 16)
    assert a.shape == b.shape and np.abs(a - b).max() <= 1
    assert float(np.abs(want["taps"]["ds"][0]).max()) > 0.5         # the graph really sounds


def test_debug_sine_flow_a_twice_and_playback(gpulib, oracle):
    """render without scan, twice (the second starts from cleared voices and equals the first), then the
    playback feed in uneven slices with the voices carried from call to call (ui_workflow.rs:195-210)."""
    p, a, s = _debug_sine_project()
    outs = {}
    for name, lib in (("gpu", gpulib), ("cpu", oracle)):
        st = State(lib, s["psr"], s["bl"], assets=a)
        assert st.refresh(p), st.messages
        first = st.render().copy()
        # State::render leaves the graph at t = 0 with its voices cleared (state.rs:575) but neither rewinds the
        # floww cursor nor the running normalize max: 'stop' (fb.set_time(0)) + a normalize reset do
        st.ctx.floww_set_time(0)
        st.ctx.reset_normalize()
        second = st.render().copy()
        assert np.array_equal(first, second), name
        st.ctx.floww_set_time(0)
        plays = [st.ctx.play(n) for n in (([3, 1, 7, 20, 40] if name == "gpu" else [1] * 71))]
        outs[name] = (first, np.concatenate(plays))
        st.close()
    x, y = pcm_to_int(outs["gpu"][0], 16), pcm_to_int(outs["cpu"][0], 16)
    assert x.shape == y.shape and np.abs(x - y).max() <= 1
    g, w = outs["gpu"][1], outs["cpu"][1]
    assert g.shape == w.shape and max_abs_diff(g, w) <= TOL_F32 * max(1.0, float(np.abs(w).max()))


# ------------------------------------------------------------------------------------ (b) sample load modes
def _wav_inputs():
    rng = np.random.default_rng(77)
    st_even = np.round(rng.uniform(-30000, 30000, 2 * 3001)).astype(np.float32)
    st_even[1::2] *= np.float32(0.25)                                  # right clearly quieter: 'loudest' picks left
    st_odd = np.round(rng.uniform(-20000, 20000, 2 * 2000 + 1)).astype(np.float32)
    st_right_loud = st_even.copy(); st_right_loud[0::2] *= np.float32(0.1)
    mono = np.round(rng.uniform(-9000, 9000, 2500)).astype(np.float32)
    silent = np.zeros(2 * 1200, np.float32)
    return {"stereo": (st_even, 2), "stereo_right_loud": (st_right_loud, 2), "stereo_odd": (st_odd, 2),
            "mono": (mono, 1), "silent": (silent, 2)}


@pytest.mark.parametrize("mode", ["", "left", "right", "loudest", "normalize-seperate", "mix-down"])
@pytest.mark.parametrize("src_sr", [48000, 44100])
def test_sample_load_modes(gpulib, oracle, mode, src_sr):
    """SampleBank::add's numeric part for every SampleLoadMethod (sample.rs:38-77,125-147,262-310): channel
    selection, the odd trailing sample, mono files, normalize / normalize_seperate / mix_down, a silent sample
    (1/0 = inf, 0*inf = NaN), then the load-time resample (sinc positions accumulated like the reference's
    `idx += t_ratio`: k_resample_index).  Same error or bit-identical buffers."""
    from termdaw_b200.state import sample_mode_from_string
    m = sample_mode_from_string(mode)
    for label, (lin, ch) in _wav_inputs().items():
        res = {}
        for name, lib in (("gpu", gpulib), ("cpu", oracle)):
            c = Context(lib, PSR, BL)
            try:
                i = c.sample_add(lin, ch, src_sr, m)
                res[name] = ("ok", c.sample_frames(i), c.sample_read(i))
            except TermdawError as e:
                res[name] = ("err", e.code, None)
            finally:
                c.close()
        assert res["gpu"][0] == res["cpu"][0], (label, mode, res["gpu"][:2], res["cpu"][:2])
        assert res["gpu"][1] == res["cpu"][1], (label, mode, res["gpu"][:2], res["cpu"][:2])
        if res["cpu"][0] == "ok":
            for g, w in zip(res["gpu"][2], res["cpu"][2]):
                assert np.array_equal(g, w, equal_nan=True), (label, mode, src_sr)


def test_sample_load_modes_expected_shapes(gpulib):
    """The decisions themselves (not only agreement with the oracle): lengths and channel identities."""
    inp = _wav_inputs()
    c = Context(gpulib, PSR, BL)
    try:
        lin, ch = inp["stereo_odd"]                                    # l gets the trailing sample (sample.rs:288-290)
        i = c.sample_add(lin, ch, PSR, 1)                              # left
        assert c.sample_frames(i) == 2001
        l, r = c.sample_read(i)
        assert np.array_equal(l, r) and abs(float(np.abs(l).max()) - 1.0) < 1e-6
        i = c.sample_add(lin, ch, PSR, 2)                              # right
        assert c.sample_frames(i) == 2000
        with pytest.raises(TermdawError):                              # stereo needs equal lengths (sample.rs:64-69)
            c.sample_add(lin, ch, PSR, 0)
        lin, ch = inp["stereo_right_loud"]
        i = c.sample_add(lin, ch, PSR, 3)                              # loudest -> right channel on both sides
        l, r = c.sample_read(i)
        want = lin[1::2] * (np.float32(1.0) / np.abs(lin[1::2]).max())
        assert np.array_equal(l, want) and np.array_equal(r, want)
        lin, ch = inp["stereo"]
        i = c.sample_add(lin, ch, PSR, 4)                              # normalize-seperate: each channel peaks at 1
        l, r = c.sample_read(i)
        assert abs(float(np.abs(l).max()) - 1.0) < 1e-6 and abs(float(np.abs(r).max()) - 1.0) < 1e-6
        i = c.sample_add(lin, ch, PSR, 5)                              # mix-down: l == r == (l + r) / max
        l, r = c.sample_read(i)
        mix = lin[0::2] + lin[1::2]
        assert np.array_equal(l, r) and np.array_equal(l, mix * (np.float32(1.0) / np.abs(mix).max()))
        lin, ch = inp["mono"]
        for m in (0, 4, 5):                                            # mono files only load as left/right/loudest
            with pytest.raises(TermdawError):
                c.sample_add(lin, ch, PSR, m)
        for m in (1, 2, 3):
            assert c.sample_frames(c.sample_add(lin, ch, PSR, m)) == lin.size
    finally:
        c.close()


def test_sample_modes_through_a_project(gpulib, oracle):
    """load_sample(name, path, mode) for every mode feeding sample_loop vertices of one graph: bit-exact taps."""
    inp = _wav_inputs()
    a = Assets()
    a.wavs["st"] = DecodedWav(inp["stereo"][0], 2, PSR, 16)
    a.wavs["odd"] = DecodedWav(inp["stereo_odd"][0], 2, PSR, 16)
    a.wavs["mono"] = DecodedWav(inp["mono"][0], 1, PSR, 16)
    p = Project()
    p.set_length(0.5); p.set_render_samplerate(PSR); p.set_render_bitdepth(16)
    loads = [("s0", "st", ""), ("s1", "st", "left"), ("s2", "st", "right"), ("s3", "st", "loudest"),
             ("s4", "st", "normalize-seperate"), ("s5", "st", "mix-down"), ("s6", "odd", "left"),
             ("s7", "odd", "loudest"), ("s8", "mono", "right"), ("s9", "mono", "loudest")]
    p.add_normalize("out", 1.0, 0.0)
    for k, (n, f, m) in enumerate(loads):
        p.load_sample(n, f, m)
        p.add_sampleloop("v" + n, 0.5 + 0.1 * k, -40.0 + 9.0 * k, n)
        p.connect("v" + n, "out")
    p.set_output("out")
    cfg = (p, a, {"psr": PSR, "bl": BL, "flow": "B"})
    got, want = run_project(gpulib, *cfg), run_project(oracle, *cfg)
    for name, (l, r) in want["taps"].items():
        assert np.array_equal(got["taps"][name][0], l) and np.array_equal(got["taps"][name][1], r), name
    assert np.array_equal(got["pcm"], want["pcm"])


# ------------------------------------------------------------------------------------ (c) ADSR KATs on the device
HIT_1 = [1.0, 1.0, 0.5, 1.0, 0.25, 1.0]                # hit_conf(1, 1, 0.5, 1, 0.25, 1)    adsr.rs:123
HIT_2 = [1.0, 1.0, 0.5, 2.0, 0.25, 1.0]                # hit_conf(1, 1, 0.5, 2, 0.25, 1)    adsr.rs:155
DIP_9 = [1.0, 1.0, 0.0, 0.5, 0.0, 0.5, 0.0, 1.0, 1.0]  # the inverted 9-float conf          adsr.rs:185-195
KAT_TOL = 1e-3                                          # the reference's own assert tolerance (adsr.rs:117-119)


def _adsr_vertex_project(conf, use_off, events, seconds):
    """constant 1.0 -> add_adsr(wet 1): the output IS the envelope value of every frame."""
    a = Assets()
    a.wavs["one"] = _ones_wav()
    a.flowws["f"] = _events(events)
    p = Project()
    p.set_length(seconds); p.set_render_samplerate(PSR); p.set_render_bitdepth(16)
    p.load_sample("one", "one", "")
    p.load_midi_floww("f", "f")
    p.add_sampleloop("src", 1.0, 0.0, "one")
    p.add_adsr("env", 1.0, 0.0, 1.0, "f", use_off, True, -1, conf)
    p.connect("src", "env")
    p.set_output("env")
    return p, a, {"psr": PSR, "bl": BL, "flow": "A"}


def _env_at(tap, t):
    return float(tap[int(round(t * PSR))])


def _check_kat(gpulib, oracle, cfg, vertex, points, shift=0.0):
    got, want = run_project(gpulib, *cfg), run_project(oracle, *cfg)
    gl, gr = got["taps"][vertex]
    wl, wr = want["taps"][vertex]
    # the envelope arithmetic has no transcendental and no reassociation: the device must equal the oracle exactly
    assert np.array_equal(gl, wl) and np.array_equal(gr, wr)
    for t, v in points:
        assert abs(_env_at(gl, t + shift) - v) <= KAT_TOL, (t, _env_at(gl, t + shift), v)
        assert abs(_env_at(wl, t + shift) - v) <= KAT_TOL, ("oracle", t, v)
    return gl


def test_adsr_kat_0_apply_adsr_on_device(gpulib, oracle):
    """adsr_0 (adsr.rs:121-133) through k_fx<adsr>, drum-hit mode (apply_adsr)."""
    cfg = _adsr_vertex_project(HIT_1, False, [(0.0, 60.0, 1.0)], 8.5)
    _check_kat(gpulib, oracle, cfg, "env",
               [(0.0, 0.0), (0.5, 0.5), (1.0, 1.0), (1.5, 0.75), (2.0, 0.5), (2.5, 0.375), (3.0, 0.25), (3.5, 0.125),
                (4.0, 0.0), (8.0, 0.0)])


def test_adsr_kat_1_and_2_apply_ads_then_apply_r_on_device(gpulib, oracle):
    """adsr_1 / adsr_2 (adsr.rs:136-165) in note-on/off mode.  Held: apply_ads of the primary.  Release: the
    vertex releases the GHOST slot first, and only when its stored release value apply_ads(conf, 0) * vel =
    std_vel * vel is non-zero does a later note-off reach the primary (extensions.rs:612-618) — so apply_r is
    reachable with a 9-float conf only.  on, on, off with std_vel = the KAT's old value (0.25 / 0.375), vel 1 and
    use_max = false (min) exposes apply_r(t, old) of the ghost: old at the note-off, old/2 half a second later, 0
    after release_sec."""
    for conf, held in ((HIT_1, [(0.5, 0.5), (1.0, 1.0), (1.5, 0.75), (2.0, 0.5), (2.5, 0.375), (3.0, 0.25), (7.0, 0.25)]),
                       (HIT_2, [(0.5, 0.5), (1.0, 1.0), (1.5, 0.75), (2.0, 0.5), (3.0, 0.375)])):
        cfg = _adsr_vertex_project(conf, True, [(0.0, 60.0, 1.0)], 7.5)
        _check_kat(gpulib, oracle, cfg, "env", held)
        # a lone note-off releases the (silent) ghost, never the primary: the held curve is unchanged
        cfg = _adsr_vertex_project(conf, True, [(0.0, 60.0, 1.0), (1.25, 60.0, 0.0)], 7.5)
        _check_kat(gpulib, oracle, cfg, "env", held)
    for old in (0.25, 0.375):
        conf9 = [old, 1.0, 1.0, 0.5, 0.5, 1.0, 0.25, 1.0, 0.0]
        p, a, s = _adsr_vertex_project(conf9, True, [(0.0, 60.0, 1.0), (1.0, 60.0, 1.0), (2.0, 60.0, 0.0)], 11.5)
        p.adsrs[0] = p.adsrs[0][:6] + (False,) + p.adsrs[0][7:]          # use_max = false
        _check_kat(gpulib, oracle, (p, a, s), "env", [(2.0, old), (2.5, old / 2), (3.0, 0.0), (11.0, 0.0)])


def test_adsr_kat_4_inverted_nine_float_conf_on_device(gpulib, oracle):
    """adsr_4 (adsr.rs:183-204): std_vel 1 dipping to 0 and back (examples/neg-adsr-env-example.lua) through
    k_fx<adsr>; the hit comes at 0.25 s so that the un-hit state before it is covered too."""
    cfg = _adsr_vertex_project(DIP_9, False, [(0.25, 60.0, 1.0)], 8.5)
    _check_kat(gpulib, oracle, cfg, "env",
               [(0.0, 1.0), (0.5, 0.5), (1.0, 0.0), (1.5, 0.0), (2.0, 0.0), (2.5, 0.5), (3.0, 1.0), (4.0, 1.0), (8.0, 1.0)],
               shift=0.25)


def _synth_env_project(conf, events, seconds):
    """A synth whose only oscillator is a square wave clipped at z = 1e-4: |output| = vel * env wherever
    |sin| >= z, so the voice kernel's envelope is observable (apply_ads held, apply_r_rt released)."""
    a = Assets()
    a.flowws["f"] = _events(events)
    p = Project()
    p.set_length(seconds); p.set_render_samplerate(PSR); p.set_render_bitdepth(16)
    p.load_midi_floww("f", "f")
    p.add_synth("syn", 1.0, 0.0, "f", 1.0, 0.0001, conf, 0.0, 0.5, [], 0.0, [])
    p.set_output("syn")
    return p, a, {"psr": PSR, "bl": BL, "flow": "A"}


def _abs_env_near(tap, t, width=64):
    i = int(round(t * PSR))
    w = np.abs(tap[max(0, i - width): i + width])
    return float(np.median(w))


@pytest.mark.parametrize("conf,nine", [(HIT_2, False), (DIP_9, True)])
def test_adsr_kat_3_apply_r_rt_through_a_synth_voice(gpulib, oracle, conf, nine):
    """adsr_3 (adsr.rs:168-180): apply_r_rt(t, rt = 3.0) inside k_voice<synth> — note held for 3 s, then the
    release.  Also the 9-float dip conf through the voice kernel (held part of adsr_4)."""
    t_off = 3.0
    cfg = _synth_env_project(conf, [(0.0, 57.0, 1.0), (t_off, 57.0, 0.0)], 4.5)
    got, want = run_project(gpulib, *cfg), run_project(oracle, *cfg)
    g, w = got["taps"]["syn"][0], want["taps"]["syn"][0]
    assert max_abs_diff(g, w) <= 2e-5            # the square edge |sin| ~ z amplifies the sin difference by 1/z * 1e-7
    mx = 1.0                                      # amp = 1 / (volume * max_vel): both confs peak at 1
    if not nine:
        pts = [(0.5, 0.5), (1.0, 1.0), (1.5, 0.75), (2.0, 0.5), (2.9, 0.3875),
               (t_off + 0.0, 0.375), (t_off + 0.5, 0.1875), (t_off + 0.99, 0.00375)]
    else:
        pts = [(0.5, 0.5), (0.99, 0.01), (1.5, 0.0), (2.5, 0.0), (t_off + 0.5, 0.5), (t_off + 0.99, 0.99)]
    for t, v in pts:
        assert abs(_abs_env_near(g, t) / mx - v) <= 2e-3, (t, _abs_env_near(g, t), v)
        assert abs(_abs_env_near(w, t) / mx - v) <= 2e-3, ("oracle", t, v)


# ------------------------------------------------------------------------------------ (d) quantiser edge cases
def _loud_project(bd, gain):
    p, a, s = S.config_1(0.2)
    p.set_render_samplerate(PSR)
    p.set_render_bitdepth(bd)
    p.norms[0] = ("sum", gain, 0.0)              # output normalize with a gain: peaks at `gain` instead of 1.0
    return p, a, s


@pytest.mark.parametrize("bd", [8, 24])
def test_too_wide_is_an_error_at_8_and_24_bit(gpulib, oracle, bd):
    """hound's write_sample refuses an i16/i32 that does not fit 8/24 bits (TooWide) and the reference unwraps
    it (state.rs:521,529): both libraries report TDG_ERR_TOO_WIDE instead of aborting."""
    for lib in (gpulib, oracle):
        with pytest.raises(AssertionError if False else TermdawError) as e:
            run_project(lib, *_loud_project(bd, 1.5), taps=False)
        assert e.value.code == -8, e.value
    got, want = run_project(gpulib, *_loud_project(bd, 0.999), taps=False), run_project(oracle, *_loud_project(bd, 0.999), taps=False)
    assert np.array_equal(got["pcm"], want["pcm"])


@pytest.mark.parametrize("bd", [16, 32])
def test_saturation_at_16_and_32_bit(gpulib, oracle, bd):
    """Rust `as i16` / `as i32` saturate (state.rs:521,529): a 1.5x too loud render clips, byte for byte."""
    got, want = run_project(gpulib, *_loud_project(bd, 1.5), taps=False), run_project(oracle, *_loud_project(bd, 1.5), taps=False)
    assert np.array_equal(got["pcm"], want["pcm"])
    v = pcm_to_int(got["pcm"], bd)
    lim = (1 << (bd - 1)) - 1
    assert int(v.max()) == lim and int(v.min()) == -lim - 1          # both rails are hit


@pytest.mark.parametrize("bd", [8, 16, 24, 32])
def test_nan_quantises_to_zero(gpulib, oracle, bd):
    """A silent sample normalises to NaN (0 * 1/0, sample.rs:125-130); NaN never wins absmax (sample.rs:12-14), so
    the output stays NaN through normalize and `NaN as i16|i32` is 0 (state.rs:521,529; 8 bit: offset 128)."""
    a = Assets()
    a.wavs["z"] = DecodedWav(np.zeros(2 * 3000, np.float32), 2, PSR, 16)
    a.wavs["k"] = S.sine_hit(6000, PSR, 80.0, 0.1)
    p = Project()
    p.set_length(0.3); p.set_render_samplerate(PSR); p.set_render_bitdepth(bd)
    p.load_sample("z", "z", ""); p.load_sample("k", "k", "")
    p.add_sampleloop("nan", 1.0, 0.0, "z")
    p.add_sampleloop("ok", 1.0, 0.0, "k")
    p.add_normalize("out", 1.0, 0.0)
    p.connect("nan", "out"); p.connect("ok", "out")
    p.set_output("out")
    cfg = (p, a, {"psr": PSR, "bl": BL, "flow": "B"})
    got, want = run_project(gpulib, *cfg), run_project(oracle, *cfg)
    assert np.isnan(want["taps"]["out"][0]).all() and np.isnan(got["taps"]["out"][0]).all()
    assert np.array_equal(got["pcm"], want["pcm"])
    assert np.all(pcm_to_int(got["pcm"], bd) == 0)


# ------------------------------------------------------------------------------------ (e) Vertex::new clamps
def test_angle_and_wet_clamps(gpulib, oracle):
    """angle is clamped to +-90 and wet to 0..1 when the vertex is built (graph.rs:255-256): a project asking for
    angle 120 / -300 and wet 1.5 / -0.5 renders exactly like the clamped project, on the device and the oracle."""
    def build(angle_a, angle_b, wet_a, wet_b):
        a = Assets()
        a.wavs["n"] = S.noise_burst(5, 9000, PSR, 0.2)
        a.flowws["h"] = S.hit_stream(6, 1.0, 6.0)
        p = Project()
        p.set_length(1.0); p.set_render_samplerate(PSR); p.set_render_bitdepth(16)
        p.load_sample("n", "n", ""); p.load_midi_floww("h", "h")
        p.add_sampleloop("src", 1.0, angle_a, "n")
        p.add_sum("s", 0.9, angle_b)
        p.add_adsr("e", 1.0, 0.0, wet_a, "h", False, True, -1, [0.01, 0.05, 0.5, 0.05, 0.2, 0.05])
        p.add_bandpass("b", 1.0, 0.0, wet_b, 900.0, 60.0, True)
        p.add_normalize("out", 1.0, 0.0)
        p.connect("src", "s"); p.connect("s", "e"); p.connect("e", "b"); p.connect("b", "out")
        p.set_output("out")
        return p, a, {"psr": PSR, "bl": BL, "flow": "B"}
    wild, tame = build(120.0, -300.0, 1.5, -0.5), build(90.0, -90.0, 1.0, 0.0)
    gw, gt = run_project(gpulib, *wild), run_project(gpulib, *tame)
    ow = run_project(oracle, *wild)
    for name in ow["taps"]:
        for ch in (0, 1):
            assert np.array_equal(gw["taps"][name][ch], gt["taps"][name][ch]), name
            assert np.array_equal(gw["taps"][name][ch], ow["taps"][name][ch]), name      # wet -0.5 -> 0: the filter is bypassed
    assert np.array_equal(gw["pcm"], ow["pcm"])
    assert float(np.abs(ow["taps"]["src"][1]).max()) < 1e-6                               # angle 90: cos 45 - sin 45 in f32, the right channel is (nearly) silent


# ------------------------------------------------------------------------------------ (f) late timeline / polyphony
def _late_project(seconds):
    """One synth, one sampsyn and one debug_sine, at most 4 voices each, over the whole timeline: every note is
    a few seconds long, notes keep coming until the end, so the oscillator argument time*hz*2pi and the f32 frame
    counter are exercised far beyond 2^24 frames (SURVEY hard parts 1 and 2)."""
    rng = np.random.default_rng(3600)
    ev, t = [], 0.5
    while t < seconds - 1.0:
        for k in range(int(rng.integers(1, 4))):
            note = float(rng.integers(40, 90)) + (0.25 * k)
            dur = float(rng.uniform(0.5, 6.0))
            ev.append((t + 0.01 * k, note, float(rng.uniform(0.3, 1.0))))
            ev.append((min(seconds - 0.25, t + dur), note, 0.0))
        t += float(rng.uniform(4.0, 9.0))
    a = Assets()
    a.flowws["n"] = _events(ev)
    a.blobs["tb"] = S.wavetable_blob(11, 16, 256, 6.0)
    p = Project()
    p.set_length(seconds); p.set_render_samplerate(PSR); p.set_render_bitdepth(16)
    p.load_midi_floww("n", "n"); p.load_resource("tb", "tb")
    p.add_synth("syn", 0.8, -20.0, "n", 0.4, 0.3, S.HIT_ADSR, 1.0, 0.8, S.NOTE_ADSR, 0.5, S.NOTE_ADSR)
    p.add_sampsyn("wt", 0.9, 15.0, "n", S.STD_ADSR, "tb")
    p.add_debug_sine("ds", 0.3, 0.0, "n")
    p.add_normalize("out", 1.0, 0.0)
    for v in ("syn", "wt", "ds"):
        p.connect(v, "out")
    p.set_output("out")
    return p, a, {"psr": PSR, "bl": BL, "flow": "B"}


def _windowed_compare(lib_a, lib_b, cfg, windows, win_frames):
    """scan + render on both libraries; f32 taps compared in `windows` (frame offsets), PCM compared whole.
    Returns (worst relative tap error, vertex, absolute frame index, worst PCM LSB, its frame index)."""
    p, a, s = cfg
    sts = []
    for lib in (lib_a, lib_b):
        st = State(lib, s["psr"], s["bl"], assets=a)
        assert st.refresh(p), st.messages
        st.ctx.enable_taps(True)
        st.scan_exact()
        sts.append((st, st.render()))
    (sa, pa), (sb, pb) = sts
    n = sa.cs * sa.bl
    worst = (0.0, None, -1)
    for f0 in windows:
        f0 = min(max(0, f0), max(0, n - win_frames))
        m = min(win_frames, n - f0)
        for i, name in enumerate(sa.names):
            ga, gb = sa.ctx.read_vertex_buffer(i, f0, m), sb.ctx.read_vertex_buffer(i, f0, m)
            for x, y in zip(ga, gb):
                peak = max(1.0, float(np.abs(y).max()))
                d = np.abs(x.astype(np.float64) - y.astype(np.float64))
                k = int(d.argmax())
                if d[k] / peak > worst[0]:
                    worst = (float(d[k] / peak), name, f0 + k)
    x, y = pcm_to_int(pa, sa.bd), pcm_to_int(pb, sb.bd)
    assert x.shape == y.shape
    dd = np.abs(x - y)
    k = int(dd.argmax())
    for st, _ in sts:
        st.close()
    return worst, int(dd[k]), k // 2


def test_late_timeline_full_hour_against_the_oracle(gpulib, oracle):
    """synth + sampsyn + debug_sine over the full 3600 s (172.8 M frames) vs the oracle: taps in windows around
    2^24 frames (349 s), the half hour and the last second; the 16-bit PCM over the whole hour."""
    seconds = 12.0 if EMULATED else 3600.0
    cfg = _late_project(seconds)
    n = int(seconds * PSR)
    windows = [0, (1 << 24) - (1 << 18), 1 << 25, n // 2, (1 << 27), n - (1 << 19)]
    worst, lsb, at = _windowed_compare(gpulib, oracle, cfg, windows, 1 << 19)
    out = {"seconds": seconds, "worst_rel_tap": worst[0], "worst_vertex": worst[1], "worst_frame": worst[2],
           "worst_pcm_lsb_16bit": lsb, "worst_pcm_frame": at}
    os.makedirs("gpurun_out", exist_ok=True)
    import json
    with open(os.path.join("gpurun_out", "late_timeline_report.json"), "w") as f:
        json.dump(out, f, indent=1)
    assert worst[0] <= TOL_F32, out
    assert lsb <= 1, out


def test_config3_all_512_voices_against_the_oracle(gpulib, oracle):
    """configs[2] with all 512 wavetable voices live, a 20 s slice (the first re-strike falls inside it)."""
    seconds = 1.5 if EMULATED else 21.0
    cfg = S.config_3(seconds / 600.0, voices=64 if EMULATED else 512)
    n = int(seconds * PSR)
    worst, lsb, at = _windowed_compare(gpulib, oracle, cfg, [0, n // 2, n - (1 << 18)], 1 << 18)
    assert worst[0] <= TOL_F32, (worst, lsb, at)
    assert lsb <= 1, (worst, lsb, at)


# ------------------------------------------------------------------------------------ refresh: bank diffs on the device
def test_refresh_drops_and_reloads_banks_on_the_device(gpulib):
    """State::refresh's sample / resource diff (state.rs:172-237, sample.rs:316-336) against the device banks:
    removed samples are freed, a changed resource path is reloaded (tests/test_host_frontend.py holds the body)."""
    import test_host_frontend as H
    H.test_refresh_diffs_samples_and_resources_like_the_reference(gpulib)


# ------------------------------------------------------------------------------------ output stage: files from pinned staging
def test_batch_render_to_files_overlapped_writer(gpulib, oracle, tmp_path):
    """termdaw_b200.batch.render_projects_to_files: groups of projects rendered by the batched call into pinned staging
    buffers (tdg_host_alloc) while a thread pool writes the previous group's WAV files.  Every file equals, byte for
    byte, the file the one-project-at-a-time path writes (State.render_to_file), and its PCM the oracle's within 1 LSB."""
    from termdaw_b200.batch import render_projects_to_files
    from termdaw_b200.wav import read_wav
    cfgs = [S.config_5_project(i, 0.03) for i in range(9)]
    states = []
    for p, a, s in cfgs:
        st = State(gpulib, s["psr"], s["bl"], assets=a)
        assert st.refresh(p), st.messages
        states.append(st)
    paths = [str(tmp_path / f"batch_{i}.wav") for i in range(len(states))]
    render_projects_to_files(gpulib, states, paths, group=4, writers=3)
    for st in states:
        st.close()
    for i, (p, a, s) in enumerate(cfgs):
        st = State(gpulib, s["psr"], s["bl"], assets=a)
        assert st.refresh(p), st.messages
        st.scan_exact()
        single = str(tmp_path / f"single_{i}.wav")
        st.render_to_file(single)
        st.close()
        assert open(single, "rb").read() == open(paths[i], "rb").read(), i
        want = run_project(oracle, p, a, s, taps=False)
        got = read_wav(paths[i])
        x = np.asarray(got.linear, np.float64)
        y = pcm_to_int(want["pcm"], want["bd"]).astype(np.float64)
        assert x.shape == y.shape and np.abs(x - y).max() <= 1, i
