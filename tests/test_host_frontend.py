"""Host-side logic on CPU: the Lua-subset reader, the WAV container, `State.refresh` ordering rules.
The oracle library stands in as the backend here only because there is no GPU in this container; the
same `State` code drives the CUDA library in the -m gpu tests."""
import os

import numpy as np
import pytest

from termdaw_b200 import lua_subset, wav
from termdaw_b200.state import Assets, DecodedWav, Project, State, set_length_chunks
from termdaw_b200 import synthetic as S

REF_LIKE_SCRIPT = """
-- Our lua file, where we do all the work
set_length(2.0);
set_render_samplerate(48000);
set_render_bitdepth(16);
set_output_file("outp.wav");
load_sample("background", "bg.wav", "");
load_sample("kick", "kick.wav", "");
load_midi_floww("bassd", "bassd.midi");
add_sample_lerp("kick", 1.0, 0.0, "kick", "bassd", -1, 40);
add_sampleloop("background", 1.0, 0.0, "background");
dip = 0.3;
adsr = { 1.0, 0.01, dip, 0.2, dip, 0.0, 0.0, 0.05, 1.0 };
add_adsr("env", 1.0, 0.0, 1.0, "bassd", false, false, -1, adsr);
add_normalize("sum", 1.0, 0.0);
connect("kick", "sum");
connect("background", "env");
connect("env", "sum");
set_output("sum");
"""


def test_lua_subset_reads_the_reference_example_shape():
    """Same statements as examples/neg-adsr-env-example.lua (paths shortened)."""
    p = lua_subset.run_script(REF_LIKE_SCRIPT)
    assert p.length_seconds == 2.0 and p.render_samplerate == 48000 and p.render_bitdepth == 16
    assert p.samples == [("background", "bg.wav", ""), ("kick", "kick.wav", "")]
    assert p.samplelerps == [("kick", 1.0, 0.0, "kick", "bassd", -1, 40)]
    (name, gain, angle, wet, floww, use_off, use_max, note, conf), = p.adsrs
    assert (name, use_off, use_max, note) == ("env", False, False, -1)
    assert conf == [1.0, 0.01, 0.3, 0.2, 0.3, 0.0, 0.0, 0.05, 1.0]
    assert p.edges == [("kick", "sum"), ("background", "env"), ("env", "sum")]
    assert p.output_vertex == "sum"


def test_lua_subset_real_names_aliases_loops_and_comments():
    p = lua_subset.run_script("""
        --[[ block
             comment ]]
        local base = 36
        for i = 0, 2 do
            add_sample_multi("m" .. i, 0.5 * (i + 1), -10, "s", "f", base + 2 * i)   -- no semicolon
        end
        add_samplemulti("alias", 1, 0, 's', 'f', -1);      -- BASELINE.json paraphrase
        add_synth("syn", 1.0, 0.0, "f", 0.4, 0.3, {}, 1.0, 0.8, { 0.01, 0.1, 0.8, 5.0, 0.2, 0.5 }, 0.0, {});
        add_bandpass("band", 1.0, 0.0, 1.0, 1000, 0, true)
    """)
    assert [m[0] for m in p.samplemultis] == ["m0", "m1", "m2", "alias"]
    assert [m[1] for m in p.samplemultis] == [0.5, 1.0, 1.5, 1]
    assert [m[5] for m in p.samplemultis] == [36, 38, 40, -1]
    assert p.synths[0][6] == [] and p.synths[0][9] == [0.01, 0.1, 0.8, 5.0, 0.2, 0.5]
    assert p.bandpasses == [("band", 1.0, 0.0, 1.0, 1000.0, 0.0, True)]


@pytest.mark.parametrize("bad", ["function f() end", "x = y.z", "add_sum('a', 1)", "undefined_fn(1)", "if x then end"])
def test_lua_subset_rejects_what_it_does_not_support(bad):
    with pytest.raises(lua_subset.LuaSubsetError):
        lua_subset.run_script(bad)


@pytest.mark.parametrize("ref_file", ["project.lua", "examples/neg-adsr-env-example.lua", "examples/sample-project.lua",
                                       "examples/sample-synth-adsr-lv2fx-example.lua", "examples/stream.lua"])
def test_lua_subset_parses_the_reference_scripts(ref_file):
    """Only where the reference checkout is present (this container); skipped on the GPU box."""
    path = os.path.join("/root/reference", ref_file)
    if not os.path.exists(path):
        pytest.skip("reference checkout not present")
    p = lua_subset.run_file(path)
    assert p.output_vertex != ""
    assert p.edges


@pytest.mark.parametrize("bd", [8, 16, 24, 32])
def test_wav_roundtrip_integer_pcm(bd, tmp_path):
    rng = np.random.default_rng(bd)
    l, r = rng.uniform(-1, 1, 777), rng.uniform(-1, 1, 777)
    pcm = wav.pcm_from_float(l, r, bd)
    path = str(tmp_path / f"x{bd}.wav")
    wav.write_wav(path, pcm, 44100, bd)
    blob = open(path, "rb").read()
    assert blob[:4] == b"RIFF" and blob[8:12] == b"WAVE"
    assert len(blob) == (44 if bd <= 16 else 68) + pcm.size          # PCM vs WAVE_FORMAT_EXTENSIBLE header
    w = wav.read_wav(path)
    assert (w.channels, w.sample_rate, w.bits) == (2, 44100, bd)
    amp = {8: 127.0, 16: 32767.0, 24: 8388607.0, 32: 2147483647.0}[bd]
    want = np.empty(1554); want[0::2], want[1::2] = l, r
    # integers come back UNSCALED (sample.rs:266-273 casts `as f32`)
    assert np.abs(w.linear / amp - want).max() <= 1.0 / amp + 1e-7
    assert float(np.abs(w.linear).max()) > 1.0


def test_wav_float_and_errors(tmp_path):
    import struct
    x = np.linspace(-1, 1, 64, dtype="<f4")
    fmt = struct.pack("<HHIIHH", 3, 1, 48000, 48000 * 4, 4, 32)
    blob = b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt) + 8 + x.nbytes) + b"WAVE" + b"fmt " + \
        struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", x.nbytes) + x.tobytes()
    w = wav.decode_wav(blob)
    assert w.channels == 1 and np.array_equal(w.linear, x)
    with pytest.raises(wav.WavError):
        wav.decode_wav(b"not a wav file at all")


def test_set_length_chunks_is_f32(oracle):
    """cs = ceil(psr as f32 * seconds / bl as f32) (state.rs:104) in f32, against the oracle's C restatement."""
    for psr, sec, bl in [(48000, 3.0, 1024), (48000, 60.0, 1024), (44100, 0.1, 256), (48000, 3600.0, 1024),
                         (48000, 21.333334, 1000), (96000, 1e-3, 64)]:
        assert set_length_chunks(psr, sec, bl) == oracle.set_length_chunks(psr, sec, bl)
    assert set_length_chunks(48000, 3.0, 1024) == 141 and set_length_chunks(48000, 3600.0, 1024) == 168750


def _tiny_assets():
    a = Assets()
    a.wavs["s.wav"] = S.noise_burst(1, 600, 48000, 0.01)
    a.flowws["f.mid"] = S.hit_stream(2, 0.5, 8.0)
    return a


def test_state_refresh_kind_order_and_edge_rules(oracle):
    """Vertices are indexed by KIND in the fixed order of state.rs:341-457, not in script order; connect
    refuses self edges, cycles and generator targets silently (graph.rs:58-78); unknown names warn."""
    p = Project()
    p.set_length(0.1)
    p.load_sample("s", "s.wav", "")
    p.load_midi_floww("f", "f.mid")
    p.add_bandpass("bp", 1, 0, 1, 1000, 50, True)
    p.add_adsr("env", 1, 0, 1, "f", False, True, -1, [0.01, 0.1, 0.8, 0.1, 0.2, 0.01])
    p.add_sample_multi("multi", 1, 0, "s", "f", -1)
    p.add_sampleloop("loop", 1, 0, "s")
    p.add_normalize("out", 1, 0)
    p.add_sum("mix", 1, 0)
    for e in [("loop", "mix"), ("multi", "mix"), ("mix", "env"), ("env", "bp"), ("bp", "out"),
              ("bp", "mix"),          # would close a cycle: refused
              ("mix", "loop"),        # generator takes no input: refused
              ("out", "out"),         # self edge: refused
              ("ghost", "out")]:      # unknown name: warning
        p.connect(*e)
    p.set_output("out")
    st = State(oracle, 48000, 256, assets=_tiny_assets())
    assert st.refresh(p), st.messages
    assert st.names == ["mix", "out", "loop", "multi", "env", "bp"]
    assert any("ghost" in m for m in st.messages)
    ok, unreachable = st.ctx.check_graph()
    assert ok and unreachable == []
    pcm = st.render()
    assert pcm.size == st.cs * st.bl * 2 * 2
    st.close()


def test_state_refresh_failure_modes(oracle):
    p = Project()
    p.load_sample("s", "missing.wav", "")
    p.add_sampleloop("loop", 1, 0, "s")
    p.add_normalize("out", 1, 0)
    p.connect("loop", "out")
    p.set_output("out")
    st = State(oracle, 48000, 256, assets=Assets())
    assert not st.refresh(p) and not st.loaded
    st.close()
    # output that receives no inputs fails the graph check (graph.rs:150-160)
    p2 = Project()
    p2.add_normalize("out", 1, 0)
    p2.set_output("out")
    st2 = State(oracle, 48000, 256, assets=Assets())
    assert not st2.refresh(p2)
    st2.close()
    # mono file with a stereo load mode errors out (sample.rs:64-69)
    a = Assets()
    a.wavs["m.wav"] = DecodedWav(np.ones(100, np.float32), 1, 48000, 16)
    p3 = Project()
    p3.load_sample("m", "m.wav", "")
    st3 = State(oracle, 48000, 256, assets=a)
    assert not st3.refresh(p3)
    st3.close()


def test_refresh_diffs_samples_and_resources_like_the_reference(oracle):
    """State::refresh diffs (name, path, mode) / (name, path) tuples against the previous refresh (state.rs:172-237):
    what left the script is dropped from the bank (SampleBank::mark_dead + refresh, sample.rs:316-336), a changed
    path under the same name is RELOADED, unchanged entries are kept.  The render after an edited refresh equals
    a fresh State's render of the edited project."""
    from helpers import run_project
    from termdaw_b200 import synthetic as S
    from termdaw_b200._capi import TermdawError

    def project(table, with_kick):
        p, a, s = S.config_3(0.004, voices=8)
        a.blobs["other.wtb"] = S.wavetable_blob(999, 8, 128, 3.0)
        p.resources[0] = ("table", table)
        if not with_kick:
            p.samples = [t for t in p.samples if t[0] != "sweep"]
            p.samplelerps = [t for t in p.samplelerps if t[3] != "sweep"]
            p.edges = [e for e in p.edges if e[0] != "l2"]
        return p, a, s

    p1, a, s = project("table.wtb", True)
    st = State(oracle, s["psr"], s["bl"], assets=a)
    assert st.refresh(p1), st.messages
    sweep_idx = st.sample_index["sweep"]
    assert st.ctx.sample_frames(sweep_idx) > 0
    p2, _, _ = project("other.wtb", False)
    st.messages.clear()
    assert st.refresh(p2), st.messages
    assert any("sample \"sweep\" will be removed" in m for m in st.messages)
    assert any("resource \"table\" will be removed" in m for m in st.messages)
    assert "sweep" not in st.sample_index and st.ctx.sample_frames(sweep_idx) == 0       # the slot is empty, memory returned
    with pytest.raises(TermdawError):
        st.ctx.add_sample_loop(1.0, 0.0, sweep_idx)                                      # a removed sample cannot be named
    st.scan_exact()
    edited = st.render().copy()
    fresh = run_project(oracle, p2, a, s, taps=False)["pcm"]
    assert np.array_equal(edited, fresh)                                                 # the NEW table is what sounds
    unchanged = run_project(oracle, *project("table.wtb", False), taps=False)["pcm"]
    assert not np.array_equal(edited, unchanged)
    st.messages.clear()
    assert st.refresh(p1), st.messages                                                   # and back: sweep is loaded again
    assert st.sample_index["sweep"] != sweep_idx and st.ctx.sample_frames(st.sample_index["sweep"]) > 0
    st.close()
