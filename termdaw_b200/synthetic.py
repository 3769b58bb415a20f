"""Seeded synthetic inputs and the five BASELINE.json configurations.

Everything the reference would read from disk is generated here: noise/sine sample WAV streams
(as hound would decode them: interleaved, integer PCM cast to f32 unscaled), floww note streams
(t seconds, note, vel; vel 0 = note-off) and wavetable blobs in the "WTB1" layout DESIGN.md defines.
Each `config_N` returns `(project, assets, settings)`; `scale` shrinks the duration for tests while
keeping the structure.  Nothing here touches the oracle or the device.
"""
from __future__ import annotations

import struct
from typing import Dict, Tuple

import numpy as np

from .state import Assets, DecodedWav, Project

PSR, BL = 48000, 1024
HIT_ADSR = [0.001, 0.02, 0.0, 0.0, 0.0, 0.0]            # project.lua:29
NOTE_ADSR = [0.01, 0.1, 0.8, 5.0, 0.2, 0.5]             # project.lua:30
STD_ADSR = [0.01, 1.0, 1.0, 1.0, 1.0, 0.4]              # project.lua:31


# ------------------------------------------------------------------------------------ raw inputs
def pcm16(x: np.ndarray) -> np.ndarray:
    """float [-1,1] -> what `reader.samples::<i32>()` + `as f32` yields for a 16-bit file."""
    return np.round(np.clip(x, -1, 1) * 32767.0).astype(np.int16).astype(np.float32)


def stereo_wav(l: np.ndarray, r: np.ndarray, sr: int) -> DecodedWav:
    lin = np.empty(l.size * 2, np.float32)
    lin[0::2], lin[1::2] = pcm16(l), pcm16(r)
    return DecodedWav(lin, 2, sr, 16)


def noise_burst(seed: int, frames: int, sr: int, tau: float) -> DecodedWav:
    rng = np.random.default_rng(seed)
    t = np.arange(frames) / sr
    env = np.exp(-t / tau)
    return stereo_wav(rng.uniform(-1, 1, frames) * env, rng.uniform(-1, 1, frames) * env, sr)


def sine_hit(frames: int, sr: int, hz: float, tau: float) -> DecodedWav:
    t = np.arange(frames) / sr
    x = np.sin(2 * np.pi * hz * t) * np.exp(-t / tau)
    return stereo_wav(x, x, sr)


def sine_sweep(frames: int, sr: int, f0: float, f1: float) -> DecodedWav:
    t = np.arange(frames) / sr
    ph = 2 * np.pi * (f0 * t + (f1 - f0) * t * t / (2 * t[-1] if frames > 1 else 1))
    x = np.sin(ph) * 0.9
    return stereo_wav(x, np.roll(x, 7), sr)


def wavetable_blob(seed: int, n_frames: int = 32, n_phase: int = 512, fps: float = 8.0) -> bytes:
    """Time-evolving table: a few decaying harmonics, brighter at the start."""
    rng = np.random.default_rng(seed)
    ph = np.arange(n_phase) / n_phase
    amps = rng.uniform(0.2, 1.0, 6) / np.arange(1, 7)
    data = np.zeros((n_frames, n_phase), np.float64)
    for j in range(n_frames):
        decay = np.exp(-j / n_frames * np.arange(6) * 0.8)
        for h in range(6):
            data[j] += amps[h] * decay[h] * np.sin(2 * np.pi * (h + 1) * ph)
    data /= np.abs(data).max()
    return b"WTB1" + struct.pack("<IIf", n_frames, n_phase, fps) + data.astype("<f4").tobytes()


def note_stream(seed: int, seconds: float, grid: float = 0.125, p_on: float = 0.7, lo: int = 36, hi: int = 84,
                dur=(0.1, 0.5), fractional: bool = False):
    """On/off note events on a grid (config 2's floww)."""
    rng = np.random.default_rng(seed)
    ev = []
    for s in range(int(seconds / grid)):
        if rng.random() < p_on:
            t0 = s * grid
            note = rng.uniform(lo, hi) if fractional else float(rng.integers(lo, hi + 1))
            vel = rng.uniform(0.3, 1.0)
            ev.append((t0, note, vel))
            ev.append((t0 + rng.uniform(*dur), note, 0.0))
    ev.sort(key=lambda e: e[0])
    a = np.array(ev, np.float32).reshape(-1, 3)
    return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy()


def hit_stream(seed: int, seconds: float, rate: float = 4.0, notes=(36, 38, 42)):
    """Drum-style hits (note-on only) at ~`rate` per second."""
    rng = np.random.default_rng(seed)
    n = max(1, int(seconds * rate))
    t = np.sort(rng.uniform(0, seconds, n)).astype(np.float32)
    note = rng.choice(np.array(notes, np.float32), n)
    vel = rng.uniform(0.3, 1.0, n).astype(np.float32)
    return t, note.astype(np.float32), vel


def held_voices(seed: int, seconds: float, voices: int, restrike: float = 20.0):
    """Config 3's floww: `voices` distinct fractional notes switched on within the first second,
    held, and re-struck (off then on) every `restrike` seconds."""
    rng = np.random.default_rng(seed)
    notes = rng.uniform(36, 96, voices)
    starts = np.sort(rng.uniform(0, min(1.0, seconds / 4), voices))
    ev = []
    for v in range(voices):
        t = starts[v]
        ev.append((t, notes[v], rng.uniform(0.3, 1.0)))
        k = 1
        while t + k * restrike < seconds:
            tt = t + k * restrike
            ev.append((tt - 0.01, notes[v], 0.0))
            ev.append((tt, notes[v], rng.uniform(0.3, 1.0)))
            k += 1
    ev.sort(key=lambda e: e[0])
    a = np.array(ev, np.float32).reshape(-1, 3)
    return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy()


# ------------------------------------------------------------------------------------ configs
def config_1(scale: float = 1.0):
    """README example (README.md:93-111): two sampleloops -> normalize, 3 s, 44.1 kHz 16-bit."""
    a = Assets()
    a.wavs["snare.wav"] = noise_burst(101, 12000, 48000, 0.05)
    a.wavs["kick.wav"] = sine_hit(19200, 48000, 60.0, 0.1)
    p = Project()
    p.set_length(3.0 * scale)
    p.set_render_samplerate(44100)
    p.set_render_bitdepth(16)
    p.set_output_file("outp.wav")
    p.load_sample("snare", "snare.wav", "")
    p.load_sample("kick", "kick.wav", "")
    p.add_sampleloop("one", 1.0, 0.0, "snare")
    p.add_sampleloop("two", 1.0, 0.0, "kick")
    p.add_normalize("sum", 1.0, 0.0)
    p.connect("one", "sum")
    p.connect("two", "sum")
    p.set_output("sum")
    return p, a, {"psr": PSR, "bl": BL, "flow": "B"}


def config_2(scale: float = 1.0):
    """floww-driven synth + ADSR -> band pass -> normalize, 60 s, 48 kHz, 24-bit."""
    seconds = 60.0 * scale
    a = Assets()
    a.flowws["notes.mid"] = note_stream(202, seconds)
    p = Project()
    p.set_length(seconds)
    p.set_render_samplerate(48000)
    p.set_render_bitdepth(24)
    p.load_midi_floww("notes", "notes.mid")
    p.add_synth("syn", 1.0, 0.0, "notes", 0.4, 0.3, HIT_ADSR, 1.0, 0.8, NOTE_ADSR, 0.5, NOTE_ADSR)
    p.add_adsr("env", 1.0, 0.0, 1.0, "notes", False, True, -1, [0.01, 0.1, 0.8, 0.1, 0.2, 0.01])
    p.add_bandpass("band", 1.0, 0.0, 1.0, 1000.0, 50.0, True)
    p.add_normalize("out", 1.0, 0.0)
    p.connect("syn", "env")
    p.connect("env", "band")
    p.connect("band", "out")
    p.set_output("out")
    return p, a, {"psr": PSR, "bl": BL, "flow": "B"}


def config_3(scale: float = 1.0, voices: int = 512):
    """Wavetable synth with `voices` simultaneous voices + 2 sample_lerp from 44.1/96 kHz sources."""
    seconds = 600.0 * scale
    a = Assets()
    a.blobs["table.wtb"] = wavetable_blob(303)
    a.flowws["pad.mid"] = held_voices(304, seconds, voices)
    a.flowws["h1.mid"] = hit_stream(305, seconds, 4.0)
    a.flowws["h2.mid"] = hit_stream(306, seconds, 4.0)
    a.wavs["burst44.wav"] = noise_burst(307, 2 * 44100, 44100, 0.4)
    a.wavs["sweep96.wav"] = sine_sweep(2 * 96000, 96000, 100.0, 4000.0)
    p = Project()
    p.set_length(seconds)
    p.set_render_samplerate(48000)
    p.set_render_bitdepth(16)
    p.load_sample("burst", "burst44.wav", "")
    p.load_sample("sweep", "sweep96.wav", "")
    p.load_resource("table", "table.wtb")
    p.load_midi_floww("pad", "pad.mid")
    p.load_midi_floww("h1", "h1.mid")
    p.load_midi_floww("h2", "h2.mid")
    p.add_sampsyn("pad", 1.0, 0.0, "pad", STD_ADSR, "table")
    p.add_sample_lerp("l1", 0.8, 30.0, "burst", "h1", -1, 40)
    p.add_sample_lerp("l2", 0.8, -30.0, "sweep", "h2", -1, 40)
    p.add_normalize("out", 1.0, 0.0)
    for v in ("pad", "l1", "l2"):
        p.connect(v, "out")
    p.set_output("out")
    return p, a, {"psr": PSR, "bl": BL, "flow": "B"}


_MIX_KINDS = ["sample_loop"] * 16 + ["sample_multi"] * 8 + ["sample_lerp"] * 8 + ["synth"] * 8 + \
             ["sampsyn"] * 4 + ["adsr"] * 6 + ["band_pass"] * 6 + ["sum"] * 7


def _shared_assets(rng: np.random.Generator, seconds: float, n_flowws: int = 4) -> Assets:
    a = Assets()
    a.wavs["s_noise"] = noise_burst(int(rng.integers(1 << 30)), 9000, 48000, 0.08)
    a.wavs["s_tone"] = sine_hit(14000, 48000, float(rng.uniform(50, 200)), 0.12)
    a.wavs["s_long"] = noise_burst(int(rng.integers(1 << 30)), 30000, 48000, 0.5)
    a.blobs["table"] = wavetable_blob(int(rng.integers(1 << 30)), 16, 256, 6.0)
    for i in range(n_flowws):
        a.flowws[f"hits{i}"] = hit_stream(int(rng.integers(1 << 30)), seconds, float(rng.uniform(2, 6)))
        a.flowws[f"notes{i}"] = note_stream(int(rng.integers(1 << 30)), seconds, grid=0.25, p_on=0.6)
    return a


def random_project(seed: int, kinds, seconds: float, bitdepth: int = 16, n_flowws: int = 4):
    """Tree-shaped random DAG (every vertex reaches the output, fan-in <= 8) over `kinds`
    (a list of generator/effect kind names) feeding one output normalize."""
    rng = np.random.default_rng(seed)
    a = _shared_assets(rng, seconds, n_flowws)
    p = Project()
    p.set_length(seconds)
    p.set_render_samplerate(48000)
    p.set_render_bitdepth(bitdepth)
    for s in ("s_noise", "s_tone", "s_long"):
        p.load_sample(s, s, "")
    p.load_resource("table", "table")
    for i in range(n_flowws):
        p.load_midi_floww(f"hits{i}", f"hits{i}")
        p.load_midi_floww(f"notes{i}", f"notes{i}")
    gens, fx = [], []
    for i, k in enumerate(kinds):
        name = f"v{i}_{k}"
        gain = float(rng.uniform(0.2, 1.5))
        angle = float(rng.uniform(-90, 90))
        smp = ("s_noise", "s_tone", "s_long")[int(rng.integers(3))]
        hits = f"hits{int(rng.integers(n_flowws))}"
        notes = f"notes{int(rng.integers(n_flowws))}"
        if k == "sample_loop":
            p.add_sampleloop(name, gain, angle, smp); gens.append(name)
        elif k == "sample_multi":
            p.add_sample_multi(name, gain, angle, smp, hits, int(rng.choice([-1, 36, 38]))); gens.append(name)
        elif k == "sample_lerp":
            p.add_sample_lerp(name, gain, angle, smp, hits, int(rng.choice([-1, 36])), int(rng.choice([0, 40, 200])))
            gens.append(name)
        elif k == "debug_sine":
            p.add_debug_sine(name, gain, angle, notes); gens.append(name)
        elif k == "synth":
            p.add_synth(name, gain, angle, notes, float(rng.choice([0.0, 0.4])), 0.3, HIT_ADSR, 1.0,
                        float(rng.uniform(0.3, 0.9)), NOTE_ADSR, float(rng.choice([0.0, 0.5])), NOTE_ADSR)
            gens.append(name)
        elif k == "sampsyn":
            p.add_sampsyn(name, gain, angle, notes, STD_ADSR, "table"); gens.append(name)
        elif k == "adsr":
            use_off = bool(rng.integers(2))
            p.add_adsr(name, gain, angle, float(rng.choice([1.0, 0.7])), notes if use_off else hits, use_off,
                       bool(rng.integers(2)), -1, [0.01, 0.1, 0.8, 0.1, 0.2, 0.05]); fx.append(name)
        elif k == "band_pass":
            p.add_bandpass(name, gain, angle, 1.0, float(rng.choice([0.0, 1000.0, 4000.0])),
                           float(rng.choice([50.0, 200.0, 0.0])), bool(rng.integers(2))); fx.append(name)
        elif k == "sum":
            p.add_sum(name, gain, angle); fx.append(name)
        else:
            raise ValueError(k)
    p.add_normalize("out", 1.0, 0.0)
    targets = fx + ["out"]
    fan = {t: 0 for t in targets}

    def pick(lo: int) -> str:
        cand = [t for t in targets[lo:] if fan[t] < 8] or ["out"]
        t = cand[int(rng.integers(len(cand)))]
        fan[t] += 1
        return t

    for j, f in enumerate(fx):                      # effects chain forward => acyclic, reaches "out"
        p.connect(f, pick(j + 1))
    for g in gens:
        p.connect(g, pick(0))
    for f in fx:                                    # an effect without inputs would be silent; feed it
        if not any(b == f for (_, b) in p.edges) and gens:
            p.connect(gens[int(rng.integers(len(gens)))], f)
    p.set_output("out")
    return p, a, {"psr": PSR, "bl": BL, "flow": "B"}


def config_4(scale: float = 1.0, n_vertices: int = 64):
    """Single long mixed graph: 63 vertices of mixed kinds + the output normalize, 3600 s, 24-bit."""
    rng = np.random.default_rng(404)
    kinds = list(_MIX_KINDS)
    rng.shuffle(kinds)
    kinds = kinds[: max(1, n_vertices - 1)]
    return random_project(404, kinds, 3600.0 * scale, bitdepth=24, n_flowws=6)


_BATCH_GEN = ["sample_loop", "sample_multi", "sample_lerp", "synth", "sampsyn"]
_BATCH_FX = ["adsr", "band_pass", "sum"]


def config_5_project(index: int, scale: float = 1.0):
    """One project of the algorithmic-composition batch: seed 5000+index, 4-16 vertices, 30 s, 16-bit."""
    seed = 5000 + index
    rng = np.random.default_rng(seed)
    n = int(rng.integers(4, 17)) - 1
    n_fx = max(1, n // 3)
    kinds = [_BATCH_GEN[int(rng.integers(len(_BATCH_GEN)))] for _ in range(n - n_fx)] + \
            [_BATCH_FX[int(rng.integers(len(_BATCH_FX)))] for _ in range(n_fx)]
    return random_project(seed, kinds, 30.0 * scale, bitdepth=16, n_flowws=2)


CONFIGS = {1: config_1, 2: config_2, 3: config_3, 4: config_4}


# ------------------------------------------------------------------------------------ projects on disk
def _lua(v) -> str:
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, str):
        return '"' + v.replace("\\", "\\\\").replace('"', '\\"') + '"'
    if isinstance(v, (list, tuple)):
        return "{" + ", ".join(_lua(x) for x in v) + "}"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    return repr(float(v))


def project_to_lua(p: Project) -> str:
    """The script text that produces `p`'s seed vectors (registered globals of state.rs:103-157)."""
    out = []

    def call(fn, *args):
        out.append(f"{fn}({', '.join(_lua(a) for a in args)})")

    if p.length_seconds is not None: call("set_length", p.length_seconds)
    if p.render_samplerate is not None: call("set_render_samplerate", p.render_samplerate)
    if p.render_bitdepth is not None: call("set_render_bitdepth", p.render_bitdepth)
    if p.output_file is not None: call("set_output_file", p.output_file)
    for t in p.samples: call("load_sample", *t)
    for t in p.resources: call("load_resource", *t)
    for t in p.midis: call("load_midi_floww", *t)
    for t in p.streams: call("declare_stream", t)
    for fn, lst in (("add_sum", p.sums), ("add_normalize", p.norms), ("add_sampleloop", p.sampleloops),
                    ("add_sample_multi", p.samplemultis), ("add_sample_lerp", p.samplelerps),
                    ("add_debug_sine", p.debugsines), ("add_synth", p.synths), ("add_sampsyn", p.sampsyns),
                    ("add_lv2fx", p.lv2fxs), ("add_adsr", p.adsrs), ("add_bandpass", p.bandpasses)):
        for t in lst:
            call(fn, *t)
    for a, b in p.edges: call("connect", a, b)
    call("set_output", p.output_vertex)
    return "\n".join(out) + "\n"


def write_project_dir(cfg, path: str, main: str = "project.lua") -> str:
    """Materialise a synthetic (project, assets, settings) triple as a termdaw project directory:
    project.toml, the Lua script, 16-bit WAV files for the samples, wavetable blobs, and every floww as a
    `<name>.floww.f32` side file (little-endian f32 triples t, note, vel) that both command-line
    front-ends (`python -m termdaw_b200.cli`, `termdaw_b200_render`) read in place of a MIDI file."""
    import os
    from .wav import encode_wav
    p, a, s = cfg
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "project.toml"), "w") as f:
        f.write(f'[project]\nname = "synthetic"\n\n[settings]\nmain = "{main}"\nbuffer_length = {s["bl"]}\n'
                f'project_samplerate = {s["psr"]}\n')
    with open(os.path.join(path, main), "w") as f:
        f.write(project_to_lua(p))
    for name, w in a.wavs.items():
        if w.bits != 16:
            raise ValueError("write_project_dir writes 16-bit samples only")
        pcm = np.asarray(w.linear, np.float32).astype("<i2").view(np.uint8)
        with open(os.path.join(path, name), "wb") as f:
            f.write(encode_wav(pcm, w.sample_rate, 16, w.channels))
    for name, b in a.blobs.items():
        with open(os.path.join(path, name), "wb") as f:
            f.write(b)
    for name, (t, n, v) in a.flowws.items():
        tri = np.stack([np.asarray(t, np.float32), np.asarray(n, np.float32), np.asarray(v, np.float32)], axis=1)
        with open(os.path.join(path, name + ".floww.f32"), "wb") as f:
            f.write(np.ascontiguousarray(tri, "<f4").tobytes())
    return path
