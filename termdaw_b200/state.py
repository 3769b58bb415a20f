"""Host-side mirror of the reference's project front-end for the render path.

`Project` is what a termdaw Lua script sees: the registered globals of `State::refresh`
(state.rs:103-157) under their real names, each one only appending a tuple to a seed vector
(state.rs:85-93).  `State` is the host half of `State` (state.rs:27-47): `refresh` rebuilds banks and
graph through the C ABI in the reference's FIXED kind order (state.rs:341-457), `scan_exact`
(state.rs:473) and `render` (state.rs:477) drive the device.  The Lua interpreter itself is not
re-implemented here (see termdaw_b200/lua_subset.py for the literal-call subset the examples use).

Inputs that the reference reads from disk (WAV samples, MIDI flowws, wavetable blobs) are resolved
through `Assets`, which holds already-decoded arrays keyed by the path string the script passes, so
the same script text works for synthetic inputs; real WAV files are decoded by termdaw_b200/wav.py.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from ._capi import Context, Library, TermdawError

SAMPLE_MODES = {"left": 1, "right": 2, "loudest": 3, "normalize-seperate": 4, "mix-down": 5}


def sample_mode_from_string(s: str) -> int:
    """SampleLoadMethod::from (sample.rs:202-211): anything unknown is Stereo."""
    return SAMPLE_MODES.get(s, 0)


def set_length_chunks(psr: int, seconds: float, bl: int) -> int:
    """`cs = (psr as f32 * seconds / bl as f32).ceil() as usize` (state.rs:104), in f32."""
    v = np.float32(psr) * np.float32(seconds) / np.float32(bl)
    return int(max(0.0, math.ceil(float(v))))


@dataclass
class DecodedWav:
    """What hound hands SampleBank::add: the interleaved sample stream, integer PCM cast `as f32`
    unscaled (sample.rs:266-273)."""
    linear: np.ndarray
    channels: int
    sample_rate: int
    bits: int = 16


@dataclass
class Assets:
    wavs: Dict[str, DecodedWav] = field(default_factory=dict)
    flowws: Dict[str, Tuple[np.ndarray, np.ndarray, np.ndarray]] = field(default_factory=dict)
    blobs: Dict[str, bytes] = field(default_factory=dict)

    def wav(self, path: str) -> DecodedWav:
        if path in self.wavs:
            return self.wavs[path]
        if os.path.exists(path):
            from .wav import read_wav
            return read_wav(path)
        raise FileNotFoundError(f"TermDaw: SampleBank: could not open file \"{path}\".")

    def floww(self, path: str):
        if path in self.flowws:
            return self.flowws[path]
        if os.path.exists(path):                        # FlowwBank::add_floww -> read_floww_from_midi (floww.rs:40-48)
            from .midi import MidiError, read_floww_from_midi
            try:
                return read_floww_from_midi(path)
            except (MidiError, OSError):
                pass
        raise FileNotFoundError(f"Could not read midi file: \"{path}\".")                  # floww.rs:45-46

    def blob(self, path: str) -> bytes:
        if path in self.blobs:
            return self.blobs[path]
        with open(path, "rb") as f:
            return f.read()


class Project:
    """The Lua-visible API.  Method names and argument order are the registered globals."""

    def __init__(self):
        self.length_seconds: Optional[float] = None
        self.render_samplerate: Optional[int] = None
        self.render_bitdepth: Optional[int] = None
        self.output_file: Optional[str] = None
        self.output_vertex: str = ""
        (self.samples, self.resources, self.midis, self.streams, self.lv2plugins, self.lv2params,
         self.sums, self.norms, self.sampleloops, self.samplemultis, self.samplelerps, self.debugsines,
         self.synths, self.sampsyns, self.lv2fxs, self.adsrs, self.bandpasses, self.edges) = ([] for _ in range(18))

    # settings (state.rs:103-109)
    def set_length(self, seconds: float): self.length_seconds = float(seconds)
    def set_render_samplerate(self, sr: int): self.render_samplerate = int(sr)
    def set_render_bitdepth(self, bd: int): self.render_bitdepth = int(bd)
    def set_output_file(self, f: str): self.output_file = str(f)
    # resources (state.rs:112-122)
    def load_sample(self, name, path, mode=""): self.samples.append((name, path, mode))
    def load_resource(self, name, path): self.resources.append((name, path))
    def load_midi_floww(self, name, path): self.midis.append((name, path))
    def declare_stream(self, name): self.streams.append(name)
    def load_lv2(self, name, uri): self.lv2plugins.append((name, uri))
    def parameter(self, plugin, name, value): self.lv2params.append((plugin, name, float(value)))
    # graph (state.rs:125-152)
    def add_sum(self, name, gain, angle): self.sums.append((name, gain, angle))
    def add_normalize(self, name, gain, angle): self.norms.append((name, gain, angle))
    def add_sampleloop(self, name, gain, angle, sample): self.sampleloops.append((name, gain, angle, sample))
    def add_sample_multi(self, name, gain, angle, sample, floww, note):
        self.samplemultis.append((name, gain, angle, sample, floww, int(note)))
    def add_sample_lerp(self, name, gain, angle, sample, floww, note, lerp_len):
        self.samplelerps.append((name, gain, angle, sample, floww, int(note), int(lerp_len)))
    def add_debug_sine(self, name, gain, angle, floww): self.debugsines.append((name, gain, angle, floww))
    def add_synth(self, name, gain, angle, floww, sq_vel, sq_z, sq_arr, tf_vel, tf_z, tf_arr, tr_vel, tr_arr):
        self.synths.append((name, gain, angle, floww, sq_vel, sq_z, list(sq_arr), tf_vel, tf_z, list(tf_arr),
                            tr_vel, list(tr_arr)))
    def add_sampsyn(self, name, gain, angle, floww, adsr_conf, resource):
        self.sampsyns.append((name, gain, angle, floww, list(adsr_conf), resource))
    def add_lv2fx(self, name, gain, angle, wet, plugin): self.lv2fxs.append((name, gain, angle, wet, plugin))
    def add_adsr(self, name, gain, angle, wet, floww, use_off, use_max, note, conf):
        self.adsrs.append((name, gain, angle, wet, floww, bool(use_off), bool(use_max), int(note), list(conf)))
    def add_bandpass(self, name, gain, angle, wet, lo, hi, pass_):
        self.bandpasses.append((name, gain, angle, wet, float(lo), float(hi), bool(pass_)))
    def connect(self, a, b): self.edges.append((a, b))
    def set_output(self, out): self.output_vertex = out

    # north-star paraphrases (BASELINE.json) kept as aliases of the real names
    add_samplemulti = add_sample_multi
    add_samplelerp = add_sample_lerp
    add_wavetable = add_sampsyn
    add_envelope = add_adsr


class State:
    """Host half of `State` (state.rs:27-47) bound to one device context."""

    def __init__(self, lib: Library, project_samplerate: int = 44100, buffer_length: int = 1024,
                 device: int = 0, assets: Optional[Assets] = None):
        # defaults: config.rs:58-77 ; main.rs:75-98 (render_sr 48000, bd 16, cs 0)
        self.lib = lib
        self.psr, self.bl = int(project_samplerate), int(buffer_length)
        self.ctx = Context(lib, self.psr, self.bl, device)
        self.assets = assets or Assets()
        self.loaded = False
        self.cs = 0
        self.render_sr = 48000
        self.bd = 16
        self.output_file = "outp.wav"
        self.output_vertex = ""
        self.cur_samples: List[Tuple[str, str, str]] = []
        self.cur_resources: List[Tuple[str, str]] = []
        self.sample_index: Dict[str, int] = {}
        self.resource_index: Dict[str, int] = {}
        self.floww_index: Dict[str, int] = {}
        self.names: List[str] = []
        self.name_map: Dict[str, int] = {}
        self.messages: List[str] = []

    def close(self):
        self.ctx.close()

    def _say(self, m: str):
        self.messages.append(m)

    def refresh(self, project: Project) -> bool:
        """State::refresh (state.rs:50-471) given the seed vectors a script produced."""
        self.loaded = False
        c = self.ctx
        if project.length_seconds is not None:
            self.cs = set_length_chunks(self.psr, project.length_seconds, self.bl)
        if project.render_samplerate is not None: self.render_sr = project.render_samplerate
        if project.render_bitdepth is not None: self.bd = project.render_bitdepth
        # (state.rs:80,165: the local starts as mem::take(&mut self.output_file), i.e. the PREVIOUS name, and is written
        # back after the script ran: a script that does not call set_output_file keeps the previous name)
        if project.output_file is not None: self.output_file = project.output_file
        self.output_vertex = project.output_vertex

        # samples: keep what did not change, drop what left the script or changed (state.rs:201-219;
        # SampleBank::mark_dead + refresh, sample.rs:316-336): the device memory of a dropped sample is freed
        new = list(project.samples)
        for t in self.cur_samples:
            if t not in new:
                self._say(f"Info: sample \"{t[0]}\" will be removed from the sample bank.")
                idx = self.sample_index.pop(t[0], None)
                if idx is not None:
                    c.sample_remove(idx)
        excluded = []
        for (name, path, mode) in new:
            if (name, path, mode) in self.cur_samples and name in self.sample_index:
                continue
            try:
                w = self.assets.wav(path)
                if name in self.sample_index:
                    raise TermdawError(-5, f"there is already a sample with name \"{name}\" present.")
                self.sample_index[name] = c.sample_add(w.linear, w.channels, w.sample_rate,
                                                       sample_mode_from_string(mode))
            except (TermdawError, FileNotFoundError) as e:
                self._say(str(e))
                excluded.append(name)
        self.cur_samples = [t for t in new if t[0] not in excluded]
        if excluded:
            return False                                                    # do_excluding!, state.rs:188-199

        # resources: diffed as (name, path) tuples like the samples (state.rs:222-237): a resource that left the
        # script is dropped, a changed path under the same name is reloaded
        new_res = list(project.resources)
        for t in self.cur_resources:
            if t not in new_res:
                self._say(f"Info: resource \"{t[0]}\" will be removed from the resource bank.")
                idx = self.resource_index.pop(t[0], None)
                if idx is not None:
                    c.resource_remove(idx)
        loaded = []
        for (name, path) in new_res:
            if (name, path) in self.cur_resources and name in self.resource_index:
                loaded.append((name, path))
                continue
            try:
                if name in self.resource_index:
                    raise TermdawError(-5, f"there is already a resource with name \"{name}\" present.")
                self.resource_index[name] = c.resource_add(self.assets.blob(path))
                loaded.append((name, path))
            except (OSError, TermdawError) as e:
                self._say(str(e))
                self.cur_resources = loaded
                return False
        self.cur_resources = loaded

        # flowws: always reloaded (state.rs:240-250)
        c.floww_reset()
        self.floww_index = {}
        for (name, path) in project.midis:
            try:
                t, n, v = self.assets.floww(path)
            except FileNotFoundError as e:
                self._say(str(e))
                return False
            self.floww_index[name] = c.floww_add(t, n, v)
        for name in project.streams:
            self.floww_index[name] = c.floww_add([], [], [])

        # graph rebuild in the reference's kind order (state.rs:328-457)
        c.graph_reset()
        self.names, self.name_map = [], {}

        def reg(name, idx):
            self.names.append(name)
            self.name_map[name] = idx

        def need(table, key, name, category):
            if key not in table:
                raise KeyError(f"Could not get {category} index for vertex \"{name}\".")
            return table[key]

        try:
            for (name, gain, angle) in project.sums:
                reg(name, c.add_sum(gain, angle))
            for (name, gain, angle) in project.norms:
                reg(name, c.add_normalize(gain, angle))
            for (name, gain, angle, sample) in project.sampleloops:
                reg(name, c.add_sample_loop(gain, angle, need(self.sample_index, sample, name, "sample")))
            for (name, gain, angle, sample, floww, note) in project.samplemultis:
                reg(name, c.add_sample_multi(gain, angle, need(self.sample_index, sample, name, "sample"),
                                             need(self.floww_index, floww, name, "floww"), note))
            for (name, gain, angle, sample, floww, note, lerp_len) in project.samplelerps:
                reg(name, c.add_sample_lerp(gain, angle, need(self.sample_index, sample, name, "sample"),
                                            need(self.floww_index, floww, name, "floww"), note, lerp_len))
            for (name, gain, angle, floww) in project.debugsines:
                reg(name, c.add_debug_sine(gain, angle, need(self.floww_index, floww, name, "floww")))
            for (name, gain, angle, floww, sqv, sqz, sqa, tfv, tfz, tfa, trv, tra) in project.synths:
                reg(name, c.add_synth(gain, angle, need(self.floww_index, floww, name, "floww"),
                                      sqv, sqz, sqa, tfv, tfz, tfa, trv, tra))
            for (name, gain, angle, floww, conf, resource) in project.sampsyns:
                reg(name, c.add_sampsyn(gain, angle, need(self.floww_index, floww, name, "floww"), conf,
                                        need(self.resource_index, resource, name, "resource")))
            for (name, gain, angle, wet, plugin) in project.lv2fxs:
                reg(name, c.add_lv2fx(gain, angle, wet, 0))                 # always rejected on this path
            for (name, gain, angle, wet, floww, use_off, use_max, note, conf) in project.adsrs:
                reg(name, c.add_adsr(gain, angle, wet, need(self.floww_index, floww, name, "floww"),
                                     use_off, use_max, note, conf))
            for (name, gain, angle, wet, lo, hi, pass_) in project.bandpasses:
                reg(name, c.add_bandpass(gain, angle, wet, lo, hi, pass_))
        except (KeyError, TermdawError) as e:
            self._say(str(e))
            return False

        for (a, b) in project.edges:                                        # state.rs:459, graph.rs:80-96
            if a not in self.name_map:
                self._say(f"TermDaw: warning: vertex \"{a}\" cannot be found and thus can't be connected.")
                continue
            if b not in self.name_map:
                self._say(f"TermDaw: warning: vertex \"{b}\" cannot be found and thus can't be connected to.")
                continue
            c.connect(self.name_map[a], self.name_map[b])

        if self.output_vertex in self.name_map:                             # state.rs:461
            c.set_output(self.name_map[self.output_vertex])
        ok, unreachable = c.check_graph()
        for i in unreachable:
            self._say(f"TermDaw: warning: vertex \"{self.names[i]}\" does not reach output.")
        if not ok:
            self._say("TermDaw: graph check failed!")
            return False
        c.reset_normalize()                                                 # state.rs:467
        self.loaded = True
        return True

    def scan_exact(self):
        """State::scan_exact (state.rs:473-475)."""
        self.ctx.normalize_scan(self.cs)

    def render(self, want_pcm: bool = True):
        """State::render (state.rs:477-577); returns the PCM data-chunk bytes."""
        return self.ctx.render(self.cs, self.render_sr, self.bd, want_pcm=want_pcm)

    def render_to_file(self, path: Optional[str] = None) -> str:
        from .wav import write_wav
        pcm = self.render()
        path = path or self.output_file
        write_wav(path, pcm, self.render_sr, self.bd, channels=2)
        return path

    def normalization_values(self) -> Dict[str, float]:
        """Graph::print_normalization_values (graph.rs:213-220)."""
        out = {}
        for i, n in enumerate(self.names):
            v = self.ctx.get_normalize_value(i)
            if v > 0.0:
                out[n] = v
        return out
