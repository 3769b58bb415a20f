"""ctypes binding of the C ABI declared in include/termdaw_b200.h.

The binder is parameterised by (shared library path, symbol prefix) because the CPU oracle under
oracle/ exports the same shape with the prefix ``tdo_``; only tests/, bench.py's cpu_baseline leg
and __graft_entry__.smoke() ever construct it with the oracle library.  The product
(`termdaw_b200.load_library()`) always binds `libtermdaw_b200.so` with prefix ``tdg_`` and raises if
the CUDA extension has not been built — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

c_f32p = C.POINTER(C.c_float)
c_u8p = C.POINTER(C.c_uint8)
c_u64p = C.POINTER(C.c_uint64)
c_u32p = C.POINTER(C.c_uint32)
c_i32p = C.POINTER(C.c_int)

ERRORS = {
    -1: "INVALID", -2: "CUDA", -3: "UNSUPPORTED_VERTEX", -4: "ADSR_CONF", -5: "SAMPLE",
    -6: "BITDEPTH", -7: "NO_OUTPUT", -8: "TOO_WIDE", -9: "CAPACITY", -10: "UNSORTED",
}


class TermdawError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class PassStats(C.Structure):
    _fields_ = [("device_ms", C.c_double), ("plan_ms", C.c_double), ("h2d_ms", C.c_double),
                ("d2h_ms", C.c_double), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("algorithmic_bytes", C.c_uint64),
                ("frames", C.c_uint64), ("voice_frames", C.c_uint64), ("reused_vertices", C.c_uint64),
                ("reused_bytes", C.c_uint64), ("inlined_vertices", C.c_uint64), ("inlined_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char_p), ("launches", C.c_uint64), ("device_ms", C.c_double),
                ("algorithmic_bytes", C.c_uint64)]


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray, t=c_f32p):
    return a.ctypes.data_as(t)


class Library:
    """One loaded shared library exposing <prefix>* symbols."""

    SIGS = {
        # name: (restype, argtypes)
        "sample_mode_from_string": (C.c_int, [C.c_char_p]),
        "ctx_create": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]),
        "ctx_destroy": (None, [C.c_void_p]),
        "last_error": (C.c_char_p, [C.c_void_p]),
        "kernel_launches": (C.c_uint64, [C.c_void_p]),
        "sample_add": (C.c_int, [C.c_void_p, c_f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]),
        "sample_remove": (C.c_int, [C.c_void_p, C.c_int]),
        "resource_remove": (C.c_int, [C.c_void_p, C.c_int]),
        "sample_frames": (C.c_int64, [C.c_void_p, C.c_int]),
        "sample_read": (C.c_int, [C.c_void_p, C.c_int, c_f32p, c_f32p, C.c_uint64]),
        "resource_add": (C.c_int, [C.c_void_p, c_u8p, C.c_uint64]),
        "floww_add": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_f32p, C.c_uint64]),
        "floww_reset": (C.c_int, [C.c_void_p]),
        "graph_reset": (C.c_int, [C.c_void_p]),
        "add_sum": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
        "add_normalize": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
        "add_sample_loop": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_int]),
        "add_sample_multi": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]),
        "add_sample_lerp": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]),
        "add_debug_sine": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_int]),
        "add_synth": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_int,
                                C.c_float, C.c_float, c_f32p, C.c_int,
                                C.c_float, C.c_float, c_f32p, C.c_int,
                                C.c_float, c_f32p, C.c_int]),
        "add_sampsyn": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_int, c_f32p, C.c_int, C.c_int]),
        "add_lv2fx": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int]),
        "add_adsr": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                               C.c_int, c_f32p, C.c_int]),
        "add_bandpass": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]),
        "connect": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
        "set_output": (C.c_int, [C.c_void_p, C.c_int]),
        "check_graph": (C.c_int, [C.c_void_p, c_i32p, C.c_int, c_i32p]),
        "reset_normalize": (C.c_int, [C.c_void_p]),
        "get_normalize_value": (C.c_float, [C.c_void_p, C.c_int]),
        "set_time": (C.c_int, [C.c_void_p, C.c_uint64]),
        "change_time": (C.c_uint64, [C.c_void_p, C.c_uint64, C.c_int]),
        "get_time": (C.c_uint64, [C.c_void_p]),
        "floww_set_time": (C.c_int, [C.c_void_p, C.c_uint64]),
        "normalize_scan": (C.c_int, [C.c_void_p, C.c_uint64]),
        "render": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, c_u8p, C.c_uint64, c_u64p]),
        "render_bound_bytes": (C.c_uint64, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]),
        "normalize_scan_many": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, c_u64p]),
        "render_many": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, c_u64p, c_u32p, c_u32p,
                                  C.POINTER(c_u8p), c_u64p, c_u64p]),
        "enable_taps": (C.c_int, [C.c_void_p, C.c_int]),
        "read_vertex_buffer": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, c_f32p, c_f32p]),
        "get_normalize_state": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_int]),
        "play": (C.c_int, [C.c_void_p, C.c_uint64, c_f32p, C.c_uint64, c_u64p]),
        "floww_append": (C.c_int, [C.c_void_p, C.c_int, c_f32p, c_f32p, c_f32p, C.c_uint64]),
    }
    # symbols only the CUDA library has (time sharding + measurement)
    GPU_ONLY = {
        "scan_segment": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
        "render_segment": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                     c_u8p, C.c_uint64, c_u64p]),
        "segment_output_frames": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, c_u64p, c_u64p]),
        "set_normalize_state": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_int]),
        "normalize_render_many": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, c_u64p, c_u32p, c_u32p,
                                            C.POINTER(c_u8p), c_u64p, c_u64p]),
        "comm_unique_id": (C.c_int, [c_u8p]),
        "comm_init_rank": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_u8p]),
        "comm_destroy": (C.c_int, [C.c_void_p]),
        "collectives": (C.c_uint64, [C.c_void_p]),
        "scan_segments_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_uint64, c_u64p, c_u64p]),
        "render_segments_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_uint64, c_u64p, c_u64p, C.c_uint32,
                                            C.c_uint32, C.POINTER(c_u8p), c_u64p, c_u64p]),
        "normalize_state_device_ptr": (C.c_void_p, [C.c_void_p, c_i32p]),
        "last_pass_stats": (C.c_int, [C.c_void_p, C.POINTER(PassStats)]),
        "set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
        "set_pass_reuse": (C.c_int, [C.c_void_p, C.c_int]),
        "host_alloc": (C.c_void_p, [C.c_uint64]),
        "host_free": (None, [C.c_void_p]),
        "play_begin": (C.c_int, [C.c_void_p, C.c_uint64]),
        "play_wait": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(c_f32p), c_u64p]),
        "set_oscillator_mode": (C.c_int, [C.c_void_p, C.c_int]),
        "rewind_many": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(c_f32p), C.POINTER(c_f32p), C.POINTER(c_f32p),
                                  c_u64p, c_u32p]),
        "refresh_normalize_render_many": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(c_f32p), C.POINTER(c_f32p), C.POINTER(c_f32p),
                                                    c_u64p, c_u32p, c_u64p, c_u32p, c_u32p, C.POINTER(c_u8p), c_u64p, c_u64p]),
        "reset_vertex_state": (C.c_int, [C.c_void_p]),
        "set_pipeline_group": (C.c_int, [C.c_void_p, C.c_uint32]),
        "set_launch_waves": (C.c_int, [C.c_uint32]),
        "set_inline_sums": (C.c_int, [C.c_void_p, C.c_int]),
        "last_pass_kernel_stats": (C.c_int, [C.c_void_p, C.POINTER(KernelStat), C.c_int]),
    }
    # symbols only the oracle has (scalar maths for the reference's known-answer tests)
    ORACLE_ONLY = {
        "adsr_eval": (C.c_float, [c_f32p, C.c_int, C.c_int, C.c_float, C.c_float]),
        "osc_eval": (C.c_float, [C.c_int, C.c_float, C.c_float, C.c_float]),
        "set_length_chunks": (C.c_uint64, [C.c_uint32, C.c_float, C.c_uint32]),
        "resample_whole": (C.c_int64, [c_f32p, c_f32p, C.c_uint64, C.c_uint32, C.c_uint32, c_f32p, c_f32p, C.c_uint64]),
    }

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing — build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        extra = self.GPU_ONLY if prefix == "tdg_" else self.ORACLE_ONLY
        for name, (res, args) in {**self.SIGS, **extra}.items():
            fn = getattr(self.dll, prefix + name)      # AttributeError == missing export
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    def exported(self) -> Sequence[str]:
        extra = self.GPU_ONLY if self.prefix == "tdg_" else self.ORACLE_ONLY
        return [self.prefix + n for n in {**self.SIGS, **extra}]


class Context:
    """One termdaw project on one device: the objects `State` owns (state.rs:27-47)."""

    def __init__(self, lib: Library, project_sr: int, buffer_len: int, device: int = 0):
        self.lib = lib
        self.sr, self.bl = int(project_sr), int(buffer_len)
        h = C.c_void_p()
        rc = lib.ctx_create(self.sr, self.bl, device, C.byref(h))
        if rc < 0 or not h:
            raise TermdawError(rc, "ctx_create failed (is a CUDA device visible and the extension built?)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers
    def _chk(self, rc: int) -> int:
        if rc < 0:
            raise TermdawError(rc, (self.lib.last_error(self.h) or b"").decode())
        return rc

    # -- banks
    def sample_add(self, linear, channels: int, src_sr: int, mode: int = 0) -> int:
        a = _f32(linear).ravel()
        return self._chk(self.lib.sample_add(self.h, _p(a), a.size, channels, src_sr, mode))

    def sample_remove(self, s: int):
        """SampleBank::mark_dead + refresh (sample.rs:316-336): free a sample that left the script or changed."""
        self._chk(self.lib.sample_remove(self.h, s))

    def resource_remove(self, r: int):
        self._chk(self.lib.resource_remove(self.h, r))

    def sample_frames(self, s: int) -> int:
        return self._chk(self.lib.sample_frames(self.h, s))

    def sample_read(self, s: int):
        n = self.sample_frames(s)
        l, r = np.empty(n, np.float32), np.empty(n, np.float32)
        self._chk(self.lib.sample_read(self.h, s, _p(l), _p(r), n))
        return l, r

    def resource_add(self, blob: bytes) -> int:
        a = np.frombuffer(blob, dtype=np.uint8).copy()
        return self._chk(self.lib.resource_add(self.h, _p(a, c_u8p), a.size))

    def floww_add(self, t, note, vel) -> int:
        t, note, vel = _f32(t), _f32(note), _f32(vel)
        assert t.size == note.size == vel.size
        return self._chk(self.lib.floww_add(self.h, _p(t), _p(note), _p(vel), t.size))

    def floww_reset(self):
        self._chk(self.lib.floww_reset(self.h))

    # -- graph
    def graph_reset(self):
        self._chk(self.lib.graph_reset(self.h))

    def add_sum(self, gain, angle):
        return self._chk(self.lib.add_sum(self.h, gain, angle))

    def add_normalize(self, gain, angle):
        return self._chk(self.lib.add_normalize(self.h, gain, angle))

    def add_sample_loop(self, gain, angle, sample):
        return self._chk(self.lib.add_sample_loop(self.h, gain, angle, sample))

    def add_sample_multi(self, gain, angle, sample, floww, note):
        return self._chk(self.lib.add_sample_multi(self.h, gain, angle, sample, floww, note))

    def add_sample_lerp(self, gain, angle, sample, floww, note, lerp_len):
        return self._chk(self.lib.add_sample_lerp(self.h, gain, angle, sample, floww, note, lerp_len))

    def add_debug_sine(self, gain, angle, floww):
        return self._chk(self.lib.add_debug_sine(self.h, gain, angle, floww))

    def add_synth(self, gain, angle, floww, sq_vol, sq_z, sq_adsr, tf_vol, tf_z, tf_adsr, tr_vol, tr_adsr):
        a, b, c = _f32(sq_adsr), _f32(tf_adsr), _f32(tr_adsr)
        return self._chk(self.lib.add_synth(self.h, gain, angle, floww, sq_vol, sq_z, _p(a), a.size,
                                            tf_vol, tf_z, _p(b), b.size, tr_vol, _p(c), c.size))

    def add_sampsyn(self, gain, angle, floww, adsr, resource):
        a = _f32(adsr)
        return self._chk(self.lib.add_sampsyn(self.h, gain, angle, floww, _p(a), a.size, resource))

    def add_lv2fx(self, gain, angle, wet, plugin):
        return self._chk(self.lib.add_lv2fx(self.h, gain, angle, wet, plugin))

    def add_adsr(self, gain, angle, wet, floww, use_off, use_max, note, conf):
        a = _f32(conf)
        return self._chk(self.lib.add_adsr(self.h, gain, angle, wet, floww, int(bool(use_off)),
                                           int(bool(use_max)), note, _p(a), a.size))

    def add_bandpass(self, gain, angle, wet, lo, hi, pass_):
        return self._chk(self.lib.add_bandpass(self.h, gain, angle, wet, lo, hi, int(bool(pass_))))

    def connect(self, a: int, b: int) -> bool:
        return bool(self.lib.connect(self.h, a, b))

    def set_output(self, v: int):
        self._chk(self.lib.set_output(self.h, v))

    def check_graph(self):
        cap = 4096
        arr = (C.c_int * cap)()
        n = C.c_int(0)
        ok = self.lib.check_graph(self.h, arr, cap, C.byref(n))
        return bool(ok), [arr[i] for i in range(min(n.value, cap))]

    def reset_vertex_state(self):
        """State::refresh on an unchanged project: every vertex back to its constructor state."""
        self._chk(self.lib.reset_vertex_state(self.h))

    def reset_normalize(self):
        self._chk(self.lib.reset_normalize(self.h))

    def get_normalize_value(self, v: int) -> float:
        return float(self.lib.get_normalize_value(self.h, v))

    # -- transport
    def set_time(self, frame: int):
        self._chk(self.lib.set_time(self.h, frame))

    def change_time(self, delta: int, plus: bool) -> int:
        return int(self.lib.change_time(self.h, delta, int(plus)))

    def get_time(self) -> int:
        return int(self.lib.get_time(self.h))

    def floww_set_time(self, frame: int):
        self._chk(self.lib.floww_set_time(self.h, frame))

    # -- hot path
    def normalize_scan(self, chunks: int):
        self._chk(self.lib.normalize_scan(self.h, chunks))

    def render(self, chunks: int, render_sr: int, bitdepth: int, want_pcm: bool = True) -> Optional[np.ndarray]:
        """Returns the PCM data-chunk bytes (uint8 array) the reference would hand to hound."""
        cap = int(self.lib.render_bound_bytes(self.h, chunks, render_sr, bitdepth))
        frames = C.c_uint64(0)
        if want_pcm:
            out = np.empty(cap, np.uint8)
            self._chk(self.lib.render(self.h, chunks, render_sr, bitdepth, _p(out, c_u8p), cap, C.byref(frames)))
            return out[: frames.value * 2 * (bitdepth // 8)]
        self._chk(self.lib.render(self.h, chunks, render_sr, bitdepth, None, 0, C.byref(frames)))
        self.last_frames = frames.value
        return None

    def enable_taps(self, on: bool = True):
        self._chk(self.lib.enable_taps(self.h, int(on)))

    def read_vertex_buffer(self, vertex: int, frame0: int, n: int):
        l, r = np.empty(n, np.float32), np.empty(n, np.float32)
        self._chk(self.lib.read_vertex_buffer(self.h, vertex, frame0, n, _p(l), _p(r)))
        return l, r

    def get_normalize_state(self, n_vertices: int):
        mx, sm = np.empty(n_vertices, np.float32), np.empty(n_vertices, np.float32)
        self._chk(self.lib.get_normalize_state(self.h, _p(mx), _p(sm), n_vertices))
        return mx, sm

    def floww_append(self, floww: int, t, note, vel):
        """tdg_floww_append + the cursor re-seat of the reference's Feed handler (stream_workflow.rs:63-71)."""
        t, note, vel = _f32(t), _f32(note), _f32(vel)
        self._chk(self.lib.floww_append(self.h, floww, _p(t), _p(note), _p(vel), t.size))
        self.floww_set_time(self.get_time())

    def play(self, chunks: int) -> np.ndarray:
        """tdg_play: the playback feed — `chunks` more chunks from the current transport position, nothing
        rewound; returns the interleaved (L, R) f32 stream (ui_workflow.rs:195-210)."""
        n = chunks * self.bl
        out = np.empty(2 * n, np.float32)
        frames = C.c_uint64(0)
        self._chk(self.lib.play(self.h, chunks, _p(out), n, C.byref(frames)))
        return out[: 2 * frames.value]

    # -- CUDA library only
    def play_begin(self, chunks: int) -> int:
        """tdg_play_begin: queue `chunks` more chunks of the playback feed; returns a ticket immediately."""
        return self._chk(self.lib.play_begin(self.h, int(chunks)))

    def play_wait(self, ticket: int) -> np.ndarray:
        """tdg_play_wait: the block of that ticket from the pinned ring, copied into an array [frames, 2]."""
        ptr, n = c_f32p(), C.c_uint64(0)
        self._chk(self.lib.play_wait(self.h, int(ticket), C.byref(ptr), C.byref(n)))
        return np.ctypeslib.as_array(ptr, shape=(n.value, 2)).copy()

    def scan_segment(self, total_chunks: int, chunk0: int, chunks: int, warmup_chunks: int):
        self._chk(self.lib.scan_segment(self.h, total_chunks, chunk0, chunks, warmup_chunks))

    def render_segment(self, total_chunks: int, chunk0: int, chunks: int, warmup_chunks: int, bitdepth: int,
                       want_pcm=True, render_sr: int = 0):
        """tdg_render_segment; render_sr = 0 (or >= the project rate) renders at the project rate."""
        cap = chunks * self.bl * 2 * (bitdepth // 8)            # the resampler only ever shortens
        frames = C.c_uint64(0)
        out = np.empty(cap if want_pcm else 0, np.uint8)
        self._chk(self.lib.render_segment(self.h, total_chunks, chunk0, chunks, warmup_chunks, render_sr, bitdepth,
                                          _p(out, c_u8p) if want_pcm else None, cap, C.byref(frames)))
        return out[: frames.value * 2 * (bitdepth // 8)] if want_pcm else None

    def set_normalize_state(self, mx, sm):
        mx, sm = _f32(mx), _f32(sm)
        self._chk(self.lib.set_normalize_state(self.h, _p(mx), _p(sm), mx.size))

    def normalize_state_device_ptr(self):
        n = C.c_int(0)
        p = self.lib.normalize_state_device_ptr(self.h, C.byref(n))
        return p, n.value

    def comm_init_rank(self, nranks: int, rank: int, unique_id: bytes):
        a = np.frombuffer(unique_id, dtype=np.uint8).copy()
        assert a.size == 128
        self._chk(self.lib.comm_init_rank(self.h, nranks, rank, _p(a, c_u8p)))

    def comm_destroy(self):
        self._chk(self.lib.comm_destroy(self.h))

    def collectives(self) -> int:
        return int(self.lib.collectives(self.h))

    def set_profiling(self, on: bool = True):
        self._chk(self.lib.set_profiling(self.h, int(on)))

    def set_pass_reuse(self, on: bool = True):
        """tdg_set_pass_reuse: render reads pass-invariant cones from the scan before it (device-wide)."""
        self._chk(self.lib.set_pass_reuse(self.h, int(on)))

    def set_oscillator_mode(self, libm_exact: bool = True):
        """tdg_set_oscillator_mode: True = glibc's sinf bit for bit (default), False = the 2-ulp f32 sine."""
        self._chk(self.lib.set_oscillator_mode(self.h, int(libm_exact)))

    def set_inline_sums(self, on: bool):
        self._chk(self.lib.set_inline_sums(self.h, 1 if on else 0))

    def set_pipeline_group(self, n: int):
        self._chk(self.lib.set_pipeline_group(self.h, int(n)))

    def last_pass_kernel_stats(self) -> dict:
        arr = (KernelStat * 32)()
        n = self.lib.last_pass_kernel_stats(self.h, arr, 32)
        return {arr[i].name.decode(): {"launches": int(arr[i].launches), "ms": float(arr[i].device_ms),
                                       "bytes": int(arr[i].algorithmic_bytes)} for i in range(n)}

    def last_pass_stats(self) -> dict:
        st = PassStats()
        self._chk(self.lib.last_pass_stats(self.h, C.byref(st)))
        return st.as_dict()


def normalize_render_many(lib: Library, ctxs, chunks, render_sr, bitdepth, want_pcm: bool = True, outs=None):
    """tdg_normalize_render_many: scan + render of every project, pipelined over project groups."""
    return render_many(lib, ctxs, chunks, render_sr, bitdepth, want_pcm, outs, fn="normalize_render_many")


def refresh_normalize_render_many(lib: Library, ctxs, flowws, chunks, render_sr, bitdepth, want_pcm: bool = True, outs=None):
    """tdg_refresh_normalize_render_many: 'stop' + 'refresh' + normalize + render of a batch of unchanged projects in one
    pipelined call.  `flowws` = (t, note, vel, counts, per_ctx): the flattened ctypes tables tdg_rewind_many takes
    (arrays of float pointers, event counts, flowws per context), or None for no floww reload."""
    pre = tuple(flowws) if flowws is not None else (None, None, None, None, None)
    return render_many(lib, ctxs, chunks, render_sr, bitdepth, want_pcm, outs, fn="refresh_normalize_render_many", pre_args=pre)


def render_many(lib: Library, ctxs: Sequence[Context], chunks: Sequence[int], render_sr: Sequence[int],
                bitdepth: Sequence[int], want_pcm: bool = True, outs: Optional[Sequence[np.ndarray]] = None,
                fn: str = "render_many", pre_args: tuple = ()):
    """tdg_render_many: one merged device schedule for `ctxs` (independent projects)."""
    call = getattr(lib, fn)
    n = len(ctxs)
    hs = (C.c_void_p * n)(*[c.h for c in ctxs])
    ch = (C.c_uint64 * n)(*chunks)
    rs = (C.c_uint32 * n)(*render_sr)
    bd = (C.c_uint32 * n)(*bitdepth)
    fw = (C.c_uint64 * n)()
    if want_pcm:
        # tdg_render_bound_bytes, without one FFI call per project: the resampler only ever shortens
        caps = [int(chunks[i]) * ctxs[i].bl * 2 * (int(bitdepth[i]) // 8) for i in range(n)]
        if outs is None:
            outs = [np.empty(c, np.uint8) for c in caps]
        ptrs = (c_u8p * n)(*[_p(o, c_u8p) for o in outs])
        cp = (C.c_uint64 * n)(*caps)
        rc = call(hs, n, *pre_args, ch, rs, bd, ptrs, cp, fw)
    else:
        rc = call(hs, n, *pre_args, ch, rs, bd, None, None, fw)
    if rc < 0:
        msgs = [(lib.last_error(c.h) or b"").decode() for c in ctxs]
        raise TermdawError(rc, next((m for m in msgs if m), ""))
    if want_pcm:
        return [outs[i][: fw[i] * 2 * (bitdepth[i] // 8)] for i in range(n)]
    return [int(x) for x in fw]


def normalize_scan_many(lib: Library, ctxs: Sequence[Context], chunks: Sequence[int]):
    n = len(ctxs)
    hs = (C.c_void_p * n)(*[c.h for c in ctxs])
    ch = (C.c_uint64 * n)(*chunks)
    rc = lib.normalize_scan_many(hs, n, ch)
    if rc < 0:
        msgs = [(lib.last_error(c.h) or b"").decode() for c in ctxs]
        raise TermdawError(rc, next((m for m in msgs if m), ""))


def comm_unique_id(lib: Library) -> bytes:
    """ncclGetUniqueId through the library (rank 0 of a time-sharded render); 128 bytes to broadcast."""
    a = np.zeros(128, np.uint8)
    rc = lib.comm_unique_id(_p(a, c_u8p))
    if rc < 0:
        raise TermdawError(rc, "NCCL is not available (libnccl.so.2 could not be loaded)")
    return a.tobytes()


def scan_segments_local(lib: Library, ctxs: Sequence[Context], total_chunks: int, bounds: Sequence[Sequence[int]]):
    """tdg_scan_segments_local: `ctxs` hold the same project; bounds = [(chunk0, chunks), ...]."""
    n = len(ctxs)
    hs = (C.c_void_p * n)(*[c.h for c in ctxs])
    c0 = (C.c_uint64 * n)(*[b[0] for b in bounds])
    cn = (C.c_uint64 * n)(*[b[1] for b in bounds])
    rc = lib.scan_segments_local(hs, n, total_chunks, c0, cn)
    if rc < 0:
        msgs = [(lib.last_error(c.h) or b"").decode() for c in ctxs]
        raise TermdawError(rc, next((m for m in msgs if m), ""))


def segment_output_frames(lib: Library, project_sr: int, buffer_len: int, render_sr: int, chunk0: int, chunks: int):
    """tdg_segment_output_frames: (first output frame, frame count) of chunks [chunk0, chunk0 + chunks) in the
    rendered file — the host's f64 replay of the resampler index when render_sr < project_sr."""
    first, n = C.c_uint64(0), C.c_uint64(0)
    rc = lib.segment_output_frames(project_sr, buffer_len, render_sr, chunk0, chunks, C.byref(first), C.byref(n))
    if rc < 0:
        raise TermdawError(rc, "tdg_segment_output_frames")
    return first.value, n.value


def render_segments_local(lib: Library, ctxs: Sequence[Context], total_chunks: int, bounds, bitdepth: int,
                          render_sr: int = 0):
    """tdg_render_segments_local; returns the per-segment PCM byte arrays (concatenate for the file)."""
    n = len(ctxs)
    hs = (C.c_void_p * n)(*[c.h for c in ctxs])
    c0 = (C.c_uint64 * n)(*[b[0] for b in bounds])
    cn = (C.c_uint64 * n)(*[b[1] for b in bounds])
    caps = [b[1] * ctxs[i].bl * 2 * (bitdepth // 8) for i, b in enumerate(bounds)]
    outs = [np.empty(max(c, 1), np.uint8) for c in caps]
    ptrs = (c_u8p * n)(*[_p(o, c_u8p) for o in outs])
    cp = (C.c_uint64 * n)(*caps)
    fw = (C.c_uint64 * n)()
    rc = lib.render_segments_local(hs, n, total_chunks, c0, cn, render_sr, bitdepth, ptrs, cp, fw)
    if rc < 0:
        msgs = [(lib.last_error(c.h) or b"").decode() for c in ctxs]
        raise TermdawError(rc, next((m for m in msgs if m), ""))
    return [outs[i][: fw[i] * 2 * (bitdepth // 8)] for i in range(n)]
