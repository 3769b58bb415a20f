"""Standard MIDI file -> floww event stream (reference: `FlowwBank::add_floww`,
floww.rs:40-48, which calls the `floww` crate's `read_floww_from_midi`).

PARITY UNPINNED.  The `floww` crate (0.1.x, Cargo.toml:24) and the `apres` MIDI reader under it are
not vendored in the reference and absent from this image; the reference holds no test or fixture for
this conversion ("MIDI velocity scaling and tempo handling: unknown").  What IS
pinned by termdaw's own code is the shape it consumes, and this module produces exactly that:

    Floww = Vec<(id, t_seconds: f32, note: f32, vel: f32)>        (floww.rs:74,105-115,131-135)
    sorted by time; a note-off is an event with vel <= 0.001      (floww.rs:112,134)

The conversion itself is DEFINED here, from the Standard MIDI File 1.0 specification:

  * header `MThd`: format 0/1/2, ntrks, division.  division > 0: ticks per quarter note, tempo from
    `FF 51 03` meta events (default 500 000 us/quarter = 120 BPM); a tempo change applies to every track
    from its tick on (format 1 keeps the tempo map in any track; all tracks are merged on one tick
    axis, format 2's independent sequences are concatenated in file order).  division < 0 (SMPTE):
    seconds = ticks / (frames_per_second * ticks_per_frame), 29 meaning 29.97.
  * events: delta times are variable-length quantities; running status is honoured; sysex (F0/F7) and
    meta events are skipped by their length; `FF 2F 00` ends a track.
  * note-on with velocity > 0  -> (t, key, velocity / 127)
    note-off, or note-on with velocity 0 -> (t, key, 0.0)
    every other channel message is ignored (termdaw's vertices read notes only); channels are merged.
  * seconds are accumulated in f64 over the tempo map (exact rational arithmetic per segment:
    ticks * us_per_quarter / division / 1e6) and rounded ONCE to f32, the type the reference stores.
  * ordering: by tick, then file order (track index, then position inside the track) — a stable merge, so
    a note-off written before a note-on at the same tick stays before it.  The drum / simple queries of
    floww.rs:99-141 depend on that order for same-frame events.
"""
from __future__ import annotations

import struct
from typing import List, Tuple, Union

import numpy as np


class MidiError(ValueError):
    pass


def _vlq(b: bytes, i: int) -> Tuple[int, int]:
    v = 0
    for _ in range(4):
        if i >= len(b):
            raise MidiError("truncated variable-length quantity")
        c = b[i]
        i += 1
        v = (v << 7) | (c & 0x7F)
        if not c & 0x80:
            return v, i
    raise MidiError("variable-length quantity longer than 4 bytes")


def _parse_track(b: bytes, track: int):
    """-> (notes [(tick, track, seq, key, vel)], tempos [(tick, track, seq, us_per_quarter)], end tick)"""
    notes, tempos = [], []
    i, tick, status, seq = 0, 0, 0, 0
    n = len(b)
    while i < n:
        dt, i = _vlq(b, i)
        tick += dt
        if i >= n:
            raise MidiError("track ends inside an event")
        c = b[i]
        if c == 0xFF:                                   # meta
            if i + 1 >= n:
                raise MidiError("truncated meta event")
            kind = b[i + 1]
            ln, j = _vlq(b, i + 2)
            data = b[j:j + ln]
            if len(data) != ln:
                raise MidiError("truncated meta event")
            i = j + ln
            status = 0                                  # meta / sysex cancel running status
            if kind == 0x51 and ln == 3:
                tempos.append((tick, track, seq, (data[0] << 16) | (data[1] << 8) | data[2]))
            elif kind == 0x2F:
                break
        elif c in (0xF0, 0xF7):                         # sysex
            ln, j = _vlq(b, i + 1)
            if j + ln > n:
                raise MidiError("truncated sysex event")
            i = j + ln
            status = 0
        else:
            if c & 0x80:
                status = c
                i += 1
            elif not status:
                raise MidiError("data byte without a running status")
            hi = status & 0xF0
            if hi in (0x80, 0x90, 0xA0, 0xB0, 0xE0):
                if i + 2 > n:
                    raise MidiError("truncated channel event")
                d0, d1 = b[i], b[i + 1]
                i += 2
                if hi == 0x90 and d1 > 0:
                    notes.append((tick, track, seq, d0 & 0x7F, d1 & 0x7F))
                elif hi == 0x80 or hi == 0x90:
                    notes.append((tick, track, seq, d0 & 0x7F, 0))
            elif hi in (0xC0, 0xD0):
                if i + 1 > n:
                    raise MidiError("truncated channel event")
                i += 1
            elif status in (0xF1, 0xF3):                # system common with one data byte
                i += 1
            elif status == 0xF2:
                i += 2
            # F6, F8..FE: no data
        seq += 1
    return notes, tempos, tick


def parse_midi(data: bytes) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Bytes of a Standard MIDI File -> (t_seconds f32[], note f32[], vel f32[]), time-sorted."""
    if len(data) < 14 or data[:4] != b"MThd":
        raise MidiError("not a Standard MIDI File (no MThd)")
    hlen, fmt, ntrks, division = struct.unpack(">IHHH", data[4:14])
    if hlen < 6:
        raise MidiError("bad MThd length")
    i = 8 + hlen
    tracks: List[bytes] = []
    while i + 8 <= len(data) and len(tracks) < ntrks:
        tag, ln = data[i:i + 4], struct.unpack(">I", data[i + 4:i + 8])[0]
        body = data[i + 8:i + 8 + ln]
        if len(body) != ln:
            raise MidiError("truncated chunk")
        i += 8 + ln
        if tag == b"MTrk":
            tracks.append(body)                         # alien chunks are skipped (SMF 1.0, section 2)
    if not tracks:
        raise MidiError("no MTrk chunk")
    notes, tempos = [], []
    offset = 0
    for k, tb in enumerate(tracks):
        nt, tp, end = _parse_track(tb, k)
        if fmt == 2 and offset:
            nt = [(t + offset, a, s, key, v) for (t, a, s, key, v) in nt]
            tp = [(t + offset, a, s, us) for (t, a, s, us) in tp]
        notes += nt
        tempos += tp
        if fmt == 2:
            offset += end
    notes.sort(key=lambda e: (e[0], e[1], e[2]))
    tempos.sort(key=lambda e: (e[0], e[1], e[2]))

    if division & 0x8000:                               # SMPTE: negative frames/s in the high byte
        fps = 256 - (division >> 8)
        fps_f = 29.97 if fps == 29 else float(fps)
        per_tick = 1.0 / (fps_f * float(division & 0xFF))
        secs = [tick * per_tick for (tick, _, _, _, _) in notes]
    else:
        if division == 0:
            raise MidiError("division of 0 ticks per quarter note")
        # piecewise-constant tempo map: seconds at the start of each segment, accumulated in f64
        seg_tick, seg_sec, seg_us = [0], [0.0], [500000]
        for (tick, _, _, us) in tempos:
            if tick == seg_tick[-1]:
                seg_us[-1] = us                         # several tempo events on one tick: the last one stands
                continue
            seg_sec.append(seg_sec[-1] + (tick - seg_tick[-1]) * seg_us[-1] / division / 1e6)
            seg_tick.append(tick)
            seg_us.append(us)
        secs = []
        k = 0
        for (tick, _, _, _, _) in notes:
            while k + 1 < len(seg_tick) and seg_tick[k + 1] <= tick:
                k += 1
            secs.append(seg_sec[k] + (tick - seg_tick[k]) * seg_us[k] / division / 1e6)
    t = np.asarray(secs, np.float64).astype(np.float32)
    note = np.asarray([e[3] for e in notes], np.float32)
    vel = (np.asarray([e[4] for e in notes], np.float32) / np.float32(127.0)).astype(np.float32)
    return t, note, vel


def read_floww_from_midi(path_or_bytes: Union[str, bytes, bytearray]):
    """`read_floww_from_midi(path)` of the floww crate (call site floww.rs:41), as defined above."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        return parse_midi(bytes(path_or_bytes))
    with open(path_or_bytes, "rb") as f:
        return parse_midi(f.read())


# ---- writer: used by the tests and by tools that want to hand termdaw a .mid of a synthetic stream ----
def _enc_vlq(v: int) -> bytes:
    out = [v & 0x7F]
    v >>= 7
    while v:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    return bytes(reversed(out))


def write_midi(events, division: int = 480, us_per_quarter: int = 500000) -> bytes:
    """events: [(tick, key, velocity 0..127)] (velocity 0 = note-off) -> format-0 SMF bytes."""
    body = bytearray()
    body += b"\x00\xFF\x51\x03" + us_per_quarter.to_bytes(3, "big")
    last = 0
    for tick, key, v in sorted(events, key=lambda e: e[0]):
        body += _enc_vlq(int(tick) - last)
        last = int(tick)
        body += bytes([0x90 if v > 0 else 0x80, int(key) & 0x7F, int(v) & 0x7F if v > 0 else 0x40])
    body += b"\x00\xFF\x2F\x00"
    return b"MThd" + struct.pack(">IHHH", 6, 0, 1, division) + b"MTrk" + struct.pack(">I", len(body)) + bytes(body)
