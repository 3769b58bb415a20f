"""MIDI -> floww (termdaw_b200/midi.py; reference call site floww.rs:40-48).  The floww crate is not
vendored, so these tests pin the DEFINITION the module documents (SMF 1.0 timing, velocity / 127,
note-off = velocity 0, stable tick order), not upstream behaviour."""
import struct

import numpy as np
import pytest

from termdaw_b200.midi import MidiError, parse_midi, write_midi


def _smf(fmt, division, tracks):
    out = b"MThd" + struct.pack(">IHHH", 6, fmt, len(tracks), division)
    for t in tracks:
        out += b"MTrk" + struct.pack(">I", len(t)) + t
    return out


def test_default_tempo_and_velocity_scaling():
    # 480 ticks per quarter at the default 120 BPM: one quarter = 0.5 s
    trk = bytes([0x00, 0x90, 60, 127, 0x83, 0x60, 0x80, 60, 64,        # +480 ticks: note off
                 0x00, 0x90, 64, 1, 0x81, 0x70, 0x90, 64, 0,           # +240 ticks: note-on vel 0 = off
                 0x00, 0xFF, 0x2F, 0x00])
    t, n, v = parse_midi(_smf(0, 480, [trk]))
    assert t.dtype == n.dtype == v.dtype == np.float32
    assert t.tolist() == [0.0, 0.5, 0.5, 0.75]
    assert n.tolist() == [60.0, 60.0, 64.0, 64.0]
    assert v.tolist() == [1.0, 0.0, float(np.float32(1) / np.float32(127)), 0.0]


def test_tempo_map_applies_across_tracks_and_running_status():
    conductor = bytes([0x00, 0xFF, 0x51, 0x03, 0x07, 0xA1, 0x20,       # 500 000 us
                       0x83, 0x60, 0xFF, 0x51, 0x03, 0x0F, 0x42, 0x40, # at tick 480: 1 000 000 us (60 BPM)
                       0x00, 0xFF, 0x2F, 0x00])
    # running status: second and third note-on reuse 0x90; a controller change in between is ignored
    notes = bytes([0x00, 0x90, 40, 100, 0x83, 0x60, 41, 100, 0x00, 0xB0, 7, 90, 0x83, 0x60, 0x90, 42, 100,
                   0x00, 0xFF, 0x2F, 0x00])
    t, n, v = parse_midi(_smf(1, 480, [conductor, notes]))
    assert n.tolist() == [40.0, 41.0, 42.0]
    assert t.tolist() == [0.0, 0.5, 1.5]                               # the third quarter lasts 1 s


def test_order_is_tick_then_file_order():
    a = bytes([0x00, 0x90, 50, 10, 0x60, 0x80, 50, 0, 0x00, 0xFF, 0x2F, 0x00])
    b = bytes([0x60, 0x90, 51, 20, 0x00, 0xFF, 0x2F, 0x00])
    t, n, v = parse_midi(_smf(1, 96, [a, b]))
    assert n.tolist() == [50.0, 50.0, 51.0]                            # at tick 96: track 0's off before track 1's on
    assert v[1] == 0.0 and v[2] > 0.0
    assert np.all(np.diff(t) >= 0)


def test_smpte_division_and_sysex_skipped():
    div = ((256 - 25) << 8) | 40                                       # 25 fps, 40 ticks per frame: 1 ms per tick
    trk = bytes([0x00, 0xF0, 0x03, 1, 2, 0xF7, 0x87, 0x68, 0x90, 70, 64, 0x00, 0xFF, 0x2F, 0x00])   # +1000 ticks
    t, n, v = parse_midi(_smf(0, div, [trk]))
    assert t.tolist() == [1.0] and n.tolist() == [70.0]


def test_writer_round_trip_and_errors():
    ev = [(0, 36, 100), (120, 36, 0), (240, 38, 90), (480, 38, 0)]
    t, n, v = parse_midi(write_midi(ev, division=480, us_per_quarter=250000))
    assert t.tolist() == [0.0, 0.0625, 0.125, 0.25]
    assert n.tolist() == [36.0, 36.0, 38.0, 38.0]
    assert [x > 0 for x in v.tolist()] == [True, False, True, False]
    for bad in (b"", b"RIFFxxxxxxxxxxxx", _smf(0, 480, [bytes([0x00, 0x90, 60])]), _smf(0, 0, [b"\x00\xFF\x2F\x00"])):
        with pytest.raises(MidiError):
            parse_midi(bad)


def test_state_loads_a_midi_file_from_disk(tmp_path):
    """Assets.floww falls through to the file system, as FlowwBank::add_floww does; unreadable -> the reference's
    message and refresh() == False (state.rs:245-249)."""
    from termdaw_b200.state import Assets
    p = tmp_path / "notes.mid"
    p.write_bytes(write_midi([(0, 60, 127), (480, 60, 0)]))
    t, n, v = Assets().floww(str(p))
    assert t.tolist() == [0.0, 0.5] and n.tolist() == [60.0, 60.0] and v.tolist() == [1.0, 0.0]
    (tmp_path / "junk.mid").write_bytes(b"not midi")
    for path in (str(tmp_path / "junk.mid"), str(tmp_path / "missing.mid")):
        with pytest.raises(FileNotFoundError, match="Could not read midi file"):
            Assets().floww(path)


# ------------------------------------------------------------------------------------ native parser == python parser
def _vlq(v):
    out = [v & 0x7F]
    v >>= 7
    while v:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    return bytes(reversed(out))


def _random_smf(rng, fmt, ntracks, division):
    """A random but well-formed SMF: note on/off with and without running status, note-on velocity 0, tempo
    changes in any track, controllers / program changes / pitch bend / sysex / text meta events in between."""
    tracks = []
    for _ in range(ntracks):
        body, status = bytearray(), None
        for _ in range(int(rng.integers(5, 60))):
            body += _vlq(int(rng.choice([0, 0, 1, 7, 96, 480, 5000, 20000])))
            kind = int(rng.integers(10))
            if kind < 5:                                   # note on / off
                st = (0x90 if rng.random() < 0.6 else 0x80) | int(rng.integers(16))
                if st != status or rng.random() < 0.3:
                    body.append(st)
                    status = st
                body += bytes([int(rng.integers(128)), int(rng.choice([0, 1, 64, 127]))])
            elif kind == 5:                                # tempo
                body += b"\xFF\x51\x03" + int(rng.integers(100000, 2000000)).to_bytes(3, "big")
                status = None
            elif kind == 6:                                # controller / pitch bend (two data bytes)
                st = int(rng.choice([0xB0, 0xE0, 0xA0])) | int(rng.integers(16))
                body += bytes([st, int(rng.integers(128)), int(rng.integers(128))])
                status = st
            elif kind == 7:                                # program change / channel pressure (one data byte)
                st = int(rng.choice([0xC0, 0xD0])) | int(rng.integers(16))
                body += bytes([st, int(rng.integers(128))])
                status = st
            elif kind == 8:                                # sysex
                n = int(rng.integers(0, 6))
                body += b"\xF0" + _vlq(n + 1) + bytes(int(x) for x in rng.integers(0, 128, n)) + b"\xF7"
                status = None
            else:                                          # text meta
                n = int(rng.integers(0, 10))
                body += b"\xFF\x01" + _vlq(n) + bytes(int(x) for x in rng.integers(32, 127, n))
                status = None
        body += b"\x00\xFF\x2F\x00"
        tracks.append(bytes(body))
    return _smf(fmt, division, tracks)


def test_native_midi_parser_equals_python_parser(tmp_path):
    """termdaw_b200_render --dump-midi (csrc/host/tdh_assets.cpp) against termdaw_b200/midi.py on random files:
    the two hosts must hand the library the same f32 triples, bit for bit."""
    import os
    import subprocess
    from helpers import ROOT
    from termdaw_b200.csrc import build as B
    host = os.path.join(ROOT, "termdaw_b200", "csrc", "termdaw_b200_render")
    if not os.path.exists(host):
        B.build()
    rng = np.random.default_rng(77)
    cases = [(0, 1, 480), (1, 3, 96), (1, 5, 960), (2, 3, 480), (0, 1, ((256 - 25) << 8) | 40), (1, 2, ((256 - 29) << 8) | 80)]
    for k in range(18):
        fmt, ntr, div = cases[k % len(cases)]
        data = _random_smf(rng, fmt, ntr, div)
        path = tmp_path / f"r{k}.mid"
        path.write_bytes(data)
        t, n, v = parse_midi(data)
        out = subprocess.run([host, "--dump-midi", str(path)], capture_output=True, text=True, timeout=60)
        assert out.returncode == 0, out.stderr
        rows = [ln.split() for ln in out.stdout.splitlines()]
        assert len(rows) == t.size, k
        got = np.asarray(rows, np.float64).reshape(-1, 3).astype(np.float32) if rows else np.zeros((0, 3), np.float32)
        assert np.array_equal(got[:, 0], t) and np.array_equal(got[:, 1], n) and np.array_equal(got[:, 2], v), k
    # and the same refusals
    for bad in (b"", b"MThd" + b"\x00" * 10, _smf(0, 480, [bytes([0x00, 0x90, 60])])):
        path = tmp_path / "bad.mid"
        path.write_bytes(bad)
        out = subprocess.run([host, "--dump-midi", str(path)], capture_output=True, text=True, timeout=60)
        assert out.returncode == 1
        with pytest.raises(MidiError):
            parse_midi(bad)
