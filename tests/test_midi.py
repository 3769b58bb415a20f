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
