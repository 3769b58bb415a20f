"""Pins the oracle against every golden vector the reference's own tests hold for this path:
the five ADSR known-answer tests of adsr.rs:116-205 (tolerance 1e-3, as in the reference)."""
import numpy as np

from termdaw_b200._capi import c_f32p

ADSR, ADS, R, R_RT = 0, 1, 2, 3


def ev(oracle, conf, fn, t, aux=0.0):
    a = np.asarray(conf, np.float32)
    return float(oracle.adsr_eval(a.ctypes.data_as(c_f32p), a.size, fn, t, aux))


def close(x, y):
    return abs(x - y) < 0.001


def test_adsr_0(oracle):   # adsr.rs:121-133
    conf = [1.0, 1.0, 0.5, 1.0, 0.25, 1.0]
    for t, v in [(0, 0), (0.5, 0.5), (1, 1), (1.5, 0.75), (2, 0.5), (2.5, 0.375), (3, 0.25), (3.5, 0.125),
                 (4, 0), (8, 0)]:
        assert close(ev(oracle, conf, ADSR, t), v), (t, v)


def test_adsr_1(oracle):   # adsr.rs:136-150
    conf = [1.0, 1.0, 0.5, 1.0, 0.25, 1.0]
    assert close(ev(oracle, conf, ADSR, 0.0), 0.0)
    for t, v in [(0.5, 0.5), (1, 1), (1.5, 0.75), (2, 0.5), (2.5, 0.375), (3, 0.25), (7, 0.25)]:
        assert close(ev(oracle, conf, ADS, t), v), (t, v)
    for t, v in [(0, 0.25), (0.5, 0.125), (1, 0), (9, 0)]:
        assert close(ev(oracle, conf, R, t, 0.25), v), (t, v)


def test_adsr_2(oracle):   # adsr.rs:153-165
    conf = [1.0, 1.0, 0.5, 2.0, 0.25, 1.0]
    assert close(ev(oracle, conf, ADSR, 0.0), 0.0)
    for t, v in [(0.5, 0.5), (1, 1), (1.5, 0.75), (2, 0.5), (3, 0.375)]:
        assert close(ev(oracle, conf, ADS, t), v), (t, v)
    for t, v in [(0, 0.375), (0.5, 0.1875), (1, 0), (9, 0)]:
        assert close(ev(oracle, conf, R, t, 0.375), v), (t, v)


def test_adsr_3(oracle):   # adsr.rs:168-180
    conf = [1.0, 1.0, 0.5, 2.0, 0.25, 1.0]
    for t, v in [(0.5, 0.5), (1, 1), (1.5, 0.75), (2, 0.5), (3, 0.375)]:
        assert close(ev(oracle, conf, ADS, t), v), (t, v)
    for t, v in [(0, 0.375), (0.5, 0.1875), (1, 0), (9, 0)]:
        assert close(ev(oracle, conf, R_RT, t, 3.0), v), (t, v)


def test_adsr_4(oracle):   # adsr.rs:183-204 — inverted "dip" envelope, 9-float conf
    conf = [1.0, 1.0, 0.0, 0.5, 0.0, 0.5, 0.0, 1.0, 1.0]
    for t, v in [(0, 1), (0.5, 0.5), (1, 0), (1.5, 0), (2, 0), (2.5, 0.5), (3, 1), (4, 1), (8, 1)]:
        assert close(ev(oracle, conf, ADSR, t), v), (t, v)


def test_adsr_conf_lengths(oracle):   # adsr.rs:94-114: 0, 6 or 9 floats; anything else is rejected
    assert ev(oracle, [], 4, 0.0) == 0.0                       # default conf: max_vel 0
    assert ev(oracle, [1, 1, 0.5, 1, 0.25, 1], 4, 0.0) == 1.0  # hit_conf: attack_vel 1
    assert np.isnan(ev(oracle, [1, 2, 3], ADSR, 0.0))
