#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/termdaw_oracle.cpp).

The reference (Rust) cannot be built or run in this environment and ships no golden audio, so these
fixtures are NOT reference output: they pin the oracle's own results (drift guard) and let the GPU box
compare the CUDA path against committed vectors in addition to the live oracle.

    python tests/golden/make_golden.py          # rewrites the fixtures
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cases():
    from termdaw_b200 import synthetic as S
    return {
        "config1_readme_44k1_16bit": (S.config_1(0.05), "B"),
        "config1_flowA_running_max": (S.config_1(0.05), "A"),
        "config2_synth_adsr_bandpass_24bit": (S.config_2(0.004), "B"),
        "config3_wavetable16_lerp": (S.config_3(0.0005, voices=16), "B"),
        "config4_mixed64": (S.config_4(0.00008, 64), "B"),
        "config5_project0": (S.config_5_project(0, 0.01), "B"),
        "config5_project5": (S.config_5_project(5, 0.01), "B"),
    }


def digest(l, r):
    """Small per-vertex summary of a tap: peak, f64 sums and 16 evenly spaced samples per channel."""
    idx = np.linspace(0, l.size - 1, 16).astype(np.int64) if l.size else np.zeros(0, np.int64)
    return np.concatenate([[np.nanmax(np.abs(np.concatenate([l, r]))) if l.size else 0.0,
                            np.nansum(l.astype(np.float64)), np.nansum(r.astype(np.float64)),
                            np.nansum(np.abs(l.astype(np.float64))), np.nansum(np.abs(r.astype(np.float64)))],
                           l[idx].astype(np.float64), r[idx].astype(np.float64)])


def main():
    from helpers import load_oracle, run_project
    lib = load_oracle()
    for name, (cfg, flow) in cases().items():
        out = run_project(lib, *cfg, flow=flow)
        names = sorted(out["taps"])
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            pcm=out["pcm"], bd=np.int64(out["bd"]), cs=np.int64(out["cs"]), flow=np.array(flow),
            tap_names=np.array(names), tap_digest=np.stack([digest(*out["taps"][n]) for n in names]),
            norm_names=np.array(sorted(out["norm"])), norm_values=np.array([out["norm"][k] for k in sorted(out["norm"])]))
        print(name, out["pcm"].size, "pcm bytes,", len(names), "taps")


if __name__ == "__main__":
    main()
