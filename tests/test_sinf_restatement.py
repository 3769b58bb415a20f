"""The device's sinf (tdg_math.cuh: m_sinf_glibc, a restatement of glibc's double-precision sinf) against the
host libm over EVERY f32 argument.  The oracle calls the host's sinf where the reference calls Rust's f32::sin
(extensions.rs:452, synth.rs:23,28); with this pinned, the CUDA oscillators are bit-exact against the oracle."""
import os
import subprocess

from helpers import ROOT


def test_m_sinf_glibc_equals_host_sinf_for_all_f32(tmp_path):
    exe = str(tmp_path / "sinf_check")
    flags = ["-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-pthread"]
    try:
        if " fma " in open("/proc/cpuinfo").read():
            flags.append("-mfma")              # fma() is exact either way; this only makes it an instruction
    except OSError:
        pass
    subprocess.check_call(["g++", *flags, os.path.join(ROOT, "tests", "sinf_check.cpp"), "-o", exe])
    stride = int(os.environ.get("TDG_SINF_STRIDE", "1"))
    out = subprocess.check_output([exe, str(stride)], text=True).split()
    checked, bad = int(out[1]), int(out[3])
    assert checked >= (1 << 32) // stride
    assert bad == 0, out
