"""Shared test helpers: oracle loading, running a (project, assets) through a backend, comparisons."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libtermdaw_oracle.so")


def load_oracle():
    from termdaw_b200._capi import Library
    src = os.path.join(ROOT, "oracle", "termdaw_oracle.cpp")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-B"], stdout=subprocess.DEVNULL)
    return Library(ORACLE_SO, "tdo_")


def run_project(lib, project, assets, settings, flow=None, taps=True, chunks=None):
    """refresh -> [normalize] -> render through `lib` (oracle or CUDA).  Returns dict with the PCM
    bytes, per-vertex pre-quantisation taps of the render pass and the normalize values."""
    from termdaw_b200.state import State
    st = State(lib, settings["psr"], settings["bl"], assets=assets)
    try:
        ok = st.refresh(project)
        assert ok, st.messages
        if chunks is not None:
            st.cs = chunks
        if taps:
            st.ctx.enable_taps(True)
        flow = flow or settings.get("flow", "B")
        if flow == "B":
            st.scan_exact()
        pcm = st.render()
        n = st.cs * st.bl
        out = {"pcm": pcm.copy(), "names": list(st.names), "cs": st.cs, "bd": st.bd,
               "norm": st.normalization_values(), "taps": {}}
        if taps:
            ok, unreachable = st.ctx.check_graph()
            for i, name in enumerate(st.names):
                if i in unreachable:
                    continue
                out["taps"][name] = st.ctx.read_vertex_buffer(i, 0, n)
        return out
    finally:
        st.close()


def pcm_to_int(pcm: np.ndarray, bd: int) -> np.ndarray:
    """Decode the interleaved little-endian data chunk back to integers (hound layout)."""
    b = np.asarray(pcm, np.uint8)
    if bd == 8:
        return b.astype(np.int32) - 128
    if bd == 16:
        return b.view("<i2").astype(np.int32)
    if bd == 24:
        t = b.reshape(-1, 3).astype(np.int32)
        v = t[:, 0] | (t[:, 1] << 8) | (t[:, 2] << 16)
        return np.where(v & 0x800000, v - (1 << 24), v)
    return b.view("<i4").astype(np.int64)


def max_abs_diff(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b)
    d[both_nan] = 0.0
    return float(np.nanmax(d)) if d.size else 0.0
