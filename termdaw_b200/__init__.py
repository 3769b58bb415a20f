"""termdaw_b200 — B200-native offline renderer for termdaw's audio graph.

Only what the hot path needs lives here: `csrc/` (CUDA kernels, host planner and the C ABI of
include/termdaw_b200.h) and the host-side mirror of the reference's project interface
(`state.Project` / `state.State`).  There is no CPU fallback: `load_library()` raises when the CUDA
extension is not built, and every render call fails with TDG_ERR_CUDA without an sm_100 device.
"""
from __future__ import annotations

import os

from ._capi import (Context, Library, TermdawError, comm_unique_id, normalize_render_many, normalize_scan_many, refresh_normalize_render_many, render_many,  # noqa: F401
                    render_segments_local, scan_segments_local, segment_output_frames)
from .state import Assets, DecodedWav, Project, State, set_length_chunks  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtermdaw_b200.so")     # the in-tree sm_100a build; there is no override and no fallback
_lib = None


def load_library() -> Library:
    """Bind the CUDA extension (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is None:
        _lib = Library(LIB_PATH, "tdg_")
    return _lib
