"""Batch render straight to WAV files: the output stage of `State::render` (state.rs:508-531, 545-574) for many
projects, with the file writes taken off the render's critical path.

The reference hands hound one sample at a time while it renders.  Here a group of projects is rendered by ONE batched
call (`tdg_normalize_render_many`: scan + render, PCM copied device -> host into PINNED staging buffers from
`tdg_host_alloc`), and while the device works on the next group a pool of host threads writes the previous group's
files: RIFF header (termdaw_b200/wav.py, byte-exact hound layout) + the pinned block, one `write` each.  Two sets of
staging buffers alternate, so a buffer is only reused once its file is on disk.
"""
from __future__ import annotations

import ctypes as C
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence

import numpy as np

from ._capi import Library, normalize_render_many, render_many
from .state import State
from .wav import wav_header


class PinnedBlock:
    """`bytes` of page-locked host memory from tdg_host_alloc, viewed as a uint8 array."""

    def __init__(self, lib: Library, nbytes: int):
        self.lib, self.nbytes = lib, max(int(nbytes), 1)
        self.ptr = lib.host_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError(f"tdg_host_alloc({self.nbytes}) failed")
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr))

    def close(self):
        if self.ptr:
            self.array = None
            self.lib.host_free(self.ptr)
            self.ptr = None


def _write_file(path: str, pcm: np.ndarray, sr: int, bd: int):
    with open(path, "wb") as f:
        f.write(wav_header(pcm.size, sr, bd, 2))
        f.write(memoryview(pcm))
    return path


def render_projects_to_files(lib: Library, states: Sequence[State], paths: Sequence[str], group: int = 64,
                             writers: int = 4, scan: bool = True) -> List[str]:
    """normalize (scan) + render of every project in `states` into `paths`, `group` projects per batched device call,
    file writes overlapping the next group's render.  Returns the paths in order."""
    assert len(states) == len(paths)
    caps = [st.cs * st.bl * 2 * (st.bd // 8) for st in states]
    sets: List[Optional[List[PinnedBlock]]] = [None, None]
    pending = [[], []]
    call = normalize_render_many if scan else render_many
    with ThreadPoolExecutor(max_workers=max(1, writers)) as pool:
        try:
            for gi, g0 in enumerate(range(0, len(states), group)):
                k = gi & 1
                for fut in pending[k]:                               # this staging set's files are on disk
                    fut.result()
                pending[k] = []
                idx = list(range(g0, min(len(states), g0 + group)))
                if sets[k] is None or any(sets[k][j].nbytes < caps[i] for j, i in enumerate(idx)) or len(sets[k]) < len(idx):
                    if sets[k]:
                        for b in sets[k]:
                            b.close()
                    sets[k] = [PinnedBlock(lib, caps[i]) for i in idx]
                outs = [sets[k][j].array[: caps[i]] for j, i in enumerate(idx)]
                sub = [states[i] for i in idx]
                pcms = call(lib, [st.ctx for st in sub], [st.cs for st in sub], [st.render_sr for st in sub],
                            [st.bd for st in sub], want_pcm=True, outs=outs)
                for st, pcm, i in zip(sub, pcms, idx):
                    pending[k].append(pool.submit(_write_file, paths[i], pcm, st.render_sr, st.bd))
            for k in (0, 1):
                for fut in pending[k]:
                    fut.result()
        finally:
            for k in (0, 1):
                for fut in pending[k]:
                    fut.cancel()
    for s in sets:
        if s:
            for b in s:
                b.close()
    return list(paths)
