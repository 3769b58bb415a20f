"""Time-sharded render of ONE long project over the GPUs of a box.

One process per GPU (torchrun); every rank loads the same project, renders its own consecutive
range of chunks and writes its PCM bytes at the range's byte offset.  What crosses the segment
boundaries is tiny and goes through NCCL inside the library (`tdg_comm_init_rank`):
normalize peaks (max-allreduce / all-gather) and band-pass affine maps (all-gather), one round per
DAG level that holds such vertices.  `torch.distributed` is used only as the launcher-side plumbing
that carries the 128-byte ncclUniqueId from rank 0 to the others (any backend: gloo works).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from ._capi import Library, comm_unique_id, segment_output_frames
from .state import State


def shard_chunks(total_chunks: int, world: int) -> List[Tuple[int, int]]:
    """[(chunk0, chunks)] per rank: consecutive, covering, sizes differing by at most one chunk
    (earlier ranks take the remainder)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    base, rem = divmod(int(total_chunks), world)
    out, c0 = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((c0, n))
        c0 += n
    return out


def pcm_offset_bytes(bounds: List[Tuple[int, int]], rank: int, buffer_len: int, bitdepth: int,
                     lib: Optional[Library] = None, project_sr: int = 0, render_sr: int = 0) -> int:
    """Byte offset of a rank's segment inside the WAV data chunk (2 channels).  With a render rate below
    the project rate (state.rs:533) the segment's first output frame comes from the library's replay of
    the resampler index (`tdg_segment_output_frames`)."""
    if lib is not None and render_sr and project_sr > render_sr:
        first, _ = segment_output_frames(lib, project_sr, buffer_len, render_sr, bounds[rank][0], bounds[rank][1])
        return first * 2 * (bitdepth // 8)
    return bounds[rank][0] * buffer_len * 2 * (bitdepth // 8)


def broadcast_unique_id(lib: Optional[Library], dist, rank: int) -> bytes:
    """rank 0 asks NCCL for an id (through the library); everybody receives it over `dist`
    (a torch.distributed-like module: needs `broadcast_object_list`)."""
    box = [comm_unique_id(lib) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("broadcast of the ncclUniqueId failed")
    return bytes(uid)


class TimeShardedRender:
    """refresh -> [normalize] -> render of one project split over `world` ranks by time."""

    def __init__(self, state: State, rank: int, world: int, dist=None):
        self.st, self.rank, self.world = state, int(rank), int(world)
        self.bounds = shard_chunks(state.cs, world)
        self.has_comm = False
        if world > 1:
            if dist is None:
                raise ValueError("world > 1 needs a torch.distributed-like module for the id broadcast")
            uid = broadcast_unique_id(state.lib, dist, rank)
            state.ctx.comm_init_rank(world, rank, uid)
            self.has_comm = True

    def scan_exact(self):
        c0, n = self.bounds[self.rank]
        self.st.ctx.scan_segment(self.st.cs, c0, n, 0)

    def render(self, want_pcm: bool = True) -> Optional[np.ndarray]:
        c0, n = self.bounds[self.rank]
        return self.st.ctx.render_segment(self.st.cs, c0, n, 0, self.st.bd, want_pcm=want_pcm, render_sr=self.st.render_sr)

    def offset_bytes(self) -> int:
        return pcm_offset_bytes(self.bounds, self.rank, self.st.bl, self.st.bd, self.st.lib, self.st.psr, self.st.render_sr)

    def close(self):
        if self.has_comm:
            self.st.ctx.comm_destroy()
            self.has_comm = False
