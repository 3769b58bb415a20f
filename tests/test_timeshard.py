"""Time sharding: host-side partitioning / id broadcast on CPU with gloo
(world_size 2), and — on the GPU — the segment hand-off arithmetic through "virtual ranks": N
contexts holding the same project render N consecutive segments in one launch sequence and must
reproduce the unsharded render (normalize peaks exactly, band-pass state within the IIR tolerance)."""
import os
import socket
import sys

import numpy as np
import pytest

from helpers import ROOT, max_abs_diff, pcm_to_int, run_project
from termdaw_b200 import synthetic as S
from termdaw_b200.timeshard import pcm_offset_bytes, shard_chunks


def test_shard_chunks_partition():
    for total in [0, 1, 7, 141, 2813, 168750]:
        for world in [1, 2, 3, 4, 8]:
            b = shard_chunks(total, world)
            assert len(b) == world and b[0][0] == 0
            assert sum(n for _, n in b) == total
            for (c0, n), (c1, _) in zip(b, b[1:]):
                assert c0 + n == c1
            sizes = [n for _, n in b]
            assert max(sizes) - min(sizes) <= 1
    assert shard_chunks(168750, 8)[7] == (147657, 21093)
    assert pcm_offset_bytes(shard_chunks(100, 4), 2, 1024, 24) == 50 * 1024 * 6


def test_segment_output_frames_partition():
    """tdg_segment_output_frames is host arithmetic (the f64 replay of SincFixedIn's index, state.rs:533-550):
    segments tile the file, and the total is what the golden configs[0] render holds."""
    from termdaw_b200 import load_library, segment_output_frames
    lib = load_library()
    g = np.load(os.path.join(ROOT, "tests", "golden", "config1_readme_44k1_16bit.npz"))
    total_frames = g["pcm"].size // 4
    cs = int(g["cs"]) if "cs" in g.files else None
    for sr, bl, rsr, chunks in [(48000, 1024, 44100, cs or 141), (48000, 1024, 32000, 37), (96000, 256, 44100, 50),
                                (48000, 64, 44100, 40), (48000, 1024, 48000, 9), (48000, 1024, 0, 9)]:
        f0, n0 = segment_output_frames(lib, sr, bl, rsr, 0, chunks)
        assert f0 == 0
        for world in (1, 2, 3, 8):
            off = 0
            for c0, n in shard_chunks(chunks, world):
                first, cnt = segment_output_frames(lib, sr, bl, rsr, c0, n)
                assert first == off
                off += cnt
            assert off == n0
        if rsr == 0 or rsr >= sr:
            assert n0 == chunks * bl
    if cs is not None:
        assert segment_output_frames(lib, 48000, 1024, 44100, 0, cs)[1] == total_frames


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from termdaw_b200 import timeshard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the id itself needs NCCL + a GPU; the plumbing is what is tested here
    timeshard.comm_unique_id = lambda lib: bytes(range(128))
    uid = timeshard.broadcast_unique_id(None, dist, rank)
    bounds = timeshard.shard_chunks(2813, world)
    import torch
    t = torch.tensor([float(bounds[rank][1])])
    dist.all_reduce(t)
    q.put((rank, uid == bytes(range(128)), bounds[rank], int(t.item())))
    dist.destroy_process_group()


def test_unique_id_broadcast_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [g[1] for g in got] == [True, True]
    assert got[0][2] == (0, 1407) and got[1][2] == (1407, 1406)
    assert got[0][3] == got[1][3] == 2813


# ------------------------------------------------------------------------------------ GPU
def _sharded(gpulib, cfg, n_seg, flow="B"):
    from termdaw_b200 import render_segments_local, scan_segments_local
    from termdaw_b200.state import State
    p, a, s = cfg
    states = []
    for _ in range(n_seg):
        st = State(gpulib, s["psr"], s["bl"], assets=a)
        assert st.refresh(p), st.messages
        states.append(st)
    cs = states[0].cs
    bounds = shard_chunks(cs, n_seg)
    ctxs = [st.ctx for st in states]
    if flow == "B":
        scan_segments_local(gpulib, ctxs, cs, bounds)
    pcms = render_segments_local(gpulib, ctxs, cs, bounds, states[0].bd, states[0].render_sr)
    norms = [st.normalization_values() for st in states]
    bd = states[0].bd
    for st in states:
        st.close()
    return np.concatenate(pcms), norms, bd


@pytest.mark.gpu
@pytest.mark.parametrize("n_seg", [2, 3, 8])
def test_virtual_ranks_bit_exact_graph(gpulib, oracle, n_seg):
    """No IIR in the graph: sharded render == oracle's unsharded render, bit for bit, both flows."""
    cfg = S.config_1(0.5)
    cfg[0].set_render_samplerate(48000)          # project rate; the resampled case is tested below
    for flow in ("B", "A"):
        want = run_project(oracle, *cfg, flow=flow, taps=False)
        pcm, norms, bd = _sharded(gpulib, cfg, n_seg, flow)
        assert np.array_equal(pcm, want["pcm"]), (n_seg, flow)
        for nv in norms:
            assert nv == pytest.approx(want["norm"], rel=0, abs=0)


@pytest.mark.gpu
@pytest.mark.parametrize("n_seg", [2, 3, 8])
def test_virtual_ranks_resampled_render(gpulib, oracle, n_seg):
    """configs[0] at its real 44.1 kHz render rate, split by time (last row): the f64 resampler
    index is replayed from chunk 0 and the previous segment's last 512 frames arrive as a halo, so the
    sharded file is the unsharded GPU file bit for bit (and the oracle's within 1 LSB)."""
    from termdaw_b200 import segment_output_frames
    from termdaw_b200.state import State
    cfg = S.config_1(0.5)
    assert cfg[0].render_samplerate == 44100
    for flow in ("B", "A"):
        p, a, s = cfg
        st = State(gpulib, s["psr"], s["bl"], assets=a)
        assert st.refresh(p), st.messages
        if flow == "B":
            st.scan_exact()
        full = st.render()
        cs, bl, psr, bd = st.cs, st.bl, st.psr, st.bd
        st.close()
        pcm, norms, _ = _sharded(gpulib, cfg, n_seg, flow)
        assert np.array_equal(pcm, full), (n_seg, flow)
        want = run_project(oracle, *cfg, flow=flow, taps=False)
        x, y = pcm_to_int(pcm, bd), pcm_to_int(want["pcm"], want["bd"])
        assert x.shape == y.shape and int(np.abs(x - y).max()) <= 1
        # the file offsets the ranks would write at
        bounds = shard_chunks(cs, n_seg)
        off = 0
        for r, (c0, n) in enumerate(bounds):
            first, cnt = segment_output_frames(gpulib, psr, bl, 44100, c0, n)
            assert first == off and pcm_offset_bytes(bounds, r, bl, bd, gpulib, psr, 44100) == off * 2 * (bd // 8)
            off += cnt
        assert off * 2 * (bd // 8) == full.size


@pytest.mark.gpu
def test_resampled_segment_with_filter_and_errors(gpulib, oracle):
    """band pass in front of the resampler (state hand-off + halo together), and the documented refusals."""
    from termdaw_b200 import TermdawError
    from termdaw_b200.state import State
    cfg = S.config_2(0.05)
    cfg[0].set_render_samplerate(32000)
    want = run_project(oracle, *cfg, taps=False)
    pcm, norms, bd = _sharded(gpulib, cfg, 4)
    a, b = pcm_to_int(pcm, bd), pcm_to_int(want["pcm"], want["bd"])
    assert a.shape == b.shape
    assert int(np.abs(a - b).max()) <= 16
    # a lone resampled segment that does not start the pass has no history to read
    p, a_, s = cfg
    st = State(gpulib, s["psr"], s["bl"], assets=a_)
    assert st.refresh(p)
    with pytest.raises(TermdawError):
        st.ctx.render_segment(st.cs, 4, 4, 0, st.bd, render_sr=32000)
    # ... unless enough warm-up chunks are rendered in front of it (approximate mode, still resampled)
    out = st.ctx.render_segment(st.cs, 4, 4, 1, st.bd, render_sr=32000)
    from termdaw_b200 import segment_output_frames
    assert out.size == segment_output_frames(gpulib, s["psr"], s["bl"], 32000, 4, 4)[1] * 2 * (st.bd // 8)
    st.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n_seg", [2, 4])
def test_virtual_ranks_filters_and_voices(gpulib, oracle, n_seg):
    """synth -> adsr -> band pass -> normalize (configs[1]) and the 64-vertex graph (configs[3]) split by
    time: filter state crosses the boundaries as affine maps, voices/envelopes by event replay."""
    for cfg, lsb in ((S.config_2(0.1), 16), (S.config_4(0.002, 64), 16)):
        want = run_project(oracle, *cfg, taps=False)
        pcm, norms, bd = _sharded(gpulib, cfg, n_seg)
        a, b = pcm_to_int(pcm, bd), pcm_to_int(want["pcm"], want["bd"])
        assert a.shape == b.shape
        assert int(np.abs(a - b).max()) <= lsb, (n_seg, int(np.abs(a - b).max()))
        for nv in norms:
            for k, v in want["norm"].items():
                assert abs(nv[k] - v) <= 2e-5 * max(1.0, abs(v))


@pytest.mark.gpu
def test_segment_equals_full_pass_single_rank(gpulib, oracle):
    """world = 1: tdg_*_segment over the whole pass is the plain pass."""
    from termdaw_b200.state import State
    from termdaw_b200.timeshard import TimeShardedRender
    p, a, s = S.config_5_project(3, 0.1)
    st = State(gpulib, s["psr"], s["bl"], assets=a)
    assert st.refresh(p)
    ts = TimeShardedRender(st, 0, 1)
    ts.scan_exact()
    pcm = ts.render()
    want = run_project(oracle, p, a, s, taps=False)
    x, y = pcm_to_int(pcm, st.bd), pcm_to_int(want["pcm"], want["bd"])
    assert x.shape == y.shape and np.abs(x - y).max() <= 1
    st.close()
