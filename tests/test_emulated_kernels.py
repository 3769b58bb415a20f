"""The REAL kernels.cu + engine.cu on the CPU (tests/emu): the product sources compiled by g++ against a
stand-in CUDA runtime whose blocks run one after the other with the threads of a block as fibers.

This is test infrastructure, like the oracle: the emulated library lives in a temporary directory, is
never importable through `termdaw_b200.load_library()` and is not what any `-m gpu` test or the bench
measures — the product still refuses to work without an sm_100 device.  What it buys is that the host
planner, the launch sequences (levels, liveness, keep arena, exchange slots of the time sharding, the
pipelined group schedule) and the kernels' arithmetic (IEEE f32 with contraction off on both sides, no
device math library call in any kernel) are checked against the oracle in the `not gpu` run as well.
What it cannot see is concurrency: races between threads or streams only show on the GPU.

The test bodies are the `-m gpu` tests themselves, called with the emulated library in place of the
CUDA one — a subset sized for the CPU suite."""
import os
import sys

import pytest

from helpers import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

import test_golden as TG  # noqa: E402
import test_gpu_parity as TP  # noqa: E402
import test_pass_reuse as TR  # noqa: E402
import test_timeshard as TS  # noqa: E402
import test_planner_invariants as TI  # noqa: E402


@pytest.fixture(scope="session")
def emulib(tmp_path_factory):
    import build_emu
    from termdaw_b200._capi import Library
    path = build_emu.build(str(tmp_path_factory.mktemp("tdg_emu")))
    lib = Library(path, "tdg_")
    assert lib.exported()
    lib.set_launch_waves(0)          # the test projects are short: always walk the maximum number of chunks per block
    return lib


# ---- inlined sums, chunks per block (tests/test_planner_invariants.py) -------------------------------------------
def test_emu_inlined_sums(emulib, oracle):
    TI.check_sum_chain(emulib, oracle)
    TI.check_batch_projects(emulib, oracle, indices=range(8), scale=0.03)


def test_emu_random_dags(emulib, oracle):
    TI.check_random_dags(emulib, oracle, seeds=range(16))


def test_emu_voice_planner_fuzz(emulib, oracle):
    TI.check_voice_planner_fuzz(emulib, oracle, seeds=range(10))


def test_emu_fused_refresh(emulib, oracle):
    TI.check_fused_refresh(emulib, oracle, n=7, scale=0.03)


def test_emu_chunks_per_block(emulib, oracle):
    try:
        TI.check_launch_waves(emulib, oracle, indices=(0, 3), scale=0.05)
    finally:
        emulib.set_launch_waves(0)


# ---- parity against the oracle (tests/test_gpu_parity.py) ------------------------------------------------
def test_emu_config1_readme_example_through_the_resampler(emulib, oracle):
    TP.test_config1_readme_example(emulib, oracle, {})
    TP.test_config1_flow_a_running_max(emulib, oracle, {})


def test_emu_config2_synth_adsr_bandpass(emulib, oracle):
    TP.test_config2_synth_adsr_bandpass(emulib, oracle, {})


@pytest.mark.parametrize("index", [0, 1, 3, 4])
def test_emu_config5_projects(emulib, oracle, index):
    TP.test_config5_batch_projects(emulib, oracle, {}, index)


@pytest.mark.parametrize("bd", [8, 16, 24, 32])
def test_emu_bitdepths(emulib, oracle, bd):
    TP.test_bitdepths(emulib, oracle, bd)


def test_emu_errors_and_refresh_cycle(emulib, oracle):
    TP.test_unsupported_and_errors(emulib)
    TP.test_refresh_cycle_is_idempotent(emulib, oracle)


def test_emu_pipelined_groups(emulib, oracle):
    TP.test_pipelined_batches_equal_oracle(emulib, oracle, 3)


@pytest.mark.parametrize("name", TG.NAMES)
def test_emu_matches_golden(emulib, name):
    TG.test_cuda_path_matches_golden(emulib, name)


# ---- time sharding through virtual ranks (tests/test_timeshard.py) ------------------------------------------
def test_emu_time_sharding_bit_exact_and_resampled(emulib, oracle):
    TS.test_virtual_ranks_bit_exact_graph(emulib, oracle, 3)
    TS.test_virtual_ranks_resampled_render(emulib, oracle, 3)
    TS.test_resampled_segment_with_filter_and_errors(emulib, oracle)


def test_emu_time_sharding_filters_and_voices(emulib, oracle):
    TS.test_virtual_ranks_filters_and_voices(emulib, oracle, 2)


# ---- pass-invariant cones (tests/test_pass_reuse.py) -----------------------------------------------------------
@pytest.mark.parametrize("name", ["c1", "c2", "p0", "p5"])
def test_emu_pass_reuse_is_bit_identical(emulib, oracle, name):
    TR.test_reuse_is_bit_identical(emulib, oracle, name)


@pytest.mark.parametrize("ops,expect", [(["scan", "render"], [True]), (["scan", ("seek", 3), "render"], [False]),
                                        (["scan", "render", "render"], [True, False]), (["scan", "shorter", "render"], [False])])
def test_emu_pass_reuse_invalidation(emulib, oracle, ops, expect):
    TR.test_reuse_only_when_the_passes_cannot_differ(emulib, oracle, ops, expect)


def test_emu_pass_reuse_two_projects_on_one_lane(emulib, oracle):
    TR.test_two_projects_interleaved_on_one_lane(emulib, oracle)


# ---- playback / stream feed (tests/test_playback.py) --------------------------------------------------------------
import test_playback as TPB  # noqa: E402


@pytest.mark.parametrize("name", ["c1", "c2", "p3"])
def test_emu_play_equals_chunkwise_reference(emulib, oracle, name):
    TPB.check_play_equals_chunkwise_reference(emulib, oracle, name)


def test_emu_play_seek_render_and_stream_feed(emulib, oracle):
    TPB.check_play_then_seek_and_render(emulib, oracle)
    TPB.check_stream_feed(emulib, oracle)
    TPB.test_play_errors(emulib)
    TPB.check_async_play_ring(emulib, oracle)


def test_emu_keep_arena_never_outgrows_free_memory(emulib, oracle, monkeypatch):
    """A device with (pretended) 1 MB free: the scan must not try to keep anything, and scan + render still equal
    the oracle — oversized renders are evaluated twice instead of failing."""
    import numpy as np
    from helpers import pcm_to_int, run_project
    from termdaw_b200 import synthetic as S
    from termdaw_b200.state import State
    cfg = S.config_5_project(40, 0.2)                          # 6 s: its kept buffers would need several MB
    p, a, s = cfg
    st = State(emulib, s["psr"], s["bl"], assets=a)
    assert st.refresh(p)
    st.scan_exact(); st.render()
    assert st.ctx.last_pass_stats()["reused_vertices"] > 0     # normal device: the cone is kept (and the arena now has this size)
    st.close()
    monkeypatch.setenv("EMU_FREE_MB", "1")
    cfg = S.config_5_project(40, 1.5)                          # a much longer render would have to grow the arena
    p, a, s = cfg
    st = State(emulib, s["psr"], s["bl"], assets=a)
    assert st.refresh(p)
    st.scan_exact()
    pcm = st.render()
    assert st.ctx.last_pass_stats()["reused_vertices"] == 0
    st.close()
    want = run_project(oracle, *cfg, taps=False)
    x, y = pcm_to_int(pcm, 16), pcm_to_int(want["pcm"], 16)
    assert x.shape == y.shape and int(np.abs(x - y).max()) <= 1


def test_emu_batched_refresh_is_the_per_project_refresh(emulib, oracle):
    """bench.py's `rewind` (tdg_rewind_many: reload flowws, vertex state, transport, normalize maxima of a whole
    batch in one call) leaves every project exactly where the reference's stop + refresh would: rendering again
    gives the first render, bit for bit, and the oracle's."""
    import numpy as np
    import bench
    from helpers import pcm_to_int, run_project
    from termdaw_b200 import normalize_scan_many, render_many, synthetic as S
    from termdaw_b200.state import State
    cfgs = [S.config_5_project(i, 0.04) for i in (0, 3, 5, 7)]
    states = []
    for p, a, s in cfgs:
        st = State(emulib, s["psr"], s["bl"], assets=a)
        assert st.refresh(p)
        states.append((st, p, a))
    ctxs = [st.ctx for st, _, _ in states]
    args = ([st.cs for st, _, _ in states], [48000] * 4, [16] * 4)

    def render_all(scan):
        if scan:
            normalize_scan_many(emulib, ctxs, args[0])
        return [x.copy() for x in render_many(emulib, ctxs, *args)]
    first_a = render_all(False)                                # fresh projects, running-max flow
    bench.rewind(states, reupload_flowws=True)
    first = render_all(True)                                   # normalize, then render
    stale_a = render_all(False)                                # no refresh: the scan's maxima are still in the vertices
    assert any(not np.array_equal(a, b) for a, b in zip(first_a, stale_a))
    for reupload in (True, False, True):
        nbytes = bench.rewind(states, reupload_flowws=reupload)
        assert (nbytes > 0) == reupload
        for a, b in zip(first_a, render_all(False)):
            assert np.array_equal(a, b)
        bench.rewind(states, reupload_flowws=reupload)
        for a, b in zip(first, render_all(True)):
            assert np.array_equal(a, b)
    for (p, a, s), pcm in zip(cfgs, first):
        want = run_project(oracle, p, a, s, taps=False)
        x, y = pcm_to_int(pcm, 16), pcm_to_int(want["pcm"], 16)
        assert x.shape == y.shape and int(np.abs(x - y).max()) <= 1
    for st, _, _ in states:
        st.close()


def test_emu_heads_only_reuse(emulib, oracle):
    TR.test_heads_only_reuse_of_sample_multi_and_lerp(emulib, oracle)
