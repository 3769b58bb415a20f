"""Planner and launch decisions that must not change a single bit: inlined sums, chunks per block, the refresh pipelined
inside the batched call, buffer liveness on arbitrary DAGs, the voice planner's bookkeeping.

* A Sum vertex whose only consumer is a Sum, Normalize or Adsr vertex is evaluated inside that consumer's kernel
  (engine.cu "inlined sums", include/termdaw_b200.h tdg_set_inline_sums): PCM with the switch on == PCM with it
  off, byte for byte, == the oracle (byte-exact where the graph holds no band pass), and the statistics say that
  vertices were in fact inlined.
* The chunk kernels let a block walk up to 8 consecutive chunks as long as the grid stays large
  (tdg_set_launch_waves): 0 (always 8) and 2^31 (always 1) give the same bytes.

The bodies take the library as an argument: tests/test_emulated_kernels.py runs them on the CPU-emulated build."""
import numpy as np
import pytest

from termdaw_b200 import synthetic as S
from termdaw_b200.state import Project, State
from helpers import run_project, pcm_to_int

pytestmark = pytest.mark.gpu


def _render(lib, cfg, inline, passes=("scan", "render"), reuse=True):
    p, a, s = cfg
    st = State(lib, s["psr"], s["bl"], assets=a)
    try:
        assert st.refresh(p), st.messages
        st.ctx.set_inline_sums(inline)
        st.ctx.set_pass_reuse(reuse)
        if "scan" in passes:
            st.scan_exact()
        pcm = st.render().copy()
        return pcm, st.ctx.last_pass_stats(), st.bd
    finally:
        st.ctx.set_inline_sums(True)
        st.ctx.set_pass_reuse(True)
        st.close()


def sum_chain_project(seconds=1.5):
    """Sums feeding a sum, an adsr and the output normalize; a sum used twice (not inlinable); a one-input sum;
    a sum of a sum (only the inner one is inlined); no band pass anywhere: everything is bit-exact."""
    rng = np.random.default_rng(77)
    a = S._shared_assets(rng, seconds, 2)
    p = Project()
    p.set_length(seconds)
    p.set_render_samplerate(48000)
    p.set_render_bitdepth(24)
    for s in ("s_noise", "s_tone", "s_long"):
        p.load_sample(s, s, "")
    p.load_resource("table", "table")
    for i in range(2):
        p.load_midi_floww(f"hits{i}", f"hits{i}")
        p.load_midi_floww(f"notes{i}", f"notes{i}")
    p.add_sampleloop("l0", 0.7, 20.0, "s_noise")
    p.add_sampleloop("l1", 1.3, -40.0, "s_tone")
    p.add_sample_multi("m0", 0.9, 10.0, "s_long", "hits0", -1)
    p.add_sample_lerp("r0", 0.8, -15.0, "s_tone", "hits1", -1, 40)
    p.add_synth("y0", 0.6, 5.0, "notes0", 0.4, 0.3, S.HIT_ADSR, 1.0, 0.6, S.NOTE_ADSR, 0.5, S.NOTE_ADSR)
    p.add_sampsyn("w0", 0.5, 0.0, "notes1", S.STD_ADSR, "table")
    p.add_sum("s_a", 0.8, 30.0)        # l0 + m0           -> s_c
    p.add_sum("s_b", 1.2, -60.0)       # r0 + y0 + w0      -> env
    p.add_sum("s_c", 0.9, 0.0)         # s_a + l1          -> out
    p.add_sum("s_one", 1.1, 12.0)      # w0                -> out
    p.add_sum("s_two", 0.5, -5.0)      # m0, consumed by s_c AND env: two consumers, keeps its buffer
    p.add_adsr("env", 1.0, 0.0, 0.7, "notes0", True, True, -1, [0.01, 0.1, 0.8, 0.1, 0.2, 0.05])
    p.add_normalize("out", 1.0, 0.0)
    for e in (("l0", "s_a"), ("m0", "s_a"), ("r0", "s_b"), ("y0", "s_b"), ("w0", "s_b"), ("s_a", "s_c"), ("l1", "s_c"),
              ("w0", "s_one"), ("m0", "s_two"), ("s_two", "s_c"), ("s_two", "env"), ("s_b", "env"),
              ("s_c", "out"), ("env", "out"), ("s_one", "out")):
        p.connect(*e)
    p.set_output("out")
    return p, a, {"psr": S.PSR, "bl": S.BL, "flow": "B"}


def check_sum_chain(lib, oracle):
    cfg = sum_chain_project()
    on, st_on, bd = _render(lib, cfg, True)
    off, st_off, _ = _render(lib, cfg, False)
    assert st_off["inlined_vertices"] == 0
    # render pass: s_b -> env and s_one -> out always; s_a -> s_c or s_c -> out (one level deep: never both)
    assert st_on["inlined_vertices"] >= 3, st_on
    assert st_on["inlined_bytes"] > 0
    assert np.array_equal(on, off)
    want = run_project(oracle, *cfg, taps=False)
    assert np.array_equal(pcm_to_int(on, bd), pcm_to_int(want["pcm"], want["bd"]))      # no band pass: byte exact
    # without pass reuse (every vertex evaluated in both passes) and render only (flow A: running maximum)
    on2, st2, _ = _render(lib, cfg, True, reuse=False)
    assert np.array_equal(on2, on) and st2["inlined_vertices"] >= 3
    a_on, _, _ = _render(lib, cfg, True, passes=("render",))
    a_off, _, _ = _render(lib, cfg, False, passes=("render",))
    assert np.array_equal(a_on, a_off)


def check_batch_projects(lib, oracle, indices=range(12), scale=0.05):
    inlined = 0
    for i in indices:
        cfg = S.config_5_project(i, scale)
        on, st_on, bd = _render(lib, cfg, True)
        off, _, _ = _render(lib, cfg, False)
        assert np.array_equal(on, off), i
        inlined += st_on["inlined_vertices"]
        want = run_project(oracle, *cfg, taps=False)
        x, y = pcm_to_int(on, bd), pcm_to_int(want["pcm"], want["bd"])
        assert x.shape == y.shape and np.abs(x - y).max() <= 1, i                        # band-pass cones: the IIR scan's tolerance
    assert inlined > 0


def check_launch_waves(lib, oracle, indices=(0, 3, 5), scale=0.2):
    """One chunk per block == eight chunks per block, for every chunk kernel (6 s projects: 282 chunks)."""
    try:
        for i in indices:
            cfg = S.config_5_project(i, scale)
            lib.set_launch_waves(0)
            many, _, bd = _render(lib, cfg, True)
            lib.set_launch_waves(1 << 31)
            one, _, _ = _render(lib, cfg, True)
            assert np.array_equal(many, one), i
        cfg = sum_chain_project(3.0)
        lib.set_launch_waves(0)
        many, _, bd = _render(lib, cfg, True)
        want = run_project(oracle, *cfg, taps=False)
        assert np.array_equal(pcm_to_int(many, bd), pcm_to_int(want["pcm"], want["bd"]))
    finally:
        lib.set_launch_waves(16)


def check_fused_refresh(lib, oracle, n=7, scale=0.04):
    """tdg_refresh_normalize_render_many (the refresh of a group done while the GPU renders the groups in front) ==
    tdg_rewind_many followed by tdg_normalize_render_many, bit for bit, repeatedly, with and without the floww reload."""
    import bench
    from termdaw_b200 import normalize_render_many, refresh_normalize_render_many
    cfgs = [S.config_5_project(i, scale) for i in range(n)]
    states = []
    for p, a, s in cfgs:
        st = State(lib, s["psr"], s["bl"], assets=a)
        assert st.refresh(p), st.messages
        states.append((st, p, a))
    try:
        ctxs = [st.ctx for st, _, _ in states]
        args = ([st.cs for st, _, _ in states], [48000] * n, [16] * n)
        ctxs[0].set_pipeline_group(3)                          # groups 3 + 3 + 1 (the taper may regroup: any grouping must do)
        bench.rewind(states, reupload_flowws=True)
        want = [x.copy() for x in normalize_render_many(lib, ctxs, *args)]
        k = bench.rewind(states, reupload_flowws="tables")
        tables = (k["t"], k["n"], k["v"], k["counts"], k["per"])
        for fl in (tables, None, tables):
            got = refresh_normalize_render_many(lib, ctxs, fl, *args)
            for a, b in zip(want, got):
                assert np.array_equal(a, b)
        ctxs[0].set_pipeline_group(0)                          # one group: the refresh simply precedes the passes
        got = refresh_normalize_render_many(lib, ctxs, tables, *args)
        for a, b in zip(want, got):
            assert np.array_equal(a, b)
        for (p, a, s), pcm in zip(cfgs, want):
            ref = run_project(oracle, p, a, s, taps=False)
            x, y = pcm_to_int(pcm, 16), pcm_to_int(ref["pcm"], 16)
            assert x.shape == y.shape and int(np.abs(x - y).max()) <= 1
    finally:
        states[0][0].ctx.set_pipeline_group(32)
        for st, _, _ in states:
            st.close()


def random_dag_project(seed, seconds=0.35):
    """A random DAG, NOT a tree: vertices with several consumers, the same producer connected twice to one consumer,
    sums of sums, sums feeding band passes / adsrs / non-output normalizes, dangling (unreachable) vertices.  What the
    planner's liveness and inlining rules must survive."""
    rng = np.random.default_rng(seed)
    a = S._shared_assets(rng, seconds, 2)
    p = Project()
    p.set_length(seconds)
    p.set_render_samplerate(48000)
    p.set_render_bitdepth(int(rng.choice([16, 24])))
    for sm in ("s_noise", "s_tone", "s_long"):
        p.load_sample(sm, sm, "")
    p.load_resource("table", "table")
    for i in range(2):
        p.load_midi_floww(f"hits{i}", f"hits{i}")
        p.load_midi_floww(f"notes{i}", f"notes{i}")
    gens, fx = [], []
    for i in range(int(rng.integers(3, 7))):
        k = int(rng.integers(5)); name = f"g{i}"
        gain, angle = float(rng.uniform(0.3, 1.4)), float(rng.uniform(-80, 80))
        smp = ("s_noise", "s_tone", "s_long")[int(rng.integers(3))]
        if k == 0: p.add_sampleloop(name, gain, angle, smp)
        elif k == 1: p.add_sample_multi(name, gain, angle, smp, f"hits{int(rng.integers(2))}", -1)
        elif k == 2: p.add_sample_lerp(name, gain, angle, smp, f"hits{int(rng.integers(2))}", -1, 40)
        elif k == 3: p.add_synth(name, gain, angle, f"notes{int(rng.integers(2))}", 0.4, 0.3, S.HIT_ADSR, 1.0, 0.6, S.NOTE_ADSR, 0.5, S.NOTE_ADSR)
        else: p.add_sampsyn(name, gain, angle, f"notes{int(rng.integers(2))}", S.STD_ADSR, "table")
        gens.append(name)
    has_bp = False
    for i in range(int(rng.integers(4, 10))):
        k = int(rng.choice([0, 0, 0, 1, 2, 3])); name = f"f{i}"
        gain, angle = float(rng.choice([1.0, 0.7, 1.3])), float(rng.choice([0.0, 25.0, -50.0]))
        if k == 0: p.add_sum(name, gain, angle)
        elif k == 1: p.add_adsr(name, gain, angle, float(rng.choice([1.0, 0.6])), f"notes{int(rng.integers(2))}", True, True, -1, [0.01, 0.1, 0.8, 0.1, 0.2, 0.05])
        elif k == 2: p.add_bandpass(name, gain, angle, 1.0, 1000.0, 200.0, True); has_bp = True
        else: p.add_normalize(name, gain, angle)
        fx.append(name)
    p.add_normalize("out", 1.0, 0.0)
    order = fx + ["out"]
    for j, f in enumerate(fx):                                  # forward edges only: acyclic
        for _ in range(int(rng.choice([1, 1, 1, 2, 3]))):       # several consumers
            t = order[int(rng.integers(j + 1, len(order)))]
            p.connect(f, t)
            if rng.random() < 0.1:
                p.connect(f, t)                                  # the same edge twice
    for g in gens:
        for _ in range(int(rng.choice([1, 1, 2]))):
            p.connect(g, order[int(rng.integers(len(order)))])
    for f in fx:                                                 # every effect gets at least one input
        if not any(b == f for (_, b) in p.edges):
            p.connect(gens[int(rng.integers(len(gens)))], f)
    p.set_output("out")
    return (p, a, {"psr": S.PSR, "bl": S.BL, "flow": "B"}), has_bp


def check_random_dags(lib, oracle, seeds=range(24)):
    inlined = 0
    for seed in seeds:
        cfg, has_bp = random_dag_project(1000 + seed)
        on, st_on, bd = _render(lib, cfg, True)
        off, _, _ = _render(lib, cfg, False)
        assert np.array_equal(on, off), seed
        on_a, _, _ = _render(lib, cfg, True, passes=("render",))
        off_a, _, _ = _render(lib, cfg, False, passes=("render",))
        assert np.array_equal(on_a, off_a), seed
        inlined += st_on["inlined_vertices"]
        want = run_project(oracle, *cfg, taps=False)
        x, y = pcm_to_int(on, bd), pcm_to_int(want["pcm"], want["bd"])
        assert x.shape == y.shape, seed
        if has_bp:
            assert np.abs(x - y).max() <= (1 if bd == 16 else 16), (seed, int(np.abs(x - y).max()))
        else:
            assert np.array_equal(x, y), seed
    assert inlined > 0


def test_random_dags_inlined_equal_not_inlined_equal_oracle(gpulib, oracle):
    check_random_dags(gpulib, oracle)


def adversarial_note_stream(rng, seconds):
    """Few distinct notes (many re-strikes of one note, some while it is still releasing), two of them within the 0.001
    matching tolerance of another, bursts of events in one frame, offs in the frame of their on, doubled ons."""
    notes = rng.uniform(40, 80, 6)
    notes = np.concatenate([notes, notes[:2] + 0.0005])
    ev, t = [], 0.0
    while t < seconds:
        t += float(rng.choice([0.0, 0.001, 0.01, 0.05, 0.2]))
        n = float(notes[int(rng.integers(len(notes)))])
        ev.append((t, n, float(rng.uniform(0.2, 1.0))) if rng.random() < 0.5 else (t, n, 0.0))
        if rng.random() < 0.15:
            ev.append((t, n, 0.0))
        if rng.random() < 0.1:
            ev.append((t, n, float(rng.uniform(0.2, 1.0))))
    ev.sort(key=lambda e: e[0])
    a = np.array(ev, np.float32).reshape(-1, 3)
    return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy()


def check_voice_planner_fuzz(lib, oracle, seeds=range(12)):
    """The voice planner (engine.cu plan_voices: note matching over the side-by-side note values, lazily erased voices,
    the drop heap) against the oracle's Vec / retain bookkeeping (extensions.rs:469-530): every tap and the PCM bit for bit,
    scan + render (voices and clocks carried from one pass into the next)."""
    from termdaw_b200.state import Assets
    for seed in seeds:
        rng = np.random.default_rng(seed)
        seconds = float(rng.choice([0.6, 1.2, 2.5]))
        a = Assets()
        a.blobs["table"] = S.wavetable_blob(seed, 8, 64, 6.0)
        a.flowws["n0"] = adversarial_note_stream(rng, seconds)
        a.flowws["n1"] = adversarial_note_stream(rng, seconds)
        p = Project()
        p.set_length(seconds); p.set_render_samplerate(48000); p.set_render_bitdepth(24)
        p.load_resource("table", "table"); p.load_midi_floww("n0", "n0"); p.load_midi_floww("n1", "n1")
        rel = float(rng.choice([0.0, 0.02, 0.3]))
        adsr = [0.01, 0.1, 0.8, 0.2, 0.2, rel]
        p.add_synth("syn", 0.8, 10.0, "n0", 0.4, 0.3, S.HIT_ADSR, 1.0, 0.6, adsr, 0.5, adsr)
        p.add_sampsyn("wt", 0.7, -20.0, "n1", [0.01, 1.0, 0.3, 1.0, 1.0, rel + 0.05], "table")
        p.add_debug_sine("ds", 0.3, 0.0, "n0")
        p.add_normalize("out", 1.0, 0.0)
        for v in ("syn", "wt", "ds"):
            p.connect(v, "out")
        p.set_output("out")
        cfg = (p, a, {"psr": 48000, "bl": int(rng.choice([256, 1024])), "flow": "B"})
        got, want = run_project(lib, *cfg), run_project(oracle, *cfg)
        assert np.array_equal(got["pcm"], want["pcm"]), seed
        for name, (l, r) in want["taps"].items():
            gl, gr = got["taps"][name]
            assert np.array_equal(gl, l) and np.array_equal(gr, r), (seed, name)


def test_voice_planner_against_adversarial_note_streams(gpulib, oracle):
    check_voice_planner_fuzz(gpulib, oracle)


def test_fused_refresh_equals_rewind_then_render(gpulib, oracle):
    check_fused_refresh(gpulib, oracle)


def test_sum_chain_inlined_is_bit_identical(gpulib, oracle):
    check_sum_chain(gpulib, oracle)


def test_batch_projects_inlined_equal_not_inlined(gpulib, oracle):
    check_batch_projects(gpulib, oracle)


def test_chunks_per_block_do_not_change_a_bit(gpulib, oracle):
    check_launch_waves(gpulib, oracle)
