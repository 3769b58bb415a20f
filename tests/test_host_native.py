"""The native (C++) host side above the C ABI: termdaw_b200/csrc/host/ -> `termdaw_b200_render`.

The reference's host is compiled Rust; the C++ mirror keeps its names and error behaviour.  On CPU the
SAME sources are linked against the oracle library (prefix tdo_ — test infrastructure, built into a
temporary directory) so that the whole host chain — project.toml, Lua subset, WAV/blob/floww loading,
refresh() order, scan, render, WAV writer — is checked byte for byte against the Python mirror driving
the same oracle.  On the GPU (`-m gpu`) the product binary must write the file the Python front-end
writes through the CUDA library."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import ORACLE_SO, ROOT, load_oracle, run_project
from termdaw_b200 import synthetic as S
from termdaw_b200.csrc import build as B
from termdaw_b200.lua_subset import run_file
from termdaw_b200.wav import encode_wav

HOST_BIN = os.path.join(ROOT, "termdaw_b200", "csrc", "termdaw_b200_render")


@pytest.fixture(scope="module")
def host_bin():
    if not os.path.exists(HOST_BIN):                         # built in-tree by __graft_entry__.build(); travels with the snapshot
        B.build()
    assert os.path.exists(HOST_BIN)
    return HOST_BIN


@pytest.fixture(scope="module")
def host_oracle_bin(tmp_path_factory):
    load_oracle()                                            # builds oracle/libtermdaw_oracle.so on demand
    out = str(tmp_path_factory.mktemp("host_oracle") / "termdaw_host_oracle")
    return B.build_host(force=True, prefix="tdo_", lib=ORACLE_SO, out=out)


def _project_as_dict(p):
    d = {"length_seconds": p.length_seconds, "render_samplerate": p.render_samplerate,
         "render_bitdepth": p.render_bitdepth, "output_file": p.output_file, "output_vertex": p.output_vertex}
    for k in ("samples", "resources", "midis", "lv2plugins", "lv2params", "sums", "norms", "sampleloops", "samplemultis",
              "samplelerps", "debugsines", "synths", "sampsyns", "lv2fxs", "adsrs", "bandpasses", "edges"):
        d[k] = [list(t) for t in getattr(p, k)]
    d["streams"] = [[s] for s in p.streams]
    return d


def _norm(x):
    """ints and floats compare by value (the Python mirror keeps Lua integers as int)"""
    if isinstance(x, bool) or x is None or isinstance(x, str):
        return x
    if isinstance(x, (int, float)):
        return float(x)
    if isinstance(x, (list, tuple)):
        return [_norm(e) for e in x]
    if isinstance(x, dict):
        return {k: _norm(v) for k, v in x.items()}
    raise TypeError(type(x))


REF_SCRIPTS = ["project.lua", "examples/neg-adsr-env-example.lua", "examples/sample-project.lua",
               "examples/sample-synth-adsr-lv2fx-example.lua", "examples/stream.lua"]


@pytest.mark.parametrize("ref_file", REF_SCRIPTS)
def test_native_lua_reader_matches_the_python_mirror_on_reference_scripts(host_bin, ref_file):
    """Only where the reference checkout is present (this container); skipped on the GPU box."""
    path = os.path.join("/root/reference", ref_file)
    if not os.path.exists(path):
        pytest.skip("reference checkout not present")
    out = subprocess.run([host_bin, "--dump-project", path], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert _norm(json.loads(out.stdout)) == _norm(_project_as_dict(run_file(path)))


def test_native_lua_reader_on_generated_scripts_and_errors(host_bin, tmp_path):
    for i, cfg in enumerate([S.config_1(0.1), S.config_2(0.1), S.config_4(0.001, 64), S.config_5_project(3, 0.1)]):
        path = tmp_path / f"p{i}.lua"
        path.write_text(S.project_to_lua(cfg[0]) + "\n-- loops, locals and arithmetic\nlocal g = 0.5\n"
                        "for i = 1, 3 do add_sum(\"extra\" .. i, g * i, -90 + 30 * i) end\n"
                        "--[[ block\ncomment ]] connect(\"extra1\", \"extra\" .. 1 + 1);\n")
        out = subprocess.run([host_bin, "--dump-project", str(path)], capture_output=True, text=True, timeout=60)
        assert out.returncode == 0, out.stderr
        assert _norm(json.loads(out.stdout)) == _norm(_project_as_dict(run_file(str(path))))
    for bad in ("add_sum(\"a\", 1)", "nosuch(1)", "x = = 2", "for i = 1, 2 do add_sum(\"a\", 1, 0)", "add_sum(\"a\", {}, 0)"):
        path = tmp_path / "bad.lua"
        path.write_text(bad + "\n")
        out = subprocess.run([host_bin, "--dump-project", str(path)], capture_output=True, text=True, timeout=60)
        assert out.returncode == 1 and "lua error: line " in out.stderr, (bad, out.stderr)


def test_native_host_fails_loudly_without_a_device(host_bin, tmp_path):
    """No CPU fallback: in a container without a GPU the context cannot be created."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d = S.write_project_dir(S.config_1(0.05), str(tmp_path / "proj"))
    out = subprocess.run([host_bin, d], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "ctx_create failed" in out.stderr


def _render_dir(binary, cfg, d, extra=()):
    S.write_project_dir(cfg, d)
    out = subprocess.run([binary, d, "-o", "native.wav", *extra], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    with open(os.path.join(d, "native.wav"), "rb") as f:
        wav = f.read()
    norms = {}
    for ln in out.stdout.splitlines():
        if ": " in ln and not ln.startswith("rendered"):
            k, v = ln.rsplit(": ", 1)
            norms[k] = float(v)
    return wav, norms


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "p0", "p5"])
def test_native_host_over_the_oracle_equals_python_over_the_oracle(host_oracle_bin, oracle, tmp_path, name):
    cfg = {"c1": lambda: S.config_1(0.2), "c2": lambda: S.config_2(0.02), "c3": lambda: S.config_3(0.002, 16),
           "p0": lambda: S.config_5_project(0, 0.03), "p5": lambda: S.config_5_project(5, 0.03)}[name]()
    for flow, extra in (("B", ()), ("A", ("--no-normalize",))):
        wav, norms = _render_dir(host_oracle_bin, cfg, str(tmp_path / f"{name}_{flow}"), extra)
        want = run_project(oracle, *cfg, flow=flow, taps=False)
        p = cfg[0]
        sr = p.render_samplerate if p.render_samplerate is not None else 48000
        assert wav == encode_wav(want["pcm"], sr, want["bd"]), (name, flow)
        assert set(norms) == set(want["norm"])
        for k, v in want["norm"].items():
            assert norms[k] == pytest.approx(v, rel=1e-7)


def test_native_wav_decoder_equals_python_decoder(host_bin, tmp_path):
    """hound's `samples::<i32>()` `as f32` for every PCM width, IEEE float, plain and extensible headers, odd chunk
    sizes and foreign chunks: csrc/host/tdh_assets.cpp against termdaw_b200/wav.py, by digest of the f32 bits."""
    import struct
    from termdaw_b200.wav import KSDATAFORMAT_SUBTYPE_FLOAT, decode_wav, encode_wav, pcm_from_float
    rng = np.random.default_rng(5)

    def fnv(lin):
        h = 1469598103934665603
        for b in np.asarray(lin, "<f4").tobytes():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h

    files = []
    for bd in (8, 16, 24, 32):
        for ch in (1, 2):
            n = int(rng.integers(1, 300))
            l, r = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
            pcm = pcm_from_float(l, r, bd) if ch == 2 else pcm_from_float(l, l, bd).reshape(n, 2, bd // 8)[:, 0].reshape(-1)
            files.append(encode_wav(pcm, int(rng.choice([22050, 44100, 48000, 96000])), bd, ch))
    # 32-bit float, plain tag 3 and extensible, with a LIST chunk of odd size in front of the data
    f32 = rng.uniform(-2, 2, 101).astype("<f4").tobytes()
    fmt3 = struct.pack("<HHIIHH", 3, 1, 48000, 48000 * 4, 4, 32)
    odd = b"LIST" + struct.pack("<I", 3) + b"abc\x00"
    files.append(b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt3) + len(odd) + 8 + len(f32)) + b"WAVE" + b"fmt " + struct.pack("<I", len(fmt3)) + fmt3 +
                 odd + b"data" + struct.pack("<I", len(f32)) + f32)
    fmte = struct.pack("<HHIIHHHHI", 0xFFFE, 1, 44100, 44100 * 4, 4, 32, 22, 32, 0x4) + KSDATAFORMAT_SUBTYPE_FLOAT
    files.append(b"RIFF" + struct.pack("<I", 4 + 8 + len(fmte) + 8 + len(f32)) + b"WAVE" + b"fmt " + struct.pack("<I", len(fmte)) + fmte +
                 b"data" + struct.pack("<I", len(f32)) + f32)
    for k, data in enumerate(files):
        path = tmp_path / f"w{k}.wav"
        path.write_bytes(data)
        w = decode_wav(data)
        out = subprocess.run([host_bin, "--dump-wav", str(path)], capture_output=True, text=True, timeout=60)
        assert out.returncode == 0, out.stderr
        ch, rate, bits, n, h = (int(x) for x in out.stdout.split())
        assert (ch, rate, bits, n) == (w.channels, w.sample_rate, w.bits, w.linear.size), k
        assert h == fnv(w.linear), k
    for bad in (b"RIFFxxxxWAVE", b"junk", files[0][:20]):
        path = tmp_path / "bad.wav"
        path.write_bytes(bad)
        assert subprocess.run([host_bin, "--dump-wav", str(path)], capture_output=True, text=True).returncode == 1


def _batch(binary, cfgs, root, extra=()):
    dirs = [S.write_project_dir(c, os.path.join(root, f"proj{i}")) for i, c in enumerate(cfgs)]
    out = subprocess.run([binary, "--batch", *dirs, "-o", "batch.wav", *extra], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    assert f"rendered {len(cfgs)} projects" in out.stdout
    return [open(os.path.join(d, "batch.wav"), "rb").read() for d in dirs]


def test_native_batch_over_the_oracle(host_oracle_bin, oracle, tmp_path):
    """`--batch`: independent project directories through the batched calls == one by one."""
    cfgs = [S.config_5_project(i, 0.03) for i in range(4)] + [S.config_1(0.1)]
    for flow, extra in (("B", ()), ("A", ("--no-normalize",))):
        wavs = _batch(host_oracle_bin, cfgs, str(tmp_path / flow), extra)
        for cfg, wav in zip(cfgs, wavs):
            want = run_project(oracle, *cfg, flow=flow, taps=False)
            sr = cfg[0].render_samplerate if cfg[0].render_samplerate is not None else 48000
            assert wav == encode_wav(want["pcm"], sr, want["bd"])


@pytest.mark.gpu
def test_native_batch_on_the_gpu(host_bin, gpulib, tmp_path):
    """The pipelined batched call (tdg_normalize_render_many) driven by the native host: every project's file is
    the file the single-project path writes."""
    cfgs = [S.config_5_project(i, 0.05) for i in range(12)]
    wavs = _batch(host_bin, cfgs, str(tmp_path / "b"))
    for cfg, wav in zip(cfgs, wavs):
        got = run_project(gpulib, *cfg, taps=False)
        assert wav == encode_wav(got["pcm"], 48000, got["bd"])


def test_native_play_over_the_oracle(host_oracle_bin, oracle, tmp_path):
    """`--play N`: the playback feed in half-second passes == the Python mirror playing chunk by chunk."""
    from termdaw_b200.state import State
    cfg = S.config_5_project(3, 0.05)
    d = S.write_project_dir(cfg, str(tmp_path / "play"))
    for extra in ((), ("--no-normalize",), ("--ahead", "5")):
        out = subprocess.run([host_oracle_bin, d, "--play", "30", "-o", "p.f32", *extra], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr
        got = np.fromfile(os.path.join(d, "p.f32"), dtype="<f4")
        p, a, s = cfg
        st = State(oracle, s["psr"], s["bl"], assets=a)
        assert st.refresh(p)
        if "--no-normalize" not in extra:
            st.scan_exact()
        want = np.concatenate([st.ctx.play(1) for _ in range(30)])
        st.close()
        assert np.array_equal(got, want), extra


def test_native_host_messages_match_the_python_mirror(host_oracle_bin, oracle, tmp_path):
    """refresh() failures: same messages, project not loaded, exit code 1."""
    from termdaw_b200.cli import DiskAssets
    from termdaw_b200.state import State
    p, a, s = S.config_1(0.05)
    p.load_sample("ghost", "no_such_file.wav", "")
    p.connect("nowhere", "sum")
    d = S.write_project_dir((p, a, s), str(tmp_path / "broken"))
    out = subprocess.run([host_oracle_bin, d], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "TermDaw: project not loaded." in out.stderr
    cwd = os.getcwd()
    os.chdir(d)
    try:
        st = State(oracle, s["psr"], s["bl"], assets=DiskAssets())
        assert not st.refresh(run_file("project.lua"))
        for m in st.messages:
            assert m in out.stderr
        st.close()
    finally:
        os.chdir(cwd)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1", "c2", "p3"])
def test_native_host_on_the_gpu_writes_the_python_front_ends_file(host_bin, gpulib, oracle, tmp_path, name):
    cfg = {"c1": lambda: S.config_1(0.5), "c2": lambda: S.config_2(0.05), "p3": lambda: S.config_5_project(3, 0.1)}[name]()
    wav, norms = _render_dir(host_bin, cfg, str(tmp_path / name))
    got = run_project(gpulib, *cfg, taps=False)
    p = cfg[0]
    sr = p.render_samplerate if p.render_samplerate is not None else 48000
    assert wav == encode_wav(got["pcm"], sr, got["bd"])
    for k, v in got["norm"].items():
        assert norms[k] == pytest.approx(v, rel=1e-6)
    # and the python command-line front-end over the same directory
    out = subprocess.run([sys.executable, "-m", "termdaw_b200.cli", str(tmp_path / name), "-o", "python.wav"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env={**os.environ, "PYTHONPATH": ROOT})
    assert out.returncode == 0, out.stderr
    with open(tmp_path / name / "python.wav", "rb") as f:
        assert f.read() == wav
