"""Multi-GPU paths on real devices: one process per GPU under torchrun, NCCL inside
the library for the time-sharded hand-off, no data-path collective for project sharding.  Needs >= 2
GPUs (`gpurun --gpus 2`); on a one-GPU box the tests skip and the same arithmetic is covered by the
"virtual ranks" tests of test_timeshard.py."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from helpers import ROOT, pcm_to_int, run_project
from termdaw_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _world():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    return 2 if n < 4 else 4


def _torchrun(world, outdir, case):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "nccl_worker.py"), str(outdir), case]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return [json.load(open(os.path.join(outdir, f"info_{k}.json"))) for k in range(world)]


def test_nccl_time_sharded_bit_exact_graph(oracle, tmp_path):
    """2 x sampleloop -> normalize split by time over real ranks: global peak by NCCL max-allreduce, PCM
    bit-identical to the oracle's unsharded render."""
    world = _world()
    infos = _torchrun(world, tmp_path, "timeshard:c1:1.0")
    cfg = S.config_1(1.0)
    cfg[0].set_render_samplerate(48000)
    want = run_project(oracle, *cfg, taps=False)
    pcm = np.concatenate([np.load(tmp_path / f"pcm_{k}.npy") for k in range(world)])
    assert np.array_equal(pcm, want["pcm"])
    for k, inf in enumerate(infos):
        assert inf["collectives"] >= 1
        assert inf["offset"] == sum(np.load(tmp_path / f"pcm_{q}.npy").size for q in range(k))
        for name, v in want["norm"].items():
            assert inf["norm"][name] == v


def test_nccl_time_sharded_resampled_render(oracle, tmp_path):
    """configs[0] at 44.1 kHz split by time over real ranks: the resampler's 512-frame history crosses the
    boundary by ncclAllGather, the f64 index by host replay; every rank writes at the replayed offset."""
    world = _world()
    infos = _torchrun(world, tmp_path, "timeshard:c1r:1.0")
    want = run_project(oracle, *S.config_1(1.0), taps=False)
    pcm = np.concatenate([np.load(tmp_path / f"pcm_{k}.npy") for k in range(world)])
    x, y = pcm_to_int(pcm, infos[0]["bd"]), pcm_to_int(want["pcm"], want["bd"])
    assert x.shape == y.shape and int(np.abs(x - y).max()) <= 1
    for k, inf in enumerate(infos):
        assert inf["collectives"] >= 2          # normalize peak + halo
        assert inf["offset"] == sum(np.load(tmp_path / f"pcm_{q}.npy").size for q in range(k))


@pytest.mark.parametrize("case,lsb", [("c2:0.1", 16), ("c4:0.002", 16)])
def test_nccl_time_sharded_filters_and_voices(oracle, tmp_path, case, lsb):
    """configs[1] and configs[3] split by time: band-pass state crosses ranks as affine maps (all-gather),
    voices / envelopes by event replay, normalize peak by max-allreduce."""
    world = _world()
    name, scale = case.split(":")
    infos = _torchrun(world, tmp_path, f"timeshard:{case}")
    cfg = {"c2": S.config_2, "c4": S.config_4}[name](float(scale))
    want = run_project(oracle, *cfg, taps=False)
    pcm = np.concatenate([np.load(tmp_path / f"pcm_{k}.npy") for k in range(world)])
    a, b = pcm_to_int(pcm, infos[0]["bd"]), pcm_to_int(want["pcm"], want["bd"])
    assert a.shape == b.shape
    assert int(np.abs(a - b).max()) <= lsb
    for inf in infos:
        for k, v in want["norm"].items():
            assert abs(inf["norm"][k] - v) <= 2e-5 * max(1.0, abs(v))


def test_project_sharding_across_ranks(oracle, tmp_path):
    """configs[4] in miniature: 8 projects dealt over the ranks, each rank renders its own with the
    batched call; every project equals the oracle's render."""
    world = _world()
    n, scale = 8, 0.05
    infos = _torchrun(world, tmp_path, f"projects:{n}:{scale}")
    total = 0
    for i in range(n):
        p, a, s = S.config_5_project(i, scale)
        want = run_project(oracle, p, a, s, taps=False)
        got = np.load(tmp_path / f"proj_{i}.npy")
        x, y = pcm_to_int(got, want["bd"]), pcm_to_int(want["pcm"], want["bd"])
        assert x.shape == y.shape and np.abs(x - y).max() <= 1, i
        total += want["cs"] * s["bl"]
    assert all(inf["total_frames"] == total for inf in infos)
