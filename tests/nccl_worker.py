"""Worker of tests/test_multigpu.py: one process per GPU (torchrun), real NCCL inside the library.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/nccl_worker.py <outdir> <case>

case "timeshard:<config>:<scale>"  every rank loads the same project and renders its time segment
                                  (scan + render); rank r writes pcm_<r>.npy
case "projects:<n>:<scale>"        rank r renders projects r, r+world, ... of the config-5 batch with
                                  the batched call (no data-path collective); writes proj_<index>.npy
Nothing here touches the oracle: the parent test compares the files with it.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    outdir, case = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import termdaw_b200
    from termdaw_b200 import synthetic as S
    from termdaw_b200.state import State
    lib = termdaw_b200.load_library()
    kind, a1, a2 = case.split(":")
    info = {"rank": rank, "world": world}
    if kind == "timeshard":
        from termdaw_b200.timeshard import TimeShardedRender
        cfg = {"c1": S.config_1, "c1r": S.config_1, "c2": S.config_2, "c4": S.config_4}[a1](float(a2))
        p, a, s = cfg
        if a1 == "c1":
            p.set_render_samplerate(48000)          # project rate; "c1r" keeps the README's 44.1 kHz (resampler halo)
        st = State(lib, s["psr"], s["bl"], device=local, assets=a)
        assert st.refresh(p), st.messages
        ts = TimeShardedRender(st, rank, world, dist)
        col0 = st.ctx.collectives()
        ts.scan_exact()
        pcm = ts.render()
        info.update(collectives=st.ctx.collectives() - col0, bounds=ts.bounds, offset=ts.offset_bytes(),
                    norm=st.normalization_values(), bd=st.bd)
        np.save(os.path.join(outdir, f"pcm_{rank}.npy"), pcm)
        ts.close()
        st.close()
    else:
        from termdaw_b200 import normalize_render_many
        n, scale = int(a1), float(a2)
        mine = list(range(rank, n, world))
        states = []
        for i in mine:
            p, a, s = S.config_5_project(i, scale)
            st = State(lib, s["psr"], s["bl"], device=local, assets=a)
            assert st.refresh(p), st.messages
            states.append(st)
        ctxs = [st.ctx for st in states]
        ctxs[0].set_pipeline_group(2)
        pcms = normalize_render_many(lib, ctxs, [st.cs for st in states], [st.render_sr for st in states],
                                     [st.bd for st in states])
        for i, pcm in zip(mine, pcms):
            np.save(os.path.join(outdir, f"proj_{i}.npy"), pcm)
        # the only cross-rank traffic of project sharding: a frame counter for reporting
        t = torch.tensor([float(sum(st.cs * st.bl for st in states))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        info.update(total_frames=int(t.item()), mine=mine)
        for st in states:
            st.close()
    with open(os.path.join(outdir, f"info_{rank}.json"), "w") as f:
        json.dump(info, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
