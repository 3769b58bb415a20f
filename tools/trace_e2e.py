"""TDG_TRACE=1 python tools/trace_e2e.py [projects]: the pipelined scan+render call of bench.py's e2e leg with the
library's job timeline on stderr (host plan / submit / device kernels / D2H per group)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, termdaw_b200
from termdaw_b200 import synthetic as S, normalize_render_many
from bench import make_states, rewind
lib = termdaw_b200.load_library()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfgs = [S.config_5_project(i, 1.0) for i in range(N)]
states = make_states(lib, cfgs, 0)
ctxs = [st.ctx for st, _, _ in states]; cs = [st.cs for st, _, _ in states]; rsr = [st.render_sr for st, _, _ in states]; bds = [st.bd for st, _, _ in states]
outs = [torch.empty(int(lib.render_bound_bytes(c.h, n, r, b)), dtype=torch.uint8, pin_memory=True).numpy() for c, n, r, b in zip(ctxs, cs, rsr, bds)]
for spec in os.environ.get("TDG_E2E_GROUPS", os.environ.get("TDG_E2E_GROUP", "32")).split(","):
    grp, _, q = spec.partition(":")                    # "64:0.75" = groups of 64, geometric tail ratio 0.75
    grp = int(grp)
    if q:
        os.environ["TDG_TAIL_RATIO"] = q
    ctxs[0].set_pipeline_group(grp)
    grp = spec
    for it in range(4):
        t0 = time.perf_counter(); rewind(states, True); t1 = time.perf_counter()
        normalize_render_many(lib, ctxs, cs, rsr, bds, want_pcm=True, outs=outs); t3 = time.perf_counter()
        print(f"group {grp} combined it {it}: rewind {1e3*(t1-t0):.1f} scan+render {1e3*(t3-t1):.1f} ms; dev span {ctxs[0].last_pass_stats()['device_ms']:.1f}", file=sys.stderr)
