import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, termdaw_b200
from termdaw_b200 import synthetic as S, normalize_scan_many, render_many, normalize_render_many
from bench import make_states, rewind
lib = termdaw_b200.load_library()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfgs = [S.config_5_project(i, 1.0) for i in range(N)]
states = make_states(lib, cfgs, 0)
ctxs = [st.ctx for st,_,_ in states]; cs=[st.cs for st,_,_ in states]; rsr=[st.render_sr for st,_,_ in states]; bds=[st.bd for st,_,_ in states]
outs = [torch.empty(int(lib.render_bound_bytes(c.h, n, r, b)), dtype=torch.uint8, pin_memory=True).numpy() for c,n,r,b in zip(ctxs,cs,rsr,bds)]
for grp in (0, 64, 64):
    ctxs[0].set_pipeline_group(grp)
    for it in range(3):
        t0=time.perf_counter(); rewind(states, True); t1=time.perf_counter()
        normalize_scan_many(lib, ctxs, cs); t2=time.perf_counter()
        render_many(lib, ctxs, cs, rsr, bds, want_pcm=True, outs=outs); t3=time.perf_counter()
        print(f"group {grp} it {it}: rewind {1e3*(t1-t0):.1f} scan {1e3*(t2-t1):.1f} render {1e3*(t3-t2):.1f} ms; dev scan+render {ctxs[0].last_pass_stats()['device_ms']:.1f}", file=sys.stderr)

ctxs[0].set_pipeline_group(64)
for it in range(4):
    t0=time.perf_counter(); rewind(states, True); t1=time.perf_counter()
    normalize_render_many(lib, ctxs, cs, rsr, bds, want_pcm=True, outs=outs); t3=time.perf_counter()
    print(f"combined it {it}: rewind {1e3*(t1-t0):.1f} scan+render {1e3*(t3-t1):.1f} ms; dev span {ctxs[0].last_pass_stats()['device_ms']:.1f}", file=sys.stderr)
