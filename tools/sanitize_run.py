"""A pipelined 41-project batch (ramp group, full groups, tapered tail; scan + render with pass reuse ON, two lanes, the
shared copy / upload streams, keep arenas read by a later pass) for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

plus one time-sharded render over virtual ranks (exchange slots written and read inside one launch sequence) and a
playback call.  Prints a checksum so that the run is known to have produced audio."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import termdaw_b200
from termdaw_b200 import synthetic as S, normalize_render_many, render_segments_local, scan_segments_local
from termdaw_b200.state import State

lib = termdaw_b200.load_library()
scale = float(os.environ.get("SAN_SCALE", "0.02"))
cfgs = [S.config_5_project(i, scale) for i in range(41)]
states = []
for p, a, s in cfgs:
    st = State(lib, s["psr"], s["bl"], assets=a)
    assert st.refresh(p), st.messages
    states.append(st)
ctxs = [st.ctx for st in states]
ctxs[0].set_pipeline_group(16)
crc = 0
for rep in range(2):
    for st in states:
        st.ctx.reset_vertex_state(); st.ctx.floww_set_time(0); st.ctx.reset_normalize()
    pcms = normalize_render_many(lib, ctxs, [st.cs for st in states], [st.render_sr for st in states], [st.bd for st in states])
    for x in pcms:
        crc = zlib.crc32(np.ascontiguousarray(x).tobytes(), crc)
print("batch crc", hex(crc))
play = states[3].ctx.play(3)
print("play", play.shape, float(np.abs(play).max()))
for st in states:
    st.close()
# time-sharded render over 3 virtual ranks (band pass + normalize + voices): config 2, short
p, a, s = S.config_2(0.02)
sts = []
for _ in range(3):
    st = State(lib, s["psr"], s["bl"], assets=a)
    assert st.refresh(p), st.messages
    sts.append(st)
cs = sts[0].cs
b = [(0, cs // 3), (cs // 3, cs // 3), (2 * (cs // 3), cs - 2 * (cs // 3))]
scan_segments_local(lib, [st.ctx for st in sts], cs, b)
out = render_segments_local(lib, [st.ctx for st in sts], cs, b, sts[0].bd, render_sr=sts[0].render_sr)
print("segments", [len(x) for x in out])
for st in sts:
    st.close()
print("sanitize_run done")
