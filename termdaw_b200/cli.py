"""`python -m termdaw_b200.cli <project dir> [--no-normalize] [--device N] [-o out.wav]`

Renders a termdaw project directory on the GPU: reads `project.toml` like `Config::read`
(config.rs:10-77: `[settings] main`, `buffer_length` default 1024, `project_samplerate` default
44100), runs the Lua-subset reader over the main script, then does what the manual workflow's
`refresh` → `normalize` → `render` commands do (ui_workflow.rs:114-135, state.rs:473-577) and writes
the WAV.  MIDI flowws are read with termdaw_b200/midi.py; a decoded event stream next to the file
(`<file>.floww.npy` or `<file>.floww.f32`, columns t/note/vel) takes precedence.  The native twin of this
front-end is `termdaw_b200/csrc/termdaw_b200_render` (csrc/host/).
"""
from __future__ import annotations

import argparse
import os
import sys
import tomllib

import numpy as np

from . import load_library
from .lua_subset import run_file
from .state import Assets, State


class DiskAssets(Assets):
    def floww(self, path: str):
        if path in self.flowws:
            return self.flowws[path]
        side = path + ".floww.npy"
        if os.path.exists(side):
            a = np.load(side).astype(np.float32).reshape(-1, 3)
            return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy()
        side = path + ".floww.f32"                       # the side file termdaw_b200_render reads too
        if os.path.exists(side):
            a = np.fromfile(side, dtype="<f4").reshape(-1, 3)
            return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy()
        return super().floww(path)                       # a real .mid file: termdaw_b200/midi.py


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="termdaw_b200")
    ap.add_argument("project_dir")
    ap.add_argument("--no-normalize", action="store_true", help="skip the 'normalize' command (running-max flow)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("-o", "--output", default=None)
    args = ap.parse_args(argv)

    os.chdir(args.project_dir)                               # main.rs:39-45
    with open("project.toml", "rb") as f:
        cfg = tomllib.load(f)
    settings = cfg["settings"]
    bl = int(settings.get("buffer_length", 1024))
    psr = int(settings.get("project_samplerate", 44100))
    project = run_file(settings["main"])

    st = State(load_library(), psr, bl, device=args.device, assets=DiskAssets())
    ok = st.refresh(project)
    for m in st.messages:
        print(m, file=sys.stderr)
    if not ok:
        print("TermDaw: project not loaded.", file=sys.stderr)
        return 1
    if not args.no_normalize:
        st.scan_exact()
    path = st.render_to_file(args.output)
    for name, v in st.normalization_values().items():
        print(f"{name}: {v}")
    print(f"rendered {st.cs * st.bl} project-rate frames -> {path}")
    st.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
