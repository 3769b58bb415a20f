"""In-tree build of libtermdaw_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["kernels.cu", "engine.cu"]
HEADERS = ["tdg_types.h", "tdg_math.cuh", "kernels.h", "engine.h", os.path.join("..", "..", "include", "termdaw_b200.h")]
OUT = os.environ.get("TDG_OUT", os.path.join(HERE, "libtermdaw_b200.so"))       # tuning variants: TDG_OUT + TDG_DEFS
DEFS = os.environ.get("TDG_DEFS", "").split()
BUILD_DIR = os.path.join(HERE, "build" + ("_" + os.path.basename(OUT).replace(".so", "") if DEFS else ""))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false / -ffp-contract=off: the reference rounds after every f32 operation (DESIGN.md "arithmetic")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(os.path.join(HERE, d)) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    objs = []
    jobs = []
    for src in SOURCES:
        obj = os.path.join(BUILD_DIR, src.replace(".cu", ".o"))
        os.makedirs(os.path.dirname(obj), exist_ok=True)
        objs.append(obj)
        if force or _stale(obj, [src] + HEADERS):
            jobs.append([NVCC, *FLAGS, *DEFS, "-c", os.path.join(HERE, src), "-o", obj])
    if jobs:
        def run(cmd):
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.check_call(cmd)
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(OUT):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT, *objs]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    if not DEFS:
        build_host(force=force, verbose=verbose)
    return OUT


HOST_DIR = os.path.join(HERE, "host")
HOST_SOURCES = ["tdh_lua.cpp", "tdh_assets.cpp", "tdh_state.cpp", "tdh_main.cpp"]
HOST_BIN = os.path.join(HERE, "termdaw_b200_render")
CXX = os.environ.get("CXX", "g++")
HOST_FLAGS = ["-O2", "-std=c++17", "-Wall", "-Wextra", "-ffp-contract=off", "-fno-fast-math"]


def build_host(force: bool = False, verbose: bool = False, prefix: str = "tdg_", lib: str | None = None,
               out: str | None = None) -> str:
    """The native host side above the C ABI (csrc/host/): `termdaw_b200_render`, linked against the CUDA
    library next to it (rpath $ORIGIN).  `prefix`/`lib`/`out` let tests/ link the same sources against the
    CPU oracle (prefix tdo_) to check the host logic where no GPU exists; the product build never does."""
    out = out or HOST_BIN
    lib = lib or OUT
    srcs = [os.path.join(HOST_DIR, s) for s in HOST_SOURCES]
    deps = srcs + [os.path.join(HOST_DIR, "tdh.h"), lib]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    libdir, libname = os.path.dirname(os.path.abspath(lib)), os.path.basename(lib)
    defs = [f"-DTDH_PREFIX={prefix}"] + (["-DTDH_ORACLE"] if prefix != "tdg_" else [])
    cmd = [CXX, *HOST_FLAGS, *defs, *srcs, "-o", out, f"-L{libdir}", f"-l:{libname}",
           "-Wl,-rpath," + ("$ORIGIN" if os.path.dirname(os.path.abspath(out)) == libdir else libdir)]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
