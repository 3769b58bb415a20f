"""TEST INFRASTRUCTURE: build the product's kernels.cu + engine.cu for the CPU with the stand-in CUDA
runtime of this directory (cuda_runtime.h, emu_runtime.cpp) -> libtermdaw_b200_emu.so, same C ABI
(prefix tdg_).  Used only by tests/test_emulated_kernels.py; the product build (termdaw_b200/csrc/build.py)
never sees this directory.

The only source transformation is the launch syntax, which g++ cannot parse:
    K<<<grid, block, smem, stream>>>(args);      ->   emu::launch(grid, block, [=]() { K(args); });
"""
from __future__ import annotations

import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "termdaw_b200", "csrc")


def _match_paren(s: str, i: int) -> int:
    """s[i] == '(' -> index of the matching ')'"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses")


def _split_top(s: str):
    out, depth, cur = [], 0, ""
    for c in s:
        if c in "(<[{":
            depth += 1
        elif c in ")>]}":
            depth -= 1
        if c == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += c
    out.append(cur.strip())
    return out


def rewrite_launches(src: str) -> str:
    # the dispatch macro passes `<<<...>>>(args)` as a macro argument: expand it textually first
    src = re.sub(r"#define TDG_VEC_DISPATCH\(.*\n", "", src)

    def expand(m):
        vec, k, rest = m.group(1), m.group(2), m.group(3)
        return f"do {{ if ({vec}) {k}<true> {rest}; else {k}<false> {rest}; }} while (0);"
    src = re.sub(r"TDG_VEC_DISPATCH\((\w+),\s*(\w+),\s*(<<<.*>>>\(.*\))\);", expand, src)
    out, pos = "", 0
    kern = re.compile(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<")
    while True:
        m = kern.search(src, pos)
        if not m:
            return out + src[pos:]
        end_cfg = src.index(">>>", m.end())
        cfg = _split_top(src[m.end():end_cfg])
        a0 = src.index("(", end_cfg)
        a1 = _match_paren(src, a0)
        semi = src.index(";", a1)
        grid, block = cfg[0], cfg[1]
        out += src[pos:m.start()] + f"emu::launch(dim3({grid}), dim3({block}), [=]() {{ {m.group(1)}{src[a0:a1 + 1]}; }})" + src[a1 + 1:semi + 1]
        pos = semi + 1


def build(out_dir: str, verbose: bool = False) -> str:
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(CSRC, "kernels.cu")) as f:
        k = rewrite_launches(f.read())
    assert "<<<" not in k
    kpath = os.path.join(out_dir, "kernels_emu.cpp")
    with open(kpath, "w") as f:
        f.write(k)
    lib = os.path.join(out_dir, "libtermdaw_b200_emu.so")
    flags = ["-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-w",
             "-D__CUDACC__", "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include"),
             *os.environ.get("TDG_DEFS", "").split()]          # tuning variants (same knob as csrc/build.py)
    try:
        if " fma " in open("/proc/cpuinfo").read():
            flags.append("-mfma")                          # fmaf() stays exact either way; this only makes it fast
    except OSError:
        pass
    cmd = [os.environ.get("CXX", "g++"), *flags, kpath, "-x", "c++", os.path.join(CSRC, "engine.cu"),
           os.path.join(HERE, "emu_runtime.cpp"), "-o", lib, "-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    import sys
    print(build(sys.argv[1] if len(sys.argv) > 1 else "/tmp/tdg_emu", verbose=True))
