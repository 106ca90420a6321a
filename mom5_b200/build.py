"""Build libmom5adv.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mom5_b200.build [--force] [--fma]

--fma builds the FMA-contracted variant libmom5adv_fma.so (not bit-exact; <= 1e-12 relative on T).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)


def lib_path(fma: bool = False) -> str:
    return os.path.join(HERE, "libmom5adv_fma.so" if fma else "libmom5adv.so")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + [
        os.path.join(ROOT, "include", "mom5adv.h")]


def build(force: bool = False, fma: bool = False, verbose: bool = False) -> str:
    out = lib_path(fma)
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(s) for s in sources()):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "--fmad=true" if fma else "--fmad=false", "-Xcompiler", "-fPIC", "-shared",
           "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
           "-o", out, os.path.join(CSRC, "capi.cu"), "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout)
    if verbose:
        print(r.stdout)
    return out


if __name__ == "__main__":
    # default: both flavours (the bit-exact library and the FMA tolerance build used by one test)
    flavours = [True] if "--fma" in sys.argv else [False] if "--no-fma" in sys.argv or "-v" in sys.argv else [False, True]
    for f in flavours:
        print(build(force="--force" in sys.argv, fma=f, verbose="-v" in sys.argv))
