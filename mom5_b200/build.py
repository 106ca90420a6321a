"""Build libmom5adv.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mom5_b200.build [--force] [--fma]
    python -m mom5_b200.build --variant NAME -DFMINB=2 -DFUNROLL=2 ...     # tuning build -> <repo>/variants_NAME.so

--fma builds the FMA-contracted variant libmom5adv_fma.so (not bit-exact; <= 1e-12 relative on T).
--variant builds the bit-exact library with extra -D block-shape macros (ZBX ZMINB ZSTAGES XWARPS XMINB YWARPS YMINB YROWS_MAX
FWARPS FMINB FUNROLL FROWS_MAX, see csrc/) next to the repo root; select it at run time with MOM5ADV_LIB=<path> (the file
travels to the GPU box with the snapshot; *.so is git-ignored).  Prints the registers / spills of the fused pass.
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


def build_variant(name: str, defines) -> str:
    out = os.path.join(ROOT, f"variants_{name}.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--fmad=false", "-Xptxas=-v",
           "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
           "-o", out, os.path.join(CSRC, "capi.cu"), "-ldl"] + list(defines)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout[-4000:])
    lines = r.stdout.splitlines()
    for n, l in enumerate(lines):   # ptxas -v: "Compiling entry function '<mangled>'" / "... bytes stack frame, ..." / "Used N registers"
        if "Compiling entry function" in l and ("k_sweby_xyILi3ELi0ELb0ELb0" in l or "k_sweby_zILi3ELi0ELb0" in l):
            print(l.split("'")[1][:40], "|", lines[n + 2].strip(), "|", lines[n + 3].strip() if n + 3 < len(lines) else "")
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        nm = sys.argv[sys.argv.index("--variant") + 1]
        print(build_variant(nm, [a for a in sys.argv[1:] if a.startswith("-D")]))
        sys.exit(0)
    # default: both flavours (the bit-exact library and the FMA tolerance build used by one test)
    flavours = [True] if "--fma" in sys.argv else [False] if "--no-fma" in sys.argv or "-v" in sys.argv else [False, True]
    for f in flavours:
        print(build(force="--force" in sys.argv, fma=f, verbose="-v" in sys.argv))
