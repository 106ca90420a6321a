"""The C-ABI library must load without a GPU and export every symbol include/mom5adv.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

from mom5_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "mom5adv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mom5adv_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    from mom5_b200 import build
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    assert lib.mom5adv_version() == 200


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 18
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/mom5adv.h but not exported"
        assert nm in _lib.SYMBOLS, f"{nm} has no ctypes signature in mom5_b200/_lib.py"
    assert sorted(_lib.SYMBOLS) == names


def test_fma_tolerance_build_exports_the_same_symbols():
    """libmom5adv_fma.so (the FMA-contracted tolerance build one GPU test loads in a subprocess) must be as current as the bit-exact
    library: same exported entry points, test hooks included (a stale build shows up here, on the CPU)"""
    from mom5_b200 import build
    path = build.build(fma=True)                      # rebuilds when a source is newer
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    for nm in list(_lib.SYMBOLS) + list(_lib.DEBUG_SYMBOLS):
        assert hasattr(lib, nm), f"{nm} missing from {path}"


def test_no_link_time_dependency_on_nccl_or_torch():
    """NCCL is dlopen'ed on first use so that a host process keeps its own copy (PyTorch bundles one)."""
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "nccl" not in out and "torch" not in out


HOT_KERNELS = ["_Z13k_sweby_z_tmaILi3ELi0ELb0EE", "_Z14k_sweby_xy_tmaILi3ELi0ELb0ELb0EE", "_Z9k_sweby_zILi3ELi0ELb0EE",
               "_Z10k_sweby_xyILi3ELi0ELb0ELb0EE"]


def _sass_by_function(path):
    import re
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    txt = subprocess.run([cuobjdump, "-sass", path], stdout=subprocess.PIPE, text=True).stdout
    out = {}
    for chunk in re.split(r"\n\s+Function : ", txt)[1:]:
        name, body = chunk.split("\n", 1)
        out[name.strip()] = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)
    return out


def test_sass_is_sm100a_fp64_without_fma_contraction(tmp_path):
    """The bit-exact build: sm_100a SASS whose only DFMAs are the ones WRITTEN in the source -- the shared-reciprocal IEEE division
    (mom5adv_internal.cuh: 5 per reciprocal, 2 per quotient, all explicit __fma_rn).  Checked mechanically for the hot kernels:
      * PTX of the --fmad=false build: every FP64 add / sub / mul carries an explicit rounding mode (.rn), which is what forbids
        ptxas to contract it; fused multiply-adds appear only as `fma.rn.f64` (explicit) and `div.rn.f64` (expanded by ptxas);
      * SASS: DFMA count of each hot kernel == its PTX `fma.rn.f64` count + (DFMAs of one IEEE division) x its `div.rn.f64` count,
        i.e. ptxas added none; and the FMA-contracted build (tolerance build) has strictly more.
    Also: the TMA staging really is in the hot kernels (UTMALDG, no LDGSTS) and the LDGSTS kernels really are the per-thread form."""
    import re
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not (os.path.exists(cuobjdump) and os.path.exists(nvcc)):
        pytest.skip("CUDA toolkit not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # DFMAs ptxas spends on one div.rn.f64 (fast path + slow-path subroutine), measured on a two-line probe kernel
    probe = tmp_path / "divprobe.cu"
    probe.write_text("__global__ void k(const double *a, const double *b, double *c) { c[threadIdx.x] = a[threadIdx.x] / b[threadIdx.x]; }\n")
    cub = tmp_path / "divprobe.cubin"
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "--fmad=false", "-cubin", "-o", str(cub), str(probe)], check=True)
    psass = subprocess.run([cuobjdump, "-sass", str(cub)], stdout=subprocess.PIPE, text=True).stdout
    dfma_per_div = len(re.findall(r"\bDFMA\b", psass))
    assert dfma_per_div >= 7
    ptx = tmp_path / "capi.ptx"
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--fmad=false", "-ptx", "-o", str(ptx),
                    os.path.join(root, "mom5_b200", "csrc", "capi.cu")], check=True, stderr=subprocess.DEVNULL)
    text = ptx.read_text()
    sass = _sass_by_function(_lib.LIB_PATH)
    sass_fma = _sass_by_function(_lib.FMA_LIB_PATH) if os.path.exists(_lib.FMA_LIB_PATH) else {}
    # (1) the whole translation unit: no FP64 add / sub / mul without an explicit rounding mode -> nothing ptxas may contract
    bare = re.findall(r"(?<![\w.])(?:add|sub|mul)\.f64\b", text)          # (atom.global.add.f64 of the total_tracer metric is not one)
    assert not bare, f"FP64 add/sub/mul without an explicit rounding mode could be contracted: {bare[:3]}"
    assert len(re.findall(r"\b(?:add|sub|mul)\.rn\.f64\b", text)) > 1000
    total_div = len(re.findall(r"\bdiv\.rn\.f64\b", text))
    for prefix in HOT_KERNELS:
        m = re.search(r"\.entry (" + re.escape(prefix) + r"\w*)\(", text)
        assert m, prefix
        name = m.group(1)
        body = text[m.start():text.index("\n}", m.start())]
        n_fma, n_div = len(re.findall(r"\bfma\.rn\.f64\b", body)), len(re.findall(r"\bdiv\.rn\.f64\b", body))
        assert n_div == 0, f"{name}: a plain division in the hot kernel body (they belong in the out-of-line exact flavour)"
        ops = sass[name]
        n_dfma = sum(1 for o in ops if o.startswith("DFMA"))
        # (2) every DFMA of the kernel is one the source wrote: the explicit __fma_rn of the shared-reciprocal division in the kernel
        # body, plus whatever out-of-line exact-flavour function (plain divisions, expanded by ptxas) ptxas chose to inline
        assert n_fma <= n_dfma <= n_fma + dfma_per_div * total_div, (name, n_dfma, n_fma, dfma_per_div, total_div)
        assert n_fma % 1 == 0 and n_fma >= 5
        assert any(o.startswith("DADD") for o in ops) and any(o.startswith("DMUL") for o in ops)
        if name in sass_fma:
            assert sum(1 for o in sass_fma[name] if o.startswith("DFMA")) > n_dfma, f"{name}: the FMA build contracts nothing?"
        if "_tma" in prefix:
            assert any(o.startswith("UTMALDG") for o in ops), f"{name}: no tensor-map copy (UTMALDG) in the TMA kernel"
            assert not any(o.startswith("LDGSTS") for o in ops), f"{name}: per-thread cp.async left in the TMA kernel"
        else:
            assert any(o.startswith("LDGSTS") for o in ops)


def test_init_without_gpu_or_bad_args_fails_loudly():
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.mom5adv_init(None, 1, None, C.byref(h))
    assert rc == -1 and b"bad arguments" in lib.mom5adv_last_error()
    g = _lib.Grid()
    g.have_obc = 1
    rc = lib.mom5adv_init(C.byref(g), 1, None, C.byref(h))
    assert rc == -5 and b"open boundaries" in lib.mom5adv_last_error()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        g.have_obc = 0
        rc = lib.mom5adv_init(C.byref(g), 1, None, C.byref(h))
        assert rc == -4 and b"no CUDA device" in lib.mom5adv_last_error()   # no silent CPU fallback
