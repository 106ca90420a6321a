"""ctypes binding of libmom5adv.so (include/mom5adv.h).  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmom5adv.so")

dp = C.POINTER(C.c_double)
dpp = C.POINTER(dp)
ip = C.POINTER(C.c_int)


class Grid(C.Structure):
    """struct mom5adv_grid"""
    _fields_ = [("isc", C.c_int), ("iec", C.c_int), ("jsc", C.c_int), ("jec", C.c_int), ("nk", C.c_int),
                ("ni_global", C.c_int), ("nj_global", C.c_int), ("layout_x", C.c_int), ("layout_y", C.c_int),
                ("x_extent", ip), ("y_extent", ip),
                ("cyclic_x", C.c_int), ("cyclic_y", C.c_int), ("tripolar", C.c_int), ("have_obc", C.c_int),
                ("dat", dp), ("datr", dp), ("dxt", dp), ("dyt", dp), ("dxte", dp), ("dyte", dp), ("dxtn", dp), ("dytn", dp),
                ("dzt", dp), ("tmask", dp)]


# every symbol include/mom5adv.h declares: name -> (restype, argtypes)
_v = C.c_void_p
SYMBOLS = {
    "mom5adv_last_error": (C.c_char_p, []),
    "mom5adv_version": (C.c_int, []),
    "mom5adv_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "mom5adv_set_device": (C.c_int, [C.c_int]),
    "mom5adv_comm_unique_id": (C.c_int, [C.c_char_p]),
    "mom5adv_comm_create": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(_v)]),
    "mom5adv_comm_from_nccl": (C.c_int, [_v, C.c_int, C.c_int, C.POINTER(_v)]),
    "mom5adv_comm_destroy": (C.c_int, [_v]),
    "mom5adv_init": (C.c_int, [C.POINTER(Grid), C.c_int, _v, C.POINTER(_v)]),
    "mom5adv_finalize": (C.c_int, [_v]),
    "mom5adv_sweby_all": (C.c_int, [_v, C.c_int, C.c_double, dpp, dpp, dpp, dp, dp, dp, dp, dpp, dpp, dpp, dpp, dpp, dpp]),
    "mom5adv_sweby_all_dev": (C.c_int, [_v, C.c_int, C.c_double, dpp, dpp, dpp, dp, dp, dp, dp, dpp, dpp, dpp, dpp, dpp, dpp, _v]),
    "mom5adv_horz": (C.c_int, [_v, C.c_int, C.c_double, dp, dp, dp, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp]),
    "mom5adv_horz_dev": (C.c_int, [_v, C.c_int, C.c_double, dp, dp, dp, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp, _v]),
    "mom5adv_vert": (C.c_int, [_v, C.c_int, dp, dp, dp, dp, dp, dp, dp]),
    "mom5adv_vert_dev": (C.c_int, [_v, C.c_int, dp, dp, dp, dp, dp, dp, dp, _v]),
    "mom5adv_tracer_update_dev": (C.c_int, [_v, C.c_int, C.c_double, dp, dp, dpp, dpp, dpp, _v]),
    "mom5adv_sweby_all_step_dev": (C.c_int, [_v, C.c_int, C.c_double, dpp, dp, dp, dp, dp, dp, dp, dpp, dpp, dpp, _v]),
    "mom5adv_continuity_dev": (C.c_int, [_v, dp, dp, dp, dp, dp, dp, _v]),
    "mom5adv_set_ppm_limiters": (C.c_int, [_v, C.c_int, C.c_int]),
    "mom5adv_adv_diss_dev": (C.c_int, [_v, C.c_int, C.c_int, C.c_double, C.c_double, dp, dp, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, _v]),
    "mom5adv_adv_diss": (C.c_int, [_v, C.c_int, C.c_int, C.c_double, C.c_double, dp, dp, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp]),
    "mom5adv_flux_int_z_dev": (C.c_int, [_v, dp, dp, _v]),
    "mom5adv_chksum_dev": (C.c_int, [_v, dp, C.c_int, C.POINTER(C.c_int64), _v]),
    "mom5adv_total_tracer_dev": (C.c_int, [_v, dp, dp, dp, _v]),
    "mom5adv_last_timing_ms": (C.c_int, [_v, C.POINTER(C.c_float)]),
    "mom5adv_kernel_launches": (C.c_int64, [_v]),
    "mom5adv_last_transfer_bytes": (C.c_int, [_v, C.POINTER(C.c_int64)]),
}
# test hooks (host-only logic, no GPU needed)
DEBUG_SYMBOLS = {
    "mom5adv_debug_plan": (C.c_int, [C.c_int] * 10 + [ip, C.c_int]),
    "mom5adv_debug_extent": (C.c_int, [C.c_int, C.c_int, C.c_int, ip, ip]),
    "mom5adv_debug_overlap_sets": (C.c_int, [C.c_int, C.c_int, C.c_int, ip]),
}

_lib = None
FMA_LIB_PATH = os.path.join(_HERE, "libmom5adv_fma.so")


def load() -> C.CDLL:
    """dlopen the in-tree library.  Raises if it has not been built (python -m mom5_b200.build).
    MOM5ADV_FMA=1 in the environment selects, for the whole process, the FMA-contracted build
    (python -m mom5_b200.build --fma): not bit-exact, <= 1e-12 relative on the updated tracer.  The two builds
    are never loaded into one process (each carries its own static CUDA runtime and the same kernel symbols)."""
    global _lib
    if _lib is None:
        fma = os.environ.get("MOM5ADV_FMA", "0") == "1"
        path = os.environ.get("MOM5ADV_LIB") or (FMA_LIB_PATH if fma else LIB_PATH)   # MOM5ADV_LIB: explicit build (tuning variants)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `python -m mom5_b200.build{' --fma' if fma else ''}` "
                               "(there is no CPU fallback for the advection path)")
        lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        for name, (res, args) in {**SYMBOLS, **DEBUG_SYMBOLS}.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class Mom5AdvError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().mom5adv_last_error().decode()
        raise Mom5AdvError(f"{what}: error {rc}: {msg}")
