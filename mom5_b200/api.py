"""Host-side mirror of the reference's operator interface for the tracer-advection path.

``TracerAdvect`` plays the role of the module ``ocean_tracer_advect_mod`` of the reference
(src/mom5/ocean_tracers/ocean_tracer_advect.F90): ``ocean_tracer_advect_init`` (OTA:507) -> constructor,
``horz_advect_tracer`` (OTA:1898), ``vert_advect_tracer`` (OTA:2095), ``advect_tracer_sweby_all`` (OTA:4104),
``ocean_tracer_advect_end`` -> ``close``.  The same names, argument meaning and error behaviour
(invalid scheme -> error, as mpp_error(FATAL) at OTA:1983-1985).

All compute happens in libmom5adv.so (hand-written sm_100a CUDA) through the C ABI of include/mom5adv.h.
Arrays are either torch CUDA tensors (device-resident mode: the ``*_dev`` entry points, asynchronous on the
current torch stream) or numpy / CPU torch arrays (host mode: H2D + kernels + D2H inside the call).
Shapes follow the Fortran memory order: (nk, nj+2, ni+2) C-order == (isd:ied, jsd:jed, nk) column-major.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Grid, check, dp, dpp

ADVECT_UPWIND = 1            # ocean_parameters.F90:149-163
ADVECT_QUICKER = 5
ADVECT_MDPPM = 8
ADVECT_MDFL_SWEBY = 9
ADVECT_DST_LINEAR = 10
ADVECT_MDFL_SWEBY_TEST = 12
ADVECT_DST_LINEAR_TEST = 14
SCHEME_IDS = {"upwind": ADVECT_UPWIND, "quicker": ADVECT_QUICKER, "mdfl_sweby": ADVECT_MDFL_SWEBY,
              "dst_linear": ADVECT_DST_LINEAR, "mdfl_sweby_test": ADVECT_MDFL_SWEBY_TEST,
              "dst_linear_test": ADVECT_DST_LINEAR_TEST, "mdppm": ADVECT_MDPPM}


def _is_torch(a) -> bool:
    return hasattr(a, "data_ptr")


def _on_device(a) -> bool:
    return _is_torch(a) and a.is_cuda


def _ptr(a) -> dp:
    if a is None:
        return dp()
    if _is_torch(a):
        assert a.is_contiguous() and str(a.dtype) == "torch.float64", "expect contiguous float64"
        return C.cast(a.data_ptr(), dp)
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "expect contiguous float64"
    return a.ctypes.data_as(dp)


def _pp(arrs: Optional[Sequence]) -> dpp:
    if arrs is None:
        return dpp()
    return (dp * len(arrs))(*[_ptr(a) for a in arrs])


def _cur_stream() -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@dataclass
class Communicator:
    """NCCL communicator owned by the library (one per rank)."""
    handle: C.c_void_p
    rank: int
    nranks: int

    @staticmethod
    def create_from_torch_distributed() -> "Communicator":
        """Bootstrap through torch.distributed: rank 0 makes the unique id, everybody gets it by broadcast."""
        import torch
        import torch.distributed as dist
        L = _lib.load()
        rank, nranks = dist.get_rank(), dist.get_world_size()
        buf = C.create_string_buffer(128)
        if rank == 0:
            check(L.mom5adv_comm_unique_id(buf), "comm_unique_id")
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().tolist())
        h = C.c_void_p()
        check(L.mom5adv_comm_create(raw, rank, nranks, C.byref(h)), "comm_create")
        return Communicator(h, rank, nranks)

    def destroy(self):
        if self.handle:
            _lib.load().mom5adv_comm_destroy(self.handle)
            self.handle = None


class TracerAdvect:
    """One rank's tracer-advection module state (replaces ocean_tracer_advect_init + mdfl_init + quicker_init)."""

    def __init__(self, block, dec=None, rank: int = 0, ntracers_max: Optional[int] = None,
                 comm: Optional[Communicator] = None, limit_with_upwind: bool = False):
        """``block``: a BlockInputs-like object (grid2d, dzt, tmask, i0,i1,j0,j1, spec) holding HOST arrays."""
        self.L = _lib.load()
        self.block = block
        self.limit_with_upwind = bool(limit_with_upwind)
        s = block.spec
        px, py = (dec.px, dec.py) if dec is not None else (1, 1)
        self.ni, self.nj, self.nk = block.ni, block.nj, block.nk
        self._keep = []

        def host(a):
            a = a.detach().cpu().numpy() if _is_torch(a) else np.asarray(a)
            a = np.ascontiguousarray(a, dtype=np.float64)
            self._keep.append(a)
            return a.ctypes.data_as(dp)

        g2 = block.grid2d
        xe = ye = None
        if dec is not None:
            xe = (C.c_int * px)(*[dec.iend[d] - dec.ibeg[d] + 1 for d in range(px)])
            ye = (C.c_int * py)(*[dec.jend[d] - dec.jbeg[d] + 1 for d in range(py)])
        grid = Grid(block.i0, block.i1, block.j0, block.j1, self.nk, s.ni, s.nj, px, py,
                    xe if xe is not None else _lib.ip(), ye if ye is not None else _lib.ip(),
                    int(s.cyclic_x), int(s.cyclic_y), int(s.tripolar), 0,
                    host(g2["dat"]), host(g2["datr"]), host(g2["dxt"]), host(g2["dyt"]), host(g2["dxte"]),
                    host(g2["dyte"]), host(g2["dxtn"]), host(g2["dytn"]), host(block.dzt), host(block.tmask))
        self.handle = C.c_void_p()
        self.ntracers_max = ntracers_max or max(len(getattr(block, "T", [])), 1)
        check(self.L.mom5adv_init(C.byref(grid), self.ntracers_max, comm.handle if comm else None, C.byref(self.handle)),
              "mom5adv_init")
        self._keep.clear()

    # ---- advect_tracer_sweby_all (OTA:4104) ----
    def advect_tracer_sweby_all(self, T: Sequence, th_tendency: Sequence, adv_tendency: Sequence, uhrho_et, vhrho_nt,
                                wrho_bt, rho_dzt, dtime: float, flux_x=None, flux_y=None, flux_z=None,
                                adv_x=None, adv_y=None, adv_z=None):
        ntr = len(T)
        dev = _on_device(T[0])
        args = [self.handle, ntr, float(dtime), _pp(T), _pp(th_tendency), _pp(adv_tendency), _ptr(uhrho_et), _ptr(vhrho_nt),  # th / adv may be None (host mode)
                _ptr(wrho_bt), _ptr(rho_dzt), _pp(flux_x), _pp(flux_y), _pp(flux_z), _pp(adv_x), _pp(adv_y), _pp(adv_z)]
        if dev:
            check(self.L.mom5adv_sweby_all_dev(*args, _cur_stream()), "mom5adv_sweby_all_dev")
        else:
            check(self.L.mom5adv_sweby_all(*args), "mom5adv_sweby_all")

    # ---- horz_advect_tracer (OTA:1898), one tracer, advect_sweby_all = .false. ----
    def horz_advect_tracer(self, scheme: int, T_taum1, th_tendency, wrk1, uhrho_et, vhrho_nt, dtime: float = 0.0,
                           T_tau=None, tmask_limit=None, wrho_bt=None, rho_dzt=None, flux_x=None, flux_y=None, flux_z=None):
        args = [self.handle, int(scheme), float(dtime), _ptr(T_taum1), _ptr(T_tau), _ptr(tmask_limit),
                int(self.limit_with_upwind), _ptr(uhrho_et), _ptr(vhrho_nt), _ptr(wrho_bt), _ptr(rho_dzt), _ptr(th_tendency),
                _ptr(wrk1), _ptr(flux_x), _ptr(flux_y), _ptr(flux_z)]
        if _on_device(T_taum1):
            check(self.L.mom5adv_horz_dev(*args, _cur_stream()), "mom5adv_horz_dev")
        else:
            check(self.L.mom5adv_horz(*args), "mom5adv_horz")

    # ---- vert_advect_tracer (OTA:2095) ----
    def vert_advect_tracer(self, scheme: int, T_taum1, th_tendency, wrk1, wrho_bt, T_tau=None, tmask_limit=None, flux_z=None):
        args = [self.handle, int(scheme), _ptr(T_taum1), _ptr(T_tau), _ptr(tmask_limit), _ptr(wrho_bt), _ptr(th_tendency),
                _ptr(wrk1), _ptr(flux_z)]
        if _on_device(th_tendency):
            check(self.L.mom5adv_vert_dev(*args, _cur_stream()), "mom5adv_vert_dev")
        else:
            check(self.L.mom5adv_vert(*args), "mom5adv_vert")

    # ---- tracer time update + halo-1 update (ocean_tracer.F90:2341-2350, ocean_model.F90:1903-1911) ----
    def tracer_update(self, T_taum1: Sequence, th_tendency: Sequence, T_taup1: Sequence, rho_dzt_taum1, rho_dztr_taup1,
                      dtime: float):
        check(self.L.mom5adv_tracer_update_dev(self.handle, len(T_taum1), float(dtime), _ptr(rho_dzt_taum1), _ptr(rho_dztr_taup1),
                                               _pp(T_taum1), _pp(th_tendency), _pp(T_taup1), _cur_stream()), "tracer_update_dev")

    # ---- update_advection_only with sweby_all as the operator, time update fused into the x/y pass ----
    def advect_sweby_all_and_update(self, T_taum1: Sequence, T_taup1: Sequence, rho_dzt_taum1, rho_dztr_taup1, uhrho_et, vhrho_nt,
                                    wrho_bt, rho_dzt_tau, dtime: float, th_tendency: Optional[Sequence] = None,
                                    adv_tendency: Optional[Sequence] = None):
        check(self.L.mom5adv_sweby_all_step_dev(self.handle, len(T_taum1), float(dtime), _pp(T_taum1), _ptr(rho_dzt_taum1),
                                                _ptr(rho_dztr_taup1), _ptr(uhrho_et), _ptr(vhrho_nt), _ptr(wrho_bt), _ptr(rho_dzt_tau),
                                                _pp(T_taup1), _pp(th_tendency) if th_tendency is not None else None,
                                                _pp(adv_tendency) if adv_tendency is not None else None, _cur_stream()),
              "sweby_all_step_dev")

    # ---- continuity: wrho_bt from the horizontal transports (ocean_advection_velocity.F90:660-669) ----
    def continuity(self, uhrho_et, vhrho_nt, wrho_bt, rho_dzt_tendency=None, mass_source=None, diverge_t=None):
        check(self.L.mom5adv_continuity_dev(self.handle, _ptr(uhrho_et), _ptr(vhrho_nt), _ptr(rho_dzt_tendency), _ptr(mass_source),
                                            _ptr(wrho_bt), _ptr(diverge_t), _cur_stream()), "continuity_dev")

    def set_ppm_limiters(self, ppm_hlimiter: int = 1, ppm_vlimiter: int = 1):
        """Tracer%ppm_hlimiter / ppm_vlimiter of the tracer the next ADVECT_MDPPM calls advect (1 cw84, 2 ifc, 3 sh)"""
        check(self.L.mom5adv_set_ppm_limiters(self.handle, int(ppm_hlimiter), int(ppm_vlimiter)), "set_ppm_limiters")

    # ---- diagnostics producers: compute_adv_diss (OTA:7547-7712), z-integrated fluxes (OTA:4317-4326) ----
    def adv_diss(self, horz_scheme: int, vert_scheme: int, T_tau, advect_tendency, uhrho_et, vhrho_nt, wrho_bt, rho_dzt_tau,
                 rho_dzt_taup1, dtime: float, adv_diss_out, conversion: float = 1.0, tmask_limit=None, t2_tendency=None):
        args = [self.handle, int(horz_scheme), int(vert_scheme), float(dtime), float(conversion), _ptr(T_tau), _ptr(tmask_limit),
                int(self.limit_with_upwind), _ptr(uhrho_et), _ptr(vhrho_nt), _ptr(wrho_bt), _ptr(rho_dzt_tau), _ptr(rho_dzt_taup1),
                _ptr(advect_tendency), _ptr(adv_diss_out), _ptr(t2_tendency)]
        if _on_device(T_tau):
            check(self.L.mom5adv_adv_diss_dev(*args, _cur_stream()), "adv_diss_dev")
        else:
            check(self.L.mom5adv_adv_diss(*args), "adv_diss")

    def flux_int_z(self, flux3d, out2d):
        check(self.L.mom5adv_flux_int_z_dev(self.handle, _ptr(flux3d), _ptr(out2d), _cur_stream()), "flux_int_z_dev")

    # ---- metrics ----
    def chksum(self, field, masked: bool = False) -> int:
        out = C.c_int64(0)
        check(self.L.mom5adv_chksum_dev(self.handle, _ptr(field), int(masked), C.byref(out), _cur_stream()), "chksum")
        return out.value

    def total_tracer(self, rho_dzt, T) -> float:
        out = C.c_double(0)
        check(self.L.mom5adv_total_tracer_dev(self.handle, _ptr(rho_dzt), _ptr(T), C.byref(out), _cur_stream()), "total_tracer")
        return out.value

    def last_timing_ms(self) -> dict:
        ms = (C.c_float * 5)()
        check(self.L.mom5adv_last_timing_ms(self.handle, ms), "last_timing_ms")
        return dict(z=ms[0], x=ms[1], y=ms[2], halo=ms[3], total=ms[4])

    def kernel_launches(self) -> int:
        return int(self.L.mom5adv_kernel_launches(self.handle))

    def last_transfer_bytes(self):
        """(host->device, device->host) bytes of the last host-pointer call"""
        b = (C.c_int64 * 2)()
        check(self.L.mom5adv_last_transfer_bytes(self.handle, b), "last_transfer_bytes")
        return int(b[0]), int(b[1])

    def close(self):
        """ocean_tracer_advect_end"""
        if self.handle:
            self.L.mom5adv_finalize(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
