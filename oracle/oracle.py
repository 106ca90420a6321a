"""ctypes front-end of the CPU oracle (oracle/mom5adv_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (mom5_b200/) never imports this module.

The oracle works on a list of in-process blocks (one per would-be MPI rank) + a layout; halo updates among
the blocks stand in for mpp_update_domains.  Inputs are ``mom5_b200.synthetic.BlockInputs``-shaped objects
(numpy or CPU torch arrays, Fortran memory order).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_mom5adv.so")

XUPDATE, YUPDATE = 1, 2
_dp = C.POINTER(C.c_double)
_dpp = C.POINTER(_dp)


class OrcBlock(C.Structure):
    _fields_ = [("ni", C.c_int), ("nj", C.c_int), ("nk", C.c_int)] + [
        (n, _dp) for n in ("dat", "datr", "dxt", "dyt", "dxte", "dyte", "dxtn", "dytn", "tmask", "dzt",
                           "tmask_h2", "dxt_h2", "dyt_h2", "quick_x", "quick_y", "curv_xp", "curv_xn",
                           "curv_yp", "curv_yn", "quick_z", "curv_zp", "curv_zn")]


class OrcLayout(C.Structure):
    _fields_ = [("ni_g", C.c_int), ("nj_g", C.c_int), ("px", C.c_int), ("py", C.c_int),
                ("ibeg", C.POINTER(C.c_int)), ("iend", C.POINTER(C.c_int)),
                ("jbeg", C.POINTER(C.c_int)), ("jend", C.POINTER(C.c_int)),
                ("cyclic_x", C.c_int), ("cyclic_y", C.c_int), ("fold_north", C.c_int)]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "mom5adv_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "mom5adv_oracle.h"))):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle_mom5adv.so"], check=True, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_chksum.restype = C.c_int64
        _lib.orc_total_tracer.restype = C.c_double
        _lib.orc_compute_extent.restype = C.c_int
    return _lib


def _np(a) -> np.ndarray:
    if a is None:
        return None
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(_dp) if a is not None else _dp()


def _pp(arrs: Sequence[Optional[np.ndarray]]):
    return (_dp * len(arrs))(*[_ptr(a) for a in arrs])


class Block:
    """One rank's static grid data + halo-2 scratch, kept alive for the C struct."""

    def __init__(self, bi):
        self.ni, self.nj, self.nk = bi.ni, bi.nj, bi.nk
        g = {k: _np(v) for k, v in bi.grid2d.items()}
        self.g = g
        self.tmask = _np(bi.tmask)
        self.dzt = _np(bi.dzt)
        ni, nj, nk = self.ni, self.nj, self.nk
        self.tmask_h2 = np.zeros((nk, nj + 4, ni + 4))
        self.dxt_h2 = np.zeros((nj + 4, ni + 4))
        self.dyt_h2 = np.zeros((nj + 4, ni + 4))
        self.quick_x = np.zeros((2, nj + 2, ni + 2)); self.quick_y = np.zeros((2, nj + 2, ni + 2))
        self.curv_xp = np.zeros((3, nj + 2, ni + 2)); self.curv_xn = np.zeros((3, nj + 2, ni + 2))
        self.curv_yp = np.zeros((3, nj + 2, ni + 2)); self.curv_yn = np.zeros((3, nj + 2, ni + 2))
        self.quick_z = np.zeros((2, nk)); self.curv_zp = np.zeros((3, nk)); self.curv_zn = np.zeros((3, nk))
        self.c = OrcBlock(ni, nj, nk, _ptr(g["dat"]), _ptr(g["datr"]), _ptr(g["dxt"]), _ptr(g["dyt"]),
                          _ptr(g["dxte"]), _ptr(g["dyte"]), _ptr(g["dxtn"]), _ptr(g["dytn"]), _ptr(self.tmask),
                          _ptr(self.dzt), _ptr(self.tmask_h2), _ptr(self.dxt_h2), _ptr(self.dyt_h2),
                          _ptr(self.quick_x), _ptr(self.quick_y), _ptr(self.curv_xp), _ptr(self.curv_xn),
                          _ptr(self.curv_yp), _ptr(self.curv_yn), _ptr(self.quick_z), _ptr(self.curv_zp),
                          _ptr(self.curv_zn))

    def h2(self):
        return np.zeros((self.nk, self.nj + 4, self.ni + 4))

    def h4(self):
        return np.zeros((self.nk, self.nj + 8, self.ni + 8))

    def d1(self):
        return np.zeros((self.nk, self.nj + 2, self.ni + 2))


class Oracle:
    """Multi-block CPU oracle for one decomposition.

    ``dec`` is a mom5_b200.domain.Decomposition-like object (ni_g, nj_g, px, py, ibeg/iend/jbeg/jend,
    cyclic_x, cyclic_y, tripolar); ``blocks_in[r]`` the BlockInputs of rank r = ix + px*iy.
    """

    def __init__(self, dec, blocks_in: List):
        self.L = lib()
        self.dec = dec
        self.nb = dec.px * dec.py
        assert len(blocks_in) == self.nb
        self.inp = blocks_in
        self.blocks = [Block(b) for b in blocks_in]
        self._carr = (OrcBlock * self.nb)(*[b.c for b in self.blocks])
        self._ib = (C.c_int * dec.px)(*dec.ibeg); self._ie = (C.c_int * dec.px)(*dec.iend)
        self._jb = (C.c_int * dec.py)(*dec.jbeg); self._je = (C.c_int * dec.py)(*dec.jend)
        self.layout = OrcLayout(dec.ni_g, dec.nj_g, dec.px, dec.py, self._ib, self._ie, self._jb, self._je,
                                int(dec.cyclic_x), int(dec.cyclic_y), int(dec.tripolar))
        self.nk = self.blocks[0].nk
        self._mdfl_ready = False
        self._quick_ready = False
        # numpy views of the dynamic inputs
        self.rho = [_np(b.rho_dzt) for b in blocks_in]
        self.u = [_np(b.uhrho_et) for b in blocks_in]
        self.v = [_np(b.vhrho_nt) for b in blocks_in]
        self.w = [_np(b.wrho_bt) for b in blocks_in]

    # ---- halo update among blocks ----
    def update(self, fields: List[np.ndarray], halo: int, flags: int, nk: Optional[int] = None):
        self.L.orc_update_halo(C.byref(self.layout), _pp(fields), C.c_int(self.nk if nk is None else nk),
                               C.c_int(halo), C.c_int(flags))

    def mdfl_init(self):
        """mdfl_init (OTA:1644-1691): tmask_mdfl with a full halo-2 update."""
        for b in self.blocks:
            self.L.orc_mdfl_init_mask(C.byref(b.c))
        self.update([b.tmask_h2 for b in self.blocks], 2, XUPDATE | YUPDATE)
        self._mdfl_ready = True
        self._quick_ready = False

    def quicker_init(self):
        """quicker_init (OTA:1442-1586)."""
        for b in self.blocks:
            self.L.orc_quicker_init_pre(C.byref(b.c))
        self.update([b.tmask_h2 for b in self.blocks], 2, XUPDATE | YUPDATE)
        for b in self.blocks:
            self.L.orc_quicker_init_edges(C.byref(b.c))
        self.update([b.dxt_h2 for b in self.blocks], 2, XUPDATE | YUPDATE, nk=1)
        self.update([b.dyt_h2 for b in self.blocks], 2, XUPDATE | YUPDATE, nk=1)
        for b in self.blocks:
            self.L.orc_quicker_init_weights(C.byref(b.c))
        self._quick_ready = True
        self._mdfl_ready = True  # tmask_quick == tmask_mdfl

    # ---- advect_tracer_sweby_all ----
    def sweby_all(self, T: List[List[np.ndarray]], th: List[List[np.ndarray]], dtime: float, diag: bool = False):
        """T[b][n], th[b][n] (th updated in place). Returns dict of per-block per-tracer outputs."""
        if not self._mdfl_ready:
            self.mdfl_init()
        nb = self.nb
        ntr = len(T[0])
        T = [[_np(t) for t in Tb] for Tb in T]
        tm = [[b.h2() for _ in range(ntr)] for b in self.blocks]
        adv = [[b.d1() for _ in range(ntr)] for b in self.blocks]
        out = dict(adv=adv, tm=tm)
        names = ("flux_x", "flux_y", "flux_z", "adv_x", "adv_y", "adv_z")
        for nme in names:
            out[nme] = [[b.d1() for _ in range(ntr)] for b in self.blocks] if diag else None

        def dg(nme, ib):
            return _pp(out[nme][ib]) if diag else _dpp()

        dt = C.c_double(dtime)
        for ib, b in enumerate(self.blocks):
            self.L.orc_sweby_all_z(C.byref(b.c), ntr, dt, _pp(T[ib]), _ptr(self.w[ib]), _ptr(self.rho[ib]),
                                   _pp(tm[ib]), dg("flux_z", ib), dg("adv_z", ib))
        for n in range(ntr):
            self.update([tm[ib][n] for ib in range(nb)], 2, XUPDATE)
        for ib, b in enumerate(self.blocks):
            self.L.orc_sweby_all_x(C.byref(b.c), ntr, dt, _pp(T[ib]), _ptr(self.u[ib]), _ptr(self.rho[ib]),
                                   _pp(tm[ib]), dg("flux_x", ib), dg("adv_x", ib))
        for n in range(ntr):
            self.update([tm[ib][n] for ib in range(nb)], 2, YUPDATE)
        for ib, b in enumerate(self.blocks):
            self.L.orc_sweby_all_y(C.byref(b.c), ntr, dt, _pp(T[ib]), _ptr(self.u[ib]), _ptr(self.v[ib]),
                                   _ptr(self.w[ib]), _ptr(self.rho[ib]), _pp(tm[ib]), _pp(th[ib]), _pp(adv[ib]),
                                   dg("flux_y", ib), dg("adv_y", ib))
        return out

    def sweby_all_timed(self, T, th, dtime: float, nthreads: int = 0):
        """Whole call through the C multi-block driver (OpenMP over blocks): the timed CPU baseline.
        Scratch (tm, adv) is allocated once per tracer count and reused, as the reference's module arrays are."""
        if not self._mdfl_ready:
            self.mdfl_init()
        nb, ntr = self.nb, len(T[0])
        if getattr(self, "_timed_ntr", None) != ntr:
            self._tm = [[b.h2() for _ in range(ntr)] for b in self.blocks]
            self._adv = [[b.d1() for _ in range(ntr)] for b in self.blocks]
            self._timed_ntr = ntr
        tm, adv = self._tm, self._adv
        flat = lambda xs: _pp([a for xb in xs for a in xb])
        args = (C.byref(self.layout), self._carr, C.c_int(ntr), C.c_double(dtime), flat(T), _pp(self.u), _pp(self.v),
                _pp(self.w), _pp(self.rho), flat(tm), flat(th), flat(adv), C.c_int(nthreads))
        t0 = time.perf_counter()
        self.L.orc_sweby_all_multiblock(*args)
        self.last_seconds = time.perf_counter() - t0
        return dict(adv=adv, tm=tm)

    # ---- advect_tracer_mdfl_sweby (one tracer) ----
    def mdfl_sweby(self, T: List[np.ndarray], dtime: float, sweby_limiter: float = 1.0):
        if not self._mdfl_ready:
            self.mdfl_init()
        nb = self.nb
        T = [_np(t) for t in T]
        tm = [b.h2() for b in self.blocks]
        fx = [b.d1() for b in self.blocks]; fy = [b.d1() for b in self.blocks]; fz = [b.d1() for b in self.blocks]
        wrk1 = [b.d1() for b in self.blocks]
        dt, sl = C.c_double(dtime), C.c_double(sweby_limiter)
        for ib, b in enumerate(self.blocks):
            self.L.orc_mdfl_sweby_z(C.byref(b.c), dt, sl, _ptr(T[ib]), _ptr(self.w[ib]), _ptr(self.rho[ib]),
                                    _ptr(tm[ib]), _ptr(fz[ib]))
        self.update(tm, 2, XUPDATE)
        for ib, b in enumerate(self.blocks):
            self.L.orc_mdfl_sweby_x(C.byref(b.c), dt, sl, _ptr(T[ib]), _ptr(self.u[ib]), _ptr(self.rho[ib]),
                                    _ptr(tm[ib]), _ptr(fx[ib]))
        self.update(tm, 2, YUPDATE)
        for ib, b in enumerate(self.blocks):
            self.L.orc_mdfl_sweby_y(C.byref(b.c), dt, sl, _ptr(T[ib]), _ptr(self.u[ib]), _ptr(self.v[ib]),
                                    _ptr(self.w[ib]), _ptr(self.rho[ib]), _ptr(tm[ib]), _ptr(fy[ib]), _ptr(wrk1[ib]))
        return dict(wrk1=wrk1, flux_x=fx, flux_y=fy, flux_z=fz, tm=tm)

    def sweby_test(self, T: List[np.ndarray], dtime: float, sweby_limiter: float = 1.0):
        """advect_tracer_mdfl_sweby_test (OTA:3469-3746) behind the dispatcher arm OTA:1970-1975"""
        if not self._mdfl_ready:
            self.mdfl_init()
        T = [_np(t) for t in T]
        tr = [b.h2() for b in self.blocks]; tms = [b.h2() for b in self.blocks]; ms = [b.h2() for b in self.blocks]
        fx = [b.d1() for b in self.blocks]; fy = [b.d1() for b in self.blocks]; fz = [b.d1() for b in self.blocks]
        wrk1 = [b.d1() for b in self.blocks]
        dt, sl = C.c_double(dtime), C.c_double(sweby_limiter)
        for ib, b in enumerate(self.blocks):
            self.L.orc_sweby_test_z(C.byref(b.c), dt, sl, _ptr(T[ib]), _ptr(self.w[ib]), _ptr(self.rho[ib]),
                                    _ptr(tr[ib]), _ptr(tms[ib]), _ptr(ms[ib]), _ptr(fz[ib]))
        for f in (tr, tms, ms):
            self.update(f, 2, XUPDATE)
        for ib, b in enumerate(self.blocks):
            self.L.orc_sweby_test_x(C.byref(b.c), dt, sl, _ptr(self.u[ib]), _ptr(tr[ib]), _ptr(tms[ib]), _ptr(ms[ib]), _ptr(fx[ib]))
        for f in (tr, tms, ms):
            self.update(f, 2, YUPDATE)
        for ib, b in enumerate(self.blocks):
            self.L.orc_sweby_test_y(C.byref(b.c), dt, sl, _ptr(T[ib]), _ptr(self.v[ib]), _ptr(self.rho[ib]),
                                    _ptr(tr[ib]), _ptr(tms[ib]), _ptr(ms[ib]), _ptr(fy[ib]), _ptr(wrk1[ib]))
        return dict(wrk1=wrk1, flux_x=fx, flux_y=fy, flux_z=fz, tracer=tr, tracermass=tms, mass=ms)

    def mdppm_init(self):
        """mdppm_init (OTA:1714-1726): tmask_mdppm with a full halo-4 update"""
        self.m4 = []
        for b in self.blocks:
            m = b.h4()
            m[:, 4:-4, 4:-4] = b.tmask[:, 1:-1, 1:-1]
            self.m4.append(m)
        self.update(self.m4, 4, XUPDATE | YUPDATE)

    def mdppm(self, T: List[np.ndarray], dtime: float, limiter: int = 1):
        """advect_tracer_mdppm (OTA:5990-6494) behind the dispatcher arm OTA:1966-1968"""
        if not hasattr(self, "m4"):
            self.mdppm_init()
        T = [_np(t) for t in T]
        tr = [b.h4() for b in self.blocks]
        fx = [b.d1() for b in self.blocks]; fy = [b.d1() for b in self.blocks]; fz = [b.d1() for b in self.blocks]
        wrk1 = [b.d1() for b in self.blocks]
        dt, lim = C.c_double(dtime), C.c_int(limiter)
        for ib, b in enumerate(self.blocks):
            self.L.orc_mdppm_z(C.byref(b.c), dt, lim, _ptr(T[ib]), _ptr(self.w[ib]), _ptr(self.rho[ib]), _ptr(self.m4[ib]),
                               _ptr(tr[ib]), _ptr(fz[ib]))
        self.update(tr, 4, XUPDATE)
        for ib, b in enumerate(self.blocks):
            self.L.orc_mdppm_x(C.byref(b.c), dt, lim, _ptr(T[ib]), _ptr(self.u[ib]), _ptr(self.rho[ib]), _ptr(self.m4[ib]),
                               _ptr(tr[ib]), _ptr(fx[ib]))
        self.update(tr, 4, YUPDATE)
        for ib, b in enumerate(self.blocks):
            self.L.orc_mdppm_y(C.byref(b.c), dt, lim, _ptr(T[ib]), _ptr(self.u[ib]), _ptr(self.v[ib]), _ptr(self.w[ib]),
                               _ptr(self.rho[ib]), _ptr(self.m4[ib]), _ptr(tr[ib]), _ptr(fy[ib]), _ptr(wrk1[ib]))
        return dict(wrk1=wrk1, flux_x=fx, flux_y=fy, flux_z=fz, tracer=tr)

    def adv_diss(self, scheme: str, T_tau, tmask_limit, limit_with_upwind, advect_tendency, rho_taup1, dtime, conversion):
        """compute_adv_diss (OTA:7547-7712) for one tracer; scheme in upwind / quicker / mdfl_sweby / dst_linear /
        mdfl_sweby_test / dst_linear_test.  Returns dict(diss=wrk4, t2_tendency=wrk1) per block."""
        T_tau = [_np(t) for t in T_tau]
        sq = [b.d1() for b in self.blocks]
        for ib, b in enumerate(self.blocks):
            self.L.orc_square(C.byref(b.c), _ptr(T_tau[ib]), _ptr(sq[ib]))
        zero = [b.d1() for b in self.blocks]
        if scheme == "upwind":
            w2, w3 = self.horz_upwind(sq)["wrk1"], self.vert_upwind(sq)["wrk1"]
        elif scheme == "quicker":
            w2 = self.horz_quicker(sq, sq, tmask_limit, limit_with_upwind)["wrk1"]
            w3 = self.vert_quicker(sq, sq, tmask_limit)["wrk1"]
        elif scheme in ("mdfl_sweby", "dst_linear"):
            w2, w3 = self.mdfl_sweby(sq, dtime, 1.0 if scheme == "mdfl_sweby" else 0.0)["wrk1"], zero
        elif scheme == "dst_linear_test":
            w2, w3 = self.sweby_test(sq, dtime, 0.0)["wrk1"], zero
        elif scheme.startswith("mdppm"):
            w2, w3 = self.mdppm(sq, dtime, {"mdppm_cw84": 1, "mdppm_ifc": 2, "mdppm_sh": 3}[scheme])["wrk1"], zero
        elif scheme == "mdfl_sweby_test":      # no arm in compute_adv_diss's select (OTA:7583-7626): wrk2 stays 0
            w2, w3 = zero, zero
        else:
            raise ValueError(scheme)
        t2 = [b.d1() for b in self.blocks]; diss = [b.d1() for b in self.blocks]
        for ib, b in enumerate(self.blocks):
            self.L.orc_adv_diss_final(C.byref(b.c), C.c_double(dtime), C.c_double(conversion), _ptr(self.rho[ib]),
                                      _ptr(_np(rho_taup1[ib])), _ptr(T_tau[ib]), _ptr(_np(advect_tendency[ib])),
                                      _ptr(w2[ib]), _ptr(w3[ib]), _ptr(t2[ib]), _ptr(diss[ib]))
        return dict(diss=diss, t2_tendency=t2)

    # ---- quicker ----
    def horz_quicker(self, Tm1, Tt, tmask_limit, limit_with_upwind: bool):
        if not self._quick_ready:
            self.quicker_init()
        Tm1 = [_np(t) for t in Tm1]; Tt = [_np(t) for t in Tt]; tl = [_np(t) for t in tmask_limit]
        tq = [b.h2() for b in self.blocks]
        fx = [b.d1() for b in self.blocks]; fy = [b.d1() for b in self.blocks]
        wrk1 = [b.d1() for b in self.blocks]
        for ib, b in enumerate(self.blocks):
            self.L.orc_quicker_prep(C.byref(b.c), _ptr(Tm1[ib]), _ptr(tq[ib]))
        self.update(tq, 2, XUPDATE | YUPDATE)
        for ib, b in enumerate(self.blocks):
            self.L.orc_horz_quicker_flux(C.byref(b.c), _ptr(Tm1[ib]), _ptr(Tt[ib]), _ptr(tq[ib]), _ptr(self.u[ib]),
                                         _ptr(self.v[ib]), _ptr(tl[ib]), C.c_int(int(limit_with_upwind)),
                                         _ptr(fx[ib]), _ptr(fy[ib]))
        if self.dec.tripolar:
            self.L.orc_fold_fix_flux(C.byref(self.layout), _pp(fx), _pp(fy), C.c_int(self.nk))
        for ib, b in enumerate(self.blocks):
            self.L.orc_horz_div(C.byref(b.c), _ptr(fx[ib]), _ptr(fy[ib]), _ptr(wrk1[ib]))
        return dict(wrk1=wrk1, flux_x=fx, flux_y=fy, tq=tq)

    def vert_quicker(self, Tm1, Tt, tmask_limit):
        if not self._quick_ready:
            self.quicker_init()
        Tm1 = [_np(t) for t in Tm1]; Tt = [_np(t) for t in Tt]; tl = [_np(t) for t in tmask_limit]
        fz = [b.d1() for b in self.blocks]; wrk1 = [b.d1() for b in self.blocks]
        for ib, b in enumerate(self.blocks):
            self.L.orc_vert_quicker(C.byref(b.c), _ptr(Tm1[ib]), _ptr(Tt[ib]), _ptr(self.w[ib]), _ptr(tl[ib]),
                                    _ptr(fz[ib]), _ptr(wrk1[ib]))
        return dict(wrk1=wrk1, flux_z=fz)

    # ---- upwind ----
    def horz_upwind(self, T):
        T = [_np(t) for t in T]
        fx = [b.d1() for b in self.blocks]; fy = [b.d1() for b in self.blocks]; wrk1 = [b.d1() for b in self.blocks]
        for ib, b in enumerate(self.blocks):
            self.L.orc_horz_upwind(C.byref(b.c), _ptr(T[ib]), _ptr(self.u[ib]), _ptr(self.v[ib]), _ptr(fx[ib]),
                                   _ptr(fy[ib]), _ptr(wrk1[ib]))
        return dict(wrk1=wrk1, flux_x=fx, flux_y=fy)

    def vert_upwind(self, T):
        T = [_np(t) for t in T]
        fz = [b.d1() for b in self.blocks]; wrk1 = [b.d1() for b in self.blocks]
        for ib, b in enumerate(self.blocks):
            self.L.orc_vert_upwind(C.byref(b.c), _ptr(T[ib]), _ptr(self.w[ib]), _ptr(fz[ib]), _ptr(wrk1[ib]))
        return dict(wrk1=wrk1, flux_z=fz)

    # ---- continuity (one block at a time; no halo exchange involved) ----
    def continuity(self, ib: int, u, v, w, tend=None, src=None):
        """w: (nk+1, ny, nx) with level 0 preset; returns (w, diverge_t)"""
        b = self.blocks[ib]
        u, v = _np(u), _np(v)
        w = np.array(_np(w), copy=True)
        div = b.d1()
        self.L.orc_continuity(C.byref(b.c), _ptr(u), _ptr(v), _ptr(_np(tend)), _ptr(_np(src)), _ptr(w), _ptr(div))
        return w, div

    # ---- metrics ----
    def chksum(self, fields: List[np.ndarray], halo: int = 1, masked: bool = False) -> int:
        s = 0
        for ib, b in enumerate(self.blocks):
            a = _np(fields[ib])
            s += self.L.orc_chksum(_ptr(a), b.ni, b.nj, a.shape[0], halo, _ptr(b.tmask) if masked else _dp())
        s &= (1 << 64) - 1
        return s - (1 << 64) if s >= (1 << 63) else s

    def total_tracer(self, T: List[np.ndarray]) -> float:
        return float(sum(self.L.orc_total_tracer(C.byref(b.c), _ptr(self.rho[ib]), _ptr(_np(T[ib])))
                         for ib, b in enumerate(self.blocks)))


def gather(dec, fields: List[np.ndarray], halo: int = 1) -> np.ndarray:
    """Assemble per-block compute domains into the global (nk, nj_g, ni_g) array."""
    nk = fields[0].shape[0]
    out = np.zeros((nk, dec.nj_g, dec.ni_g))
    for r in range(dec.px * dec.py):
        i0, i1, j0, j1 = dec.extent(r)
        f = fields[r]
        out[:, j0 - 1:j1, i0 - 1:i1] = f[:, halo:halo + (j1 - j0 + 1), halo:halo + (i1 - i0 + 1)]
    return out


def split_blocks(dec, global_block) -> List:
    """Cut every rank's BlockInputs out of the global BlockInputs."""
    return [global_block.sub_block(*dec.extent(r)) for r in range(dec.px * dec.py)]
