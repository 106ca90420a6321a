#!/usr/bin/env python
"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN FORTRAN TEXT.

    python tests/golden/gen_from_reference.py [case ...]

Runs only where /root/reference exists (the build container).  It reads the loop nests of
    /root/reference/src/mom5/ocean_tracers/ocean_tracer_advect.F90
at generation time (nothing is copied into this repo), translates them statement by statement with
tests/golden/f90interp.py and executes them in IEEE binary64 on small synthetic single-domain cases from
mom5_b200.synthetic.  Inputs and outputs are stored together in tests/golden/<case>.npz; the C oracle
(oracle/mom5adv_oracle.c) must reproduce every output array bit for bit (tests/test_oracle_golden.py).

Routines executed from the reference text (line ranges located by their `subroutine`/`function` headers):
    mdfl_init (mask part) .......... tmask_mdfl
    quicker_init (weights part) .... quick_*, curv_*, dxt_quick, tmask_quick
    advect_tracer_sweby_all ........ th_tendency, T_prog%wrk1, all registered diagnostics
    advect_tracer_mdppm + ppm_limit_cw84 / _ifc / _sh ... piecewise-parabolic scheme, halo 4, three limiters
    advect_tracer_mdfl_sweby_test .. mass-weighted variant (tracer, tracer mass and cell mass carried through the sweeps)
    horz_advect_tracer ............. dispatcher arms upwind / quicker / mdfl_sweby / dst_linear / *_test
    vert_advect_tracer ............. dispatcher arms upwind / quicker
    compute_adv_diss ............... advective dissipation diagnostic (scheme applied to the squared tracer)
What is NOT reference text: the halo filler standing in for FMS mpp_update_domains (single domain: cyclic
wrap, folded north edge, walls untouched; CGRID_NE fold fix) -- that restates
src/shared/mpp/include/mpp_domains_define.inc:4865-4885,2535-2549 and is checked separately against FMS's own
known-answer pattern (test_mpp_domains.F90:5628-5634, 3749-3766) in tests/test_halo_kat.py.
"""
from __future__ import annotations

import math
import os
import re
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from f90interp import FArray, FList, Obj, S, fsq, nint, translate_block, translate_routine  # noqa: E402
from mom5_b200.domain import XUPDATE, YUPDATE, Decomposition  # noqa: E402
from mom5_b200.synthetic import make_case  # noqa: E402

OTA = "/root/reference/src/mom5/ocean_tracers/ocean_tracer_advect.F90"
OPARAM = "/root/reference/src/mom5/ocean_core/ocean_parameters.F90"
OAV = "/root/reference/src/mom5/ocean_core/ocean_advection_velocity.F90"
OOP = "/root/reference/src/mom5/ocean_core/ocean_operators.F90"

GOLDEN_CASES = {
    # name -> (synthetic case, overrides)
    "g_tripolar": ("mini_tripolar", dict(ni=20, nj=14, nk=7, ntr=3)),
    "g_walls": ("mini_walls", dict(ni=13, nj=11, nk=6, ntr=2)),
    "g_torus": ("mini_torus", dict(ni=16, nj=12, nk=5, ntr=2)),
    "g_walls_rough": ("mini_walls", dict(ni=18, nj=10, nk=8, ntr=2, cfl=0.9, seed=1234)),
}


def find_routine(src, kind, name):
    first = next(n for n, l in enumerate(src, 1) if re.match(rf"\s*{kind}\s+{name}\b", l, flags=re.I))
    last = next(n for n in range(first, len(src) + 1) if re.match(rf"\s*end\s+{kind}\s+{name}\b", src[n - 1], flags=re.I))
    return first, last


def find_line(src, pattern, start=1):
    return next(n for n in range(start, len(src) + 1) if re.search(pattern, src[n - 1]))


def to_farray(t, lows):
    """torch/numpy (nk?, ny, nx) C-order -> FArray with Fortran lower bounds"""
    a = np.array(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float64, copy=True)
    return FArray(data=a, lo=list(lows))


class Halo:
    """Single-domain stand-in for mpp_update_domains on halo-h scalar fields + the CGRID_NE fold fix."""

    def __init__(self, dec: Decomposition, isc, iec, jsc, jec):
        self.dec, self.ni, self.nj = dec, iec - isc + 1, jec - jsc + 1

    def scalar(self, f: FArray, flags):
        # a section actual argument such as field(:,:,:) arrives rebased to lower bound 1: recover the true
        # bounds from the shape (the domain's halo width), as the dummy argument's declaration would
        ni, nj = self.ni, self.nj
        hx, hy = (f.a.shape[-1] - ni) // 2, (f.a.shape[-2] - nj) // 2
        ilo, ihi, jlo, jhi = 1 - hx, ni + hx, 1 - hy, nj + hy
        src = f.a.copy()
        for j in range(jlo, jhi + 1):
            j_in = 1 <= j <= nj
            for i in range(ilo, ihi + 1):
                i_in = 1 <= i <= ni
                if i_in and j_in:
                    continue
                want = (flags & XUPDATE) if (j_in and not i_in) else (flags & YUPDATE) if (i_in and not j_in) \
                    else ((flags & XUPDATE) and (flags & YUPDATE))
                if not want:
                    continue
                s = self.dec.map_source(i, j)
                if s is None:
                    continue  # solid wall: untouched
                f.a[..., j - jlo, i - ilo] = src[..., s[1] - jlo, s[0] - ilo]

    def cgrid_ne(self, fx: FArray, fy: FArray):
        """mpp_update_domains(flux_x, flux_y, Dom_flux, gridtype=CGRID_NE), halo 1, folded north + cyclic x."""
        ni, nj = self.ni, self.nj
        ilo, jlo = 1 - (fx.a.shape[-1] - ni) // 2, 1 - (fx.a.shape[-2] - nj) // 2
        sx, sy = fx.a.copy(), fy.a.copy()
        # E/W halo columns (cyclic): plain copies for both components
        if self.dec.cyclic_x:
            for f, s in ((fx, sx), (fy, sy)):
                f.a[..., 1 - jlo:nj + 1 - jlo, 0 - ilo] = s[..., 1 - jlo:nj + 1 - jlo, ni - ilo]
                f.a[..., 1 - jlo:nj + 1 - jlo, ni + 1 - ilo] = s[..., 1 - jlo:nj + 1 - jlo, 1 - ilo]
        if self.dec.tripolar:
            middle = (1 + ni) // 2 + 1
            for i in range(middle, ni + 1):  # fold line, NORTH-position component, eastern half
                fy.a[..., nj - jlo, i - ilo] = -sy[..., nj - jlo, ni + 1 - i - ilo]
        # (north halo row j=nj+1 of either component is never read by the divergence loop)


def build_env(gen, b, src):
    s = gen.s
    ni, nj, nk, ntr = s.ni, s.nj, s.nk, len(b.T)
    isc, iec, jsc, jec = 1, ni, 1, nj
    isd, ied, jsd, jed = 0, ni + 1, 0, nj + 1
    dec = s.decomposition(1, 1)
    env = dict(FArray=FArray, S=S, nint=nint, fsq=fsq, min=min, max=max, abs=abs, real=float,
               oneSixth=1. / 6., r12=1. / 12., twoThirds=2. / 3., fourThirds=4. / 3.,   # `real, parameter ::` locals of the PPM routines
               sign=lambda a, b: math.copysign(abs(a), b),   # IEEE processors: the sign BIT of b (also for b = -0.0)
               isc=isc, iec=iec, jsc=jsc, jec=jec, isd=isd, ied=ied, jsd=jsd, jed=jed, nk=nk,
               num_prog_tracers=ntr, XUPDATE=XUPDATE, YUPDATE=YUPDATE, CGRID_NE=2, FATAL=2,
               onesixth=1.0 / 6.0, have_obc=False, async_domain_update=False, limit_with_upwind=False,
               advect_sweby_all=True, zero_tracer_advect_horz=False, zero_tracer_advect_vert=False,
               compute_gyre_overturn_diagnose=False, module_is_initialized=True, index_temp=1, index_salt=2,
               CLOCK_ROUTINE=0)
    # ADVECT_* scheme ids straight from ocean_parameters.F90
    for l in open(OPARAM):
        m = re.match(r"\s*integer, parameter, public :: (ADVECT_\w+)\s*=\s*(\d+)", l)
        if m:
            env[m.group(1)] = int(m.group(2))
    g = b.grid2d
    Grd = Obj(tripolar=bool(s.tripolar), dzt=to_farray(b.dzt, [1]), tmask=to_farray(b.tmask, [isd, jsd, 1]),
              **{k: to_farray(g[k], [isd, jsd]) for k in ("dat", "datr", "dxt", "dyt", "dxte", "dyte", "dxtn", "dytn")})
    env["Grd"] = Grd
    env["Adv_vel"] = Obj(uhrho_et=to_farray(b.uhrho_et, [isd, jsd, 1]), vhrho_nt=to_farray(b.vhrho_nt, [isd, jsd, 1]),
                         wrho_bt=to_farray(b.wrho_bt, [isd, jsd, 0]))
    taum1, tau, taup1 = 1, 2, 3
    env["Time"] = Obj(taum1=taum1, tau=tau, taup1=taup1)
    rho4 = np.stack([b.rho_dzt.numpy() * 0.97, b.rho_dzt.numpy(), b.rho_dzt.numpy() * 1.01])  # (3,nk,ny,nx): only tau is read
    env["Thickness"] = Obj(rho_dzt=FArray(data=rho4.copy(), lo=[isd, jsd, 1, 1]))
    env["Dens"] = Obj()
    tracers = []
    for n in range(ntr):
        f4 = np.stack([b.T[n].numpy(), b.T_tau[n].numpy(), np.zeros_like(b.T[n].numpy())])
        tracers.append(Obj(field=FArray(data=f4.copy(), lo=[isd, jsd, 1, 1]),
                           th_tendency=to_farray(b.th_tendency[n], [isd, jsd, 1]),
                           wrk1=FArray([(isd, ied), (jsd, jed), (1, nk)], fill=-777.0),
                           tmask_limit=to_farray(b.tmask_limit[n], [isd, jsd, 1]),
                           conversion=1.0, complete=(n == ntr - 1), name=f"tr{n + 1}",
                           horz_advect_scheme=0, vert_advect_scheme=0, ppm_hlimiter=1, ppm_vlimiter=1))
    env["T_prog"] = FList(tracers)
    d1 = lambda: FArray([(isd, ied), (jsd, jed), (1, nk)])
    h2 = lambda: FArray([(isc - 2, iec + 2), (jsc - 2, jec + 2), (1, nk)])
    for nm in ("flux_x", "flux_y", "flux_z", "wrk1", "wrk2", "wrk3", "wrk4", "advect_tendency", "neutral_temp_advect",
               "neutral_salt_advect"):
        env[nm] = d1()
    for nm in ("tmask_mdfl", "tracer_mdfl", "mass_mdfl", "tracermass_mdfl", "tmask_quick", "tracer_quick"):
        env[nm] = h2()
    for nm in ("tmask_mdppm", "tracer_mdppm"):     # halo 4 (OTA:1706-1710)
        env[nm] = FArray([(isc - 4, iec + 4), (jsc - 4, jec + 4), (1, nk)])
    env["Dom_mdppm"] = Obj(domain2d="halo")
    env["tracer_mdfl_all"] = FList([Obj(field=h2()) for _ in range(ntr)])
    env["dxt_quick"] = FArray([(isc - 2, iec + 2), (jsc - 2, jec + 2)])
    env["dyt_quick"] = FArray([(isc - 2, iec + 2), (jsc - 2, jec + 2)])
    for nm, c in (("quick_x", 2), ("quick_y", 2), ("curv_xp", 3), ("curv_xn", 3), ("curv_yp", 3), ("curv_yn", 3)):
        env[nm] = FArray([(isd, ied), (jsd, jed), (1, c)])
    for nm, c in (("quick_z", 2), ("curv_zp", 3), ("curv_zn", 3)):
        env[nm] = FArray([(1, nk), (1, c)])
    env["Dom_mdfl"] = env["Dom_quicker"] = env["Dom_flux"] = Obj(domain2d="halo")
    env["Dom"] = Obj(maskmap=None, domain2d="halo")

    halo = Halo(dec, isc, iec, jsc, jec)
    diags = {}
    DIAG_IDS = ["zflux_adv", "advection_z", "xflux_adv", "advection_x", "xflux_adv_int_z", "yflux_adv", "advection_y",
                "yflux_adv_int_z", "sweby_advect", "horz_advect", "vert_advect", "tracer_advection", "psom_advect",
                "tracer_advection_on_nrho", "tracer_adv_diss"]
    OFF = {"psom_advect", "tracer_advection_on_nrho", "tracer_adv_diss"}   # run_case switches tracer_adv_diss on for the arms
    env["__OFF__"] = OFF
    env["index_temp_sq"] = env["index_salt_sq"] = -1
    env["id_tracer2_advection"] = lambda n: -1
    for q, nm in enumerate(DIAG_IDS):
        env["id_" + nm] = (lambda n, q=q, nm=nm: -1 if nm in OFF else (q + 1) * 100 + n)

    def diagnose(Time, Grd_, id_, data, *a, **k):
        nm, n = DIAG_IDS[id_ // 100 - 1], id_ % 100
        diags[f"{nm}.{n}"] = np.array(data.a, copy=True)

    def update_domains(*args, flags=XUPDATE | YUPDATE, complete=None, gridtype=None, **kw):
        fields = [a for a in args if isinstance(a, FArray)]
        if gridtype is not None:
            halo.cgrid_ne(fields[0], fields[1])
        else:
            for f in fields:
                halo.scalar(f, flags)
        return 1

    def start_update(field, dom, flags=XUPDATE | YUPDATE, complete=None, **kw):
        halo.scalar(field, flags)
        return 1

    def fatal(*a, **k):
        raise RuntimeError("mpp_error: " + " ".join(str(x) for x in a))

    noop = lambda *a, **k: 0
    env.update(mpp_clock_begin=noop, mpp_clock_end=noop, mpp_clock_id=noop, set_ocean_domain=noop,
               watermass_diag=noop, gyre_overturn_diagnose=noop,
               store_ocean_obc_tracer_flux=noop, ocean_obc_zero_boundary=noop,
               mpp_update_domains=update_domains, mpp_start_update_domains=start_update,
               mpp_complete_update_domains=noop, diagnose_3d=diagnose, diagnose_2d=diagnose,
               diagnose_3d_rho=noop, mpp_error=fatal)
    # id_clock_* names are referenced as plain variables
    for l in src:
        for m in re.finditer(r"\b(id_clock_\w+)", l):
            env.setdefault(m.group(1), 0)
    return env, diags


def run_case(name):
    base, over = GOLDEN_CASES[name]
    gen = make_case(base, **over)
    b = gen.block(with_tau=True)
    src = open(OTA).read().split("\n")
    env, diags = build_env(gen, b, src)
    s = gen.s
    arrays = [k for k, v in env.items() if isinstance(v, FArray)]

    # Sequence association: advect_tracer_mdppm passes the level-k slab of a halo-4 array to the limiter kernels by its first
    # element, `tracer_mdppm(isc-4,jsc-4,k)`, received as dimension(isc-4:iec+4,jsc-4:jec+4).  The translator has no
    # storage association, so those actual arguments are rewritten as the equivalent array sections.
    src_sa = [l.replace("tracer_mdppm(isc-4,jsc-4,k)", "tracer_mdppm(:,:,k)").replace("dak(isc-4,jsc-4,k)", "dak(:,:,k)") for l in src]

    def load_routine(kind, nm):
        first, last = find_routine(src, kind, nm)
        code, _ = translate_routine(src_sa if nm == "advect_tracer_mdppm" else src, first, last, array_names=arrays)
        if nm == "compute_adv_diss":   # `logical :: use_psom=.false.` -- the translator skips declarations, initialisers included
            head, rest = code.split("\n", 1)
            code = head + "\n    use_psom = False\n" + rest
        exec(compile(code, f"<OTA:{first}-{last} {nm}>", "exec"), env)
        return first, last

    cites = {}
    for kind, nm in (("subroutine", "advect_tracer_sweby_all"), ("function", "advect_tracer_mdfl_sweby"),
                     ("function", "advect_tracer_mdfl_sweby_test"),
                     ("subroutine", "ppm_limit_cw84"), ("subroutine", "ppm_limit_ifc"), ("subroutine", "ppm_limit_sh"),
                     ("function", "advect_tracer_mdppm"),
                     ("function", "horz_advect_tracer_upwind"), ("function", "vert_advect_tracer_upwind"),
                     ("function", "horz_advect_tracer_quicker"), ("function", "vert_advect_tracer_quicker"),
                     ("subroutine", "compute_adv_diss"),
                     ("subroutine", "horz_advect_tracer"), ("subroutine", "vert_advect_tracer")):
        cites[nm] = load_routine(kind, nm)

    # compute_adv_diss re-runs the advection functions on the squared tracer, which overwrites the module-level flux
    # arrays: keep flux_z as it was when vert_advect_tracer sent its diagnostics
    real_adv_diss, snap = env["compute_adv_diss"], {}

    def adv_diss_hook(*a, **k):
        snap["flux_z"] = env["flux_z"].a.copy()
        return real_adv_diss(*a, **k)

    env["compute_adv_diss"] = adv_diss_hook

    out = {}
    # ---- mdfl_init: mask fill + halo update (tmask_mdfl = 0.0 ... mpp_update_domains) ----
    f0, l0 = find_routine(src, "subroutine", "mdfl_init")
    a = find_line(src, r"^\s*tmask_mdfl\s*=\s*0\.0", f0)
    z = find_line(src, r"call mpp_update_domains\(tmask_mdfl", a)
    code = translate_block(src, a, z, array_names=arrays)
    code = "\n".join(l for l in code.split("\n") if not re.match(r"\s*(mass_mdfl|tracermass_mdfl)", l))
    exec(compile(code, f"<OTA:{a}-{z} mdfl_init>", "exec"), env)
    out["tmask_mdfl"] = env["tmask_mdfl"].a.copy()

    # ---- mdppm_init: halo-4 mask (OTA:1714-1726) ----
    f0, l0 = find_routine(src, "subroutine", "mdppm_init")
    a = find_line(src, r"^\s*tmask_mdppm\s*=\s*0\.0", f0)
    z = find_line(src, r"call mpp_update_domains\(tmask_mdppm", a)
    code = translate_block(src, a, z, array_names=arrays)
    code = "\n".join(l for l in code.split("\n") if not re.match(r"\s*(mass_mdppm|tracermass_mdppm)", l))
    exec(compile(code, f"<OTA:{a}-{z} mdppm_init>", "exec"), env)
    out["tmask_mdppm"] = env["tmask_mdppm"].a.copy()

    T_prog = env["T_prog"]
    ntr = len(T_prog.items)
    th0 = [t.th_tendency.a.copy() for t in T_prog.items]

    def reset():
        for n, t in enumerate(T_prog.items):
            t.th_tendency.a[...] = th0[n]
            t.wrk1.a[...] = -777.0
        diags.clear()

    # ---- advect_tracer_sweby_all via the dispatcher (advect_sweby_all=.true., ntracer=1) ----
    t0 = time.time()
    env["advect_sweby_all"] = True
    env["horz_advect_tracer"](env["Time"], env["Adv_vel"], env["Thickness"], env["Dens"], T_prog, T_prog(1), 1, s.dtime)
    for n, t in enumerate(T_prog.items, 1):
        out[f"sweby_all.th_tendency.{n}"] = t.th_tendency.a.copy()
        out[f"sweby_all.wrk1.{n}"] = t.wrk1.a.copy()
        out[f"sweby_all.tm.{n}"] = env["tracer_mdfl_all"](n).field.a.copy()
    for k, v in diags.items():
        out[f"sweby_all.diag.{k}"] = v
    print(f"  sweby_all: {time.time() - t0:.1f}s, {len(diags)} diagnostics")

    # ---- per-tracer dispatcher arms (with the advective-dissipation diagnostic of vert_advect_tracer's tail) ----
    env["advect_sweby_all"] = False
    env["__OFF__"].discard("tracer_adv_diss")
    f0, l0 = find_routine(src, "subroutine", "quicker_init")
    a = find_line(src, r"^\s*quick_x\s*=\s*0\.0", f0)
    z = find_line(src, r"curv_zn\(k,3\)", a) + 1  # through the enddo of the k loop
    code = translate_block(src, a, z, array_names=arrays)
    exec(compile(code, f"<OTA:{a}-{z} quicker_init>", "exec"), env)
    for nm in ("quick_x", "quick_y", "curv_xp", "curv_xn", "curv_yp", "curv_yn", "quick_z", "curv_zp", "curv_zn",
               "dxt_quick", "dyt_quick", "tmask_quick"):
        out[f"quicker_init.{nm}"] = env[nm].a.copy()

    arms = [("upwind", "ADVECT_UPWIND", False), ("quicker", "ADVECT_QUICKER", False), ("quicker_lim", "ADVECT_QUICKER", True),
            ("mdfl_sweby", "ADVECT_MDFL_SWEBY", False), ("dst_linear", "ADVECT_DST_LINEAR", False),
            ("mdfl_sweby_test", "ADVECT_MDFL_SWEBY_TEST", False), ("dst_linear_test", "ADVECT_DST_LINEAR_TEST", False),
            ("mdppm_cw84", "ADVECT_MDPPM", False), ("mdppm_ifc", "ADVECT_MDPPM", False), ("mdppm_sh", "ADVECT_MDPPM", False)]
    for tag, scheme, lim in arms:
        t0 = time.time()
        reset()
        env["limit_with_upwind"] = lim
        n = min(2, ntr) if tag in ("quicker_lim", "dst_linear_test", "mdppm_ifc") else 1
        tr = T_prog(n)
        tr.ppm_hlimiter = tr.ppm_vlimiter = {"mdppm_ifc": 2, "mdppm_sh": 3}.get(tag, 1)
        tr.conversion = 3992.1 if n == 1 else 1.0     # only compute_adv_diss output is kept from the diagnostics of the arms
        tr.horz_advect_scheme = tr.vert_advect_scheme = env[scheme]
        env["horz_advect_tracer"](env["Time"], env["Adv_vel"], env["Thickness"], env["Dens"], T_prog, tr, n, s.dtime)
        out[f"{tag}.horz.wrk1"] = tr.wrk1.a.copy()
        out[f"{tag}.horz.th_tendency"] = tr.th_tendency.a.copy()
        out[f"{tag}.flux_x"] = env["flux_x"].a.copy()
        out[f"{tag}.flux_y"] = env["flux_y"].a.copy()
        if tag.startswith("mdppm"):
            out[f"{tag}.tracer_mdppm"] = env["tracer_mdppm"].a.copy()
        if tag.endswith("_test"):   # the three running fields after the y sweep (compute domain is what matters)
            for nm in ("tracer_mdfl", "tracermass_mdfl", "mass_mdfl"):
                out[f"{tag}.{nm}"] = env[nm].a.copy()
        env["vert_advect_tracer"](env["Time"], env["Adv_vel"], env["Dens"], env["Thickness"], T_prog, tr, n, s.dtime)
        out[f"{tag}.vert.wrk1"] = tr.wrk1.a.copy()
        out[f"{tag}.vert.th_tendency"] = tr.th_tendency.a.copy()
        out[f"{tag}.flux_z"] = snap["flux_z"]
        out[f"{tag}.adv_diss"] = diags[f"tracer_adv_diss.{n}"].copy()          # wrk4 of compute_adv_diss
        out[f"{tag}.adv_diss.t2_tendency"] = env["wrk1"].a.copy()             # advection operator on the squared tracer
        out[f"{tag}.advect_tendency"] = env["advect_tendency"].a.copy()
        out[f"{tag}.tracer"] = np.array(n)
        out[f"{tag}.conversion"] = np.array(tr.conversion)
        print(f"  {tag}: {time.time() - t0:.1f}s")

    # ---- continuity: diverge_t + wrho_bt recurrence (ocean_advection_velocity.F90 C-grid block; BDX_ET / BDY_NT) ----
    osrc = open(OOP).read().split("\n")
    vsrc = open(OAV).read().split("\n")
    env["halo"] = 1
    for nm in ("BDX_ET", "BDY_NT"):
        f0, l0 = find_routine(osrc, "function", nm)
        code, _ = translate_routine(osrc, f0, l0, array_names=arrays)
        exec(compile(code, f"<ocean_operators.F90:{f0}-{l0} {nm}>", "exec"), env)
    ldiv = find_line(vsrc, r"Adv_vel%diverge_t\(:,:,k\) = Grd%tmask\(:,:,k\)\*\(BDX_ET")
    lw0 = find_line(vsrc, r"Adv_vel%wrho_bt\(:,:,0\) = -\(pme", ldiv)
    lw1 = next(n for n in range(lw0, len(vsrc) + 1) if re.match(r"\s*enddo", vsrc[n - 1]))
    rng = np.random.default_rng(11)
    shp2 = tuple(b.grid2d["dat"].shape)
    env["pme"] = to_farray(1e-5 * rng.standard_normal(shp2), [0, 0])
    env["river"] = to_farray(1e-6 * rng.standard_normal(shp2), [0, 0])
    tend = 1e-4 * rng.standard_normal(tuple(b.rho_dzt.shape))
    msrc = 1e-5 * rng.standard_normal(tuple(b.rho_dzt.shape))
    env["Thickness"].rho_dzt_tendency = to_farray(tend, [0, 0, 1])
    env["Thickness"].mass_source = to_farray(msrc, [0, 0, 1])
    env["tmp"] = FArray([(0, s.ni + 1), (0, s.nj + 1)])
    env["Adv_vel"].diverge_t = FArray([(0, s.ni + 1), (0, s.nj + 1), (1, s.nk)])
    env["Adv_vel"].wrho_bt = FArray([(0, s.ni + 1), (0, s.nj + 1), (0, s.nk)])
    snippet = ["do k=1,nk", vsrc[ldiv - 1], "enddo"] + vsrc[lw0 - 1:lw1]
    code = translate_block(snippet, 1, len(snippet), array_names=arrays + ["tmp", "pme", "river"])
    exec(compile(code, f"<ocean_advection_velocity.F90:{ldiv},{lw0}-{lw1} continuity>", "exec"), env)
    out["continuity.diverge_t"] = env["Adv_vel"].diverge_t.a.copy()
    out["continuity.wrho_bt"] = env["Adv_vel"].wrho_bt.a.copy()
    out["continuity.in.tend"] = tend
    out["continuity.in.src"] = msrc
    print(f"  continuity: OAV:{ldiv},{lw0}-{lw1}")

    # ---- inputs ----
    inp = dict(ni=s.ni, nj=s.nj, nk=s.nk, ntr=ntr, dtime=s.dtime, cyclic_x=s.cyclic_x, cyclic_y=s.cyclic_y,
               tripolar=s.tripolar, dzt=b.dzt.numpy(), tmask=b.tmask.numpy(), rho_dzt=b.rho_dzt.numpy(),
               uhrho_et=b.uhrho_et.numpy(), vhrho_nt=b.vhrho_nt.numpy(), wrho_bt=b.wrho_bt.numpy())
    for k, v in b.grid2d.items():
        inp["grid." + k] = v.numpy()
    for n in range(ntr):
        inp[f"T.{n + 1}"] = b.T[n].numpy()
        inp[f"T_tau.{n + 1}"] = b.T_tau[n].numpy()
        inp[f"th0.{n + 1}"] = th0[n]
        inp[f"tmask_limit.{n + 1}"] = b.tmask_limit[n].numpy()
    meta = "; ".join(f"{k}=OTA:{a}-{z}" for k, (a, z) in cites.items())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), __cites__=np.array(meta),
                        **{"in." + k: v for k, v in inp.items()}, **{"out." + k: v for k, v in out.items()})
    print(f"  wrote {name}.npz ({os.path.getsize(os.path.join(HERE, name + '.npz')) / 1e3:.0f} kB)")


if __name__ == "__main__":
    if not os.path.exists(OTA):
        sys.exit("the reference tree is not available here; golden vectors are committed under tests/golden/")
    for nm in (sys.argv[1:] or list(GOLDEN_CASES)):
        print(nm)
        run_case(nm)
