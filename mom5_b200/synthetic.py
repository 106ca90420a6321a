"""Deterministic synthetic inputs for the tracer-advection path.

Every field is a POINTWISE function of the global indices (ig, jg, k) plus a counter-based hash
(SplitMix64 of the linear global index), so any rank can generate exactly its own block -- including its
halo ring, which is evaluated at the halo point's *source* point (cyclic wrap / tripolar fold, see
``Decomposition.map_source``) -- without ever materialising the global array, on CPU or directly in HBM.
A block cut out of a larger generated block is bit-identical to generating that block directly.

Array convention: torch tensors in C order with shape (nk, nj+2, ni+2) == Fortran (isd:ied, jsd:jed, nk)
column-major, float64.  ``wrho_bt`` has shape (nk+1, nj+2, ni+2) == Fortran (…, 0:nk)
(src/mom5/ocean_core/ocean_advection_velocity.F90:342).

The reference's experiment inputs (box1, torus1, om3_core3, …) are network downloads that are not available;
only their shapes are taken from the docs.  Nothing here reads /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from .domain import Decomposition

F64 = torch.float64
I64 = torch.int64

RHO0 = 1035.0
EPSLN = 1.0e-40  # src/shared/constants/constants.F90:123

# field ids for the hash streams
_FID = dict(T=16, rho=1, kmt=2, fu=3, fv=4, th=5, tlimit=6, ttau=7)


def _lsr(z: torch.Tensor, n: int) -> torch.Tensor:
    """logical shift right on int64 bit patterns"""
    return torch.bitwise_right_shift(z, n) & ((1 << (64 - n)) - 1)


def _wrap64(c: int) -> int:
    c &= (1 << 64) - 1
    return c - (1 << 64) if c >= (1 << 63) else c


_C0 = _wrap64(0x9E3779B97F4A7C15)
_C1 = _wrap64(0xBF58476D1CE4E5B9)
_C2 = _wrap64(0x94D049BB133111EB)


def splitmix64_uniform(idx: torch.Tensor, stream: int) -> torch.Tensor:
    """U[0,1) from SplitMix64(idx + (stream+1)*golden); idx int64 tensor. 53-bit mantissa, exact in float64."""
    z = idx + _wrap64((stream + 1) * 0x9E3779B97F4A7C15)
    z = z * 1 + _C0  # one increment of the SplitMix64 state
    z = (z ^ _lsr(z, 30)) * _C1
    z = (z ^ _lsr(z, 27)) * _C2
    z = z ^ _lsr(z, 31)
    return _lsr(z, 11).to(F64) * (1.0 / 9007199254740992.0)


def splitmix64_uniform_py(idx: int, stream: int) -> float:
    """Pure-python twin of splitmix64_uniform (for tests)."""
    M = (1 << 64) - 1
    z = (idx + (stream + 1) * 0x9E3779B97F4A7C15) & M
    z = (z + 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    z = z ^ (z >> 31)
    return (z >> 11) * (1.0 / 9007199254740992.0)


@dataclass
class CaseSpec:
    """One synthetic configuration (global)."""
    name: str
    ni: int
    nj: int
    nk: int
    ntr: int
    cyclic_x: bool = False
    cyclic_y: bool = False
    tripolar: bool = False
    grid: str = "mercator"      # "mercator" | "uniform" | "sector"
    bathy: str = "rough"        # "rough" (~35 % land, partial depth) | "flat"
    flow: str = "modes"         # "modes" (3-D divergent, zero vertical sum per face) | "zonal" (torus)
    dzt_kind: str = "stretched"  # "stretched" 10 -> 370 m | "uniform"
    dtime: float = 3600.0
    cfl: float = 0.5            # target max directional CFL
    seed: int = 20240601
    rho_noise: float = 0.01     # relative cell-to-cell variation of rho_dzt
    flow_scale: Optional[float] = None  # set by calibrate(); None -> computed on first use (needs the global max)

    @property
    def cells(self) -> int:
        return self.ni * self.nj * self.nk

    def decomposition(self, px: int = 1, py: int = 1) -> Decomposition:
        return Decomposition(self.ni, self.nj, px, py, cyclic_x=self.cyclic_x, cyclic_y=self.cyclic_y,
                             tripolar=self.tripolar)


# BASELINE.json configs (shapes from SURVEY.md section 8 / BASELINE.md section 3)
CASES: Dict[str, CaseSpec] = {
    "box1": CaseSpec("box1", 24, 35, 18, 2, grid="sector", bathy="flat", dzt_kind="uniform", dtime=7200.0, seed=20240601 + 1),
    "torus": CaseSpec("torus", 256, 256, 50, 2, cyclic_x=True, cyclic_y=True, grid="uniform", bathy="flat",
                      flow="zonal", dzt_kind="uniform", dtime=10800.0, cfl=0.5, seed=20240601 + 2),
    "gyre": CaseSpec("gyre", 256, 256, 50, 8, grid="sector", bathy="flat", dtime=7200.0, seed=20240601 + 2),
    "global_1deg": CaseSpec("global_1deg", 360, 300, 50, 10, cyclic_x=True, tripolar=True, dtime=3600.0, seed=20240601 + 3),
    "global_025deg": CaseSpec("global_025deg", 1440, 1080, 50, 3, cyclic_x=True, tripolar=True, dtime=1800.0, seed=20240601 + 4),
    "global_01deg": CaseSpec("global_01deg", 3600, 2700, 75, 3, cyclic_x=True, tripolar=True, dtime=600.0, seed=20240601 + 5),
    # small cases for fast parity tests
    "mini_tripolar": CaseSpec("mini_tripolar", 40, 30, 12, 3, cyclic_x=True, tripolar=True, dtime=3600.0, seed=77),
    "mini_walls": CaseSpec("mini_walls", 33, 27, 9, 2, grid="sector", dtime=3600.0, seed=78),
    "mini_torus": CaseSpec("mini_torus", 32, 24, 8, 2, cyclic_x=True, cyclic_y=True, grid="uniform", bathy="flat",
                           flow="zonal", dzt_kind="uniform", dtime=10800.0, seed=79),
}


@dataclass
class BlockInputs:
    """Everything one rank hands to the advection path for its block (halo-1 data domain)."""
    spec: CaseSpec
    i0: int
    i1: int
    j0: int
    j1: int
    grid2d: Dict[str, torch.Tensor]          # dat datr dxt dyt dxte dyte dxtn dytn : (nj+2, ni+2)
    dzt: torch.Tensor                        # (nk,)
    tmask: torch.Tensor                      # (nk, nj+2, ni+2)
    rho_dzt: torch.Tensor
    uhrho_et: torch.Tensor
    vhrho_nt: torch.Tensor
    wrho_bt: torch.Tensor                    # (nk+1, nj+2, ni+2)
    T: List[torch.Tensor] = field(default_factory=list)            # taum1 fields
    T_tau: List[torch.Tensor] = field(default_factory=list)        # tau fields (three-level quicker)
    th_tendency: List[torch.Tensor] = field(default_factory=list)
    tmask_limit: List[torch.Tensor] = field(default_factory=list)

    @property
    def ni(self) -> int:
        return self.i1 - self.i0 + 1

    @property
    def nj(self) -> int:
        return self.j1 - self.j0 + 1

    @property
    def nk(self) -> int:
        return self.spec.nk

    def sub_block(self, i0: int, i1: int, j0: int, j1: int) -> "BlockInputs":
        """Cut the block (global 1-based, inclusive) with its halo-1 ring out of this (larger) block."""
        a, b = i0 - self.i0, i1 - self.i0 + 3
        c, d = j0 - self.j0, j1 - self.j0 + 3

        def cut(t):
            return t[..., c:d, a:b].contiguous()

        return BlockInputs(self.spec, i0, i1, j0, j1, {k: cut(v) for k, v in self.grid2d.items()}, self.dzt,
                           cut(self.tmask), cut(self.rho_dzt), cut(self.uhrho_et), cut(self.vhrho_nt), cut(self.wrho_bt),
                           [cut(t) for t in self.T], [cut(t) for t in self.T_tau], [cut(t) for t in self.th_tendency],
                           [cut(t) for t in self.tmask_limit])


class Generator:
    """Pointwise field evaluators for one CaseSpec."""

    def __init__(self, spec: CaseSpec, device="cpu"):
        self.s = spec
        self.dev = torch.device(device)
        self.dec = spec.decomposition()
        s = spec
        k = torch.arange(1, s.nk + 1, dtype=F64)
        if s.dzt_kind == "uniform":
            dzt = torch.full((s.nk,), 4000.0 / s.nk, dtype=F64)
        else:
            dzt = 10.0 + 360.0 * ((k - 1.0) / max(s.nk - 1, 1)) ** 2
        self.dzt = dzt.to(self.dev)
        self.zt = (torch.cumsum(dzt, 0) - 0.5 * dzt).to(self.dev)
        # vertical profiles with zero sum over the top K levels: P[K, k-1], K = 0..nk
        P = torch.zeros((s.nk + 1, s.nk), dtype=F64)
        for K in range(1, s.nk + 1):
            kk = torch.arange(1, K + 1, dtype=F64)
            prof = torch.cos(math.pi * (kk - 0.5) / K) + 0.5 * torch.cos(3.0 * math.pi * (kk - 0.5) / K)
            w = dzt[:K]
            mu = (w * prof).sum() / w.sum()
            P[K, :K] = w * (prof - mu)
        self.P = P.to(self.dev)

    # ---- index plumbing -------------------------------------------------------------------------
    def _src(self, ig: torch.Tensor, jg: torch.Tensor):
        """map (possibly out-of-range) global indices to source indices; returns (ig, jg, valid)."""
        s = self.s
        ig, jg = torch.broadcast_tensors(ig, jg)
        ig, jg = ig.clone(), jg.clone()
        valid = torch.ones_like(ig, dtype=torch.bool)
        north = jg > s.nj
        if s.tripolar:
            ig = torch.where(north, s.ni + 1 - ig, ig)
            jg = torch.where(north, 2 * s.nj + 1 - jg, jg)
        elif s.cyclic_y:
            jg = torch.where(north, jg - s.nj, jg)
        else:
            valid &= ~north
        south = jg < 1
        if s.cyclic_y:
            jg = torch.where(south, jg + s.nj, jg)
        else:
            valid &= ~south
        if s.cyclic_x:
            ig = torch.where(ig < 1, ig + s.ni, ig)
            ig = torch.where(ig > s.ni, ig - s.ni, ig)
        valid &= (ig >= 1) & (ig <= s.ni) & (jg >= 1) & (jg <= s.nj)
        return ig.clamp(1, s.ni), jg.clamp(1, s.nj), valid

    def _noise(self, fid: int, ig, jg, k=None):
        s = self.s
        idx = (jg - 1) * s.ni + (ig - 1)
        if k is not None:
            idx = idx + (k - 1) * (s.ni * s.nj)
        return splitmix64_uniform(idx, s.seed * 64 + fid)

    # ---- 2-D metrics (evaluated at source points; walls clamp = edge replication) ----------------
    def _dxt_dyt(self, ig, jg):
        s = self.s
        ig, jg, _ = self._src(ig, jg)
        x = ig.to(F64)
        y = jg.to(F64)
        if s.grid == "uniform":
            one = torch.ones_like(x)
            return 1.0e5 * one, 1.0e5 * one
        R = 6.371e6
        if s.grid == "sector":  # 2 x 2 degree sector, 10N .. (box1: doc/testcase_ocean_solo.pdf section 5)
            dlon = dlat = math.radians(2.0)
            lat = math.radians(10.0) + (y - 0.5) * dlat
            lat = lat.clamp(max=math.radians(80.0))
            return R * torch.cos(lat) * dlon * (1.0 + 0.0 * x), R * dlat * (1.0 + 0.02 * torch.sin(2 * math.pi * x / s.ni))
        # mercator-like global: -78 .. 88, dx ~ cos(lat), mild i-dependence (tripolar distortion stand-in)
        dlon = 2 * math.pi / s.ni
        lat = math.radians(-78.0) + (y - 0.5) * (math.radians(166.0) / s.nj)
        dlat = math.radians(166.0) / s.nj
        cosl = torch.cos(lat).clamp(min=0.05)
        sym = torch.cos(2 * math.pi * (x - 0.5) / s.ni)  # symmetric under i -> ni+1-i, as the fold requires
        return R * cosl * dlon * (1.0 + 0.05 * sym), R * dlat * (1.0 + 0.05 * sym * torch.sin(lat) ** 2)

    def metrics(self, ig, jg) -> Dict[str, torch.Tensor]:
        dxt, dyt = self._dxt_dyt(ig, jg)
        dxt_e, dyt_e = self._dxt_dyt(ig + 1, jg)
        dxt_n, dyt_n = self._dxt_dyt(ig, jg + 1)
        dat = dxt * dyt                                   # ocean_grids.F90:844
        return dict(dxt=dxt, dyt=dyt, dat=dat, datr=1.0 / (dat + EPSLN),  # ocean_grids.F90:962
                    dxte=0.5 * (dxt + dxt_e), dyte=0.5 * (dyt + dyt_e),
                    dxtn=0.5 * (dxt + dxt_n), dytn=0.5 * (dyt + dyt_n))

    # ---- bathymetry ------------------------------------------------------------------------------
    def kmt(self, ig, jg) -> torch.Tensor:
        """number of wet levels of the column; 0 = land; 0 outside solid walls."""
        s = self.s
        ig, jg, valid = self._src(ig, jg)
        if s.bathy == "flat":
            return torch.where(valid, torch.full_like(ig, s.nk), torch.zeros_like(ig))
        x = 2 * math.pi * (ig.to(F64) - 0.5) / s.ni
        y = math.pi * (jg.to(F64) - 0.5) / s.nj
        # symmetric in i -> ni+1-i near the northern edge is NOT required (the fold maps T cells one-to-one)
        h = (0.55 + 0.30 * torch.sin(3 * x + 0.7) * torch.sin(2 * y + 0.3) + 0.22 * torch.cos(5 * x - 1.1) * torch.cos(3 * y)
             + 0.10 * (self._noise(_FID["kmt"], ig, jg) - 0.5))
        land = h < 0.42
        lev = torch.round(s.nk * (0.25 + 1.4 * (h - 0.42))).to(I64).clamp(2, s.nk)
        return torch.where(valid & ~land, lev, torch.zeros_like(lev))

    # ---- 3-D fields ------------------------------------------------------------------------------
    def rho_dzt(self, ig, jg, k):
        igs, jgs, _ = self._src(ig, jg)
        return RHO0 * self.dzt[k - 1] * (1.0 + self.s.rho_noise * (2.0 * self._noise(_FID["rho"], igs, jgs, k) - 1.0))

    def _face_amp(self, fid, ig, jg, a, b, ph):
        s = self.s
        x = 2 * math.pi * (ig.to(F64)) / s.ni
        y = 2 * math.pi * (jg.to(F64)) / s.nj
        return (torch.sin(a * x + b * y + ph) + 0.35 * torch.cos((a + 2) * x - (b + 1) * y + 2 * ph)
                + 0.15 * (2.0 * self._noise(fid, ig, jg) - 1.0))

    def uhrho_raw(self, ig, jg, k):
        """unscaled uhrho_et at the east face of (ig, jg)."""
        s = self.s
        igs, jgs, valid = self._src(ig, jg)
        if s.flow == "zonal":
            m = self.metrics(igs, jgs)
            return RHO0 * self.dzt[k - 1] * m["dxte"] / s.dtime * torch.ones_like(m["dxte"]) * valid
        # face wet depth = min(kmt(i), kmt(i+1)); the east neighbour is evaluated from the SOURCE point so that
        # a halo face is the same physical face as its source
        K = torch.minimum(self.kmt(igs, jgs), self.kmt(igs + 1, jgs))
        m = self.metrics(igs, jgs)
        F = self._face_amp(_FID["fu"], igs, jgs, 2.0, 1.0, 0.4)
        return RHO0 * m["dxte"] / s.dtime * F * self.P[K, k - 1] * valid

    def vhrho_raw(self, ig, jg, k):
        s = self.s
        igs, jgs, valid = self._src(ig, jg)
        if s.flow == "zonal":
            return torch.zeros_like(igs, dtype=F64) * valid
        K = torch.minimum(self.kmt(igs, jgs), self.kmt(igs, jgs + 1))
        m = self.metrics(igs, jgs)
        F = self._face_amp(_FID["fv"], igs, jgs, 1.0, 3.0, 1.3)
        v = RHO0 * m["dytn"] / s.dtime * F * self.P[K, k - 1] * valid
        if s.tripolar:
            # the north face of row nj is shared by (i, nj) and (ni+1-i, nj) with opposite orientation:
            # make the two copies exactly antisymmetric, as a consistent C-grid transport is
            top = jgs == s.nj
            west = igs <= s.ni // 2
            vm = None
            if bool(top.any()):
                igm = s.ni + 1 - igs
                Km = torch.minimum(self.kmt(igm, jgs), self.kmt(igm, jgs + 1))
                mm = self.metrics(igm, jgs)
                Fm = self._face_amp(_FID["fv"], igm, jgs, 1.0, 3.0, 1.3)
                vm = -(RHO0 * mm["dytn"] / s.dtime * Fm * self.P[Km, k - 1])
                v = torch.where(top & ~west, vm, v)
        return v

    def tracer(self, n, ig, jg, k, tau=False):
        s = self.s
        igs, jgs, _ = self._src(ig, jg)
        x = igs.to(F64)
        y = jgs.to(F64)
        z = self.zt[k - 1]
        u = self._noise(_FID["T"] + n, igs, jgs, k)
        if n == 0:    # temperature: stratified + large-scale + noise
            t = 25.0 * torch.exp(-z / 800.0) + 2.0 * torch.sin(2 * math.pi * x / s.ni) * torch.cos(math.pi * y / s.nj) + 1e-3 * (2 * u - 1)
        elif n == 1:  # salinity
            t = 35.0 + 0.5 * (2 * u - 1) + 0.0 * z
        elif n == 2:  # square pulse passive tracer in [0,1]
            inside = ((x > 0.25 * s.ni) & (x < 0.5 * s.ni) & (y > 0.3 * s.nj) & (y < 0.6 * s.nj))
            t = inside.to(F64) + 0.0 * z
        elif n == 3:  # Gaussian passive tracer
            t = torch.exp(-(((x - 0.6 * s.ni) / (0.08 * s.ni)) ** 2 + ((y - 0.4 * s.nj) / (0.1 * s.nj)) ** 2)) + 0.0 * z
        else:         # level-tagged / random passive tracers (gyre: tracer n = 1 on level n)
            lvl = ((n - 4) * 3) % s.nk + 1
            t = (torch.as_tensor(k, device=x.device) == lvl).to(F64) * torch.ones_like(x) + 0.05 * u
        if tau:
            t = t + 1e-2 * (2 * self._noise(_FID["ttau"] + 8 * n, igs, jgs, k) - 1)
        return t

    # ---- block assembly --------------------------------------------------------------------------
    def _index_grids(self, i0, i1, j0, j1, halo=1):
        ig = torch.arange(i0 - halo, i1 + halo + 1, dtype=I64, device=self.dev)[None, None, :]
        jg = torch.arange(j0 - halo, j1 + halo + 1, dtype=I64, device=self.dev)[None, :, None]
        k = torch.arange(1, self.s.nk + 1, dtype=I64, device=self.dev)[:, None, None]
        return ig, jg, k

    def _cfl_max(self, i0, i1, j0, j1) -> float:
        """max directional CFL of the UNSCALED flow on the block's compute domain (+ its W/S faces)."""
        b = self._assemble(i0, i1, j0, j1, scale=1.0, ntr=0)
        dt = self.s.dtime
        g = b.grid2d
        rho, u, v, w = b.rho_dzt, b.uhrho_et, b.vhrho_nt, b.wrho_bt
        cx = (u[:, 1:-1, :-1] * dt * 2.0 / ((rho[:, 1:-1, :-1] + rho[:, 1:-1, 1:]) * g["dxte"][1:-1, :-1])).abs().max()
        cy = (v[:, :-1, 1:-1] * dt * 2.0 / ((rho[:, :-1, 1:-1] + rho[:, 1:, 1:-1]) * g["dytn"][:-1, 1:-1])).abs().max()
        cz = (w[1:, 1:-1, 1:-1] * dt / rho[:, 1:-1, 1:-1]).abs().max()
        return float(torch.stack([cx, cy, cz]).max())

    def calibrate(self, i0=None, i1=None, j0=None, j1=None, reduce_max=None) -> float:
        """Set spec.flow_scale so that the global max directional CFL equals spec.cfl.
        ``reduce_max(float)->float`` performs the cross-rank max when each rank holds only its block."""
        s = self.s
        if s.flow_scale is not None:
            return s.flow_scale
        i0, i1, j0, j1 = i0 or 1, i1 or s.ni, j0 or 1, j1 or s.nj
        c = self._cfl_max(i0, i1, j0, j1)
        if reduce_max is not None:
            c = reduce_max(c)
        # round the scale to a power of two times cfl so it is reproducible independent of the reduction order
        s.flow_scale = s.cfl / c if c > 0 else 1.0
        return s.flow_scale

    def _assemble(self, i0, i1, j0, j1, scale: float, ntr: int, with_tau=False) -> BlockInputs:
        s = self.s
        ig, jg, k = self._index_grids(i0, i1, j0, j1)
        ig2, jg2 = ig[0], jg[0]
        grid2d = {n: t.expand(jg2.shape[0], ig2.shape[1]).contiguous() for n, t in self.metrics(ig2, jg2).items()}
        kmt = self.kmt(ig, jg)
        tmask = (k <= kmt).to(F64).expand(s.nk, jg.shape[1], ig.shape[2]).contiguous()
        shp = (s.nk, jg.shape[1], ig.shape[2])
        rho = self.rho_dzt(ig, jg, k).expand(shp).contiguous()
        u = (self.uhrho_raw(ig, jg, k) * scale).expand(shp).contiguous()
        v = (self.vhrho_raw(ig, jg, k) * scale).expand(shp).contiguous()
        # continuity (ocean_advection_velocity.F90:635-641, ocean_operators.F90:945-958) on i0-1+1.. : we need
        # w on the compute domain only; halo ring of w is evaluated the same way from one more ring of u, v
        igw, jgw, _ = self._index_grids(i0, i1, j0, j1, halo=2)
        mw = self.metrics(igw[0], jgw[0])
        uw = self.uhrho_raw(igw, jgw, k) * scale
        vw = self.vhrho_raw(igw, jgw, k) * scale
        U = mw["dyte"] * uw
        V = mw["dxtn"] * vw
        kmtw = self.kmt(igw, jgw)
        tmw = (k <= kmtw).to(F64)
        div = tmw[:, 1:, 1:] * ((U[:, 1:, 1:] - U[:, 1:, :-1]) * mw["datr"][1:, 1:] + (V[:, 1:, 1:] - V[:, :-1, 1:]) * mw["datr"][1:, 1:])
        w = torch.zeros((s.nk + 1,) + tuple(div.shape[1:]), dtype=F64, device=self.dev)
        for kk in range(1, s.nk + 1):
            w[kk] = (div[kk - 1] + w[kk - 1]) * tmw[kk - 1, 1:, 1:]
        w = w[:, :-1, :-1].contiguous()  # rows/cols of ring 1 .. ring 1  -> the halo-1 data domain
        b = BlockInputs(s, i0, i1, j0, j1, grid2d, self.dzt, tmask, rho, u, v, w)
        for n in range(ntr):
            b.T.append(self.tracer(n, ig, jg, k).expand(shp).contiguous())
            if with_tau:
                b.T_tau.append(self.tracer(n, ig, jg, k, tau=True).expand(shp).contiguous())
            igs, jgs, _ = self._src(ig, jg)
            b.th_tendency.append((1e-3 * (2 * self._noise(_FID["th"] + 8 * n, igs, jgs, k) - 1)).expand(shp).contiguous())
            b.tmask_limit.append((tmask * (self._noise(_FID["tlimit"] + 8 * n, igs, jgs, k) < 0.15).to(F64)).contiguous())
        return b

    def block(self, i0=None, i1=None, j0=None, j1=None, ntr=None, with_tau=False) -> BlockInputs:
        s = self.s
        i0, i1, j0, j1 = i0 or 1, i1 or s.ni, j0 or 1, j1 or s.nj
        if s.flow_scale is None:
            if (i0, i1, j0, j1) != (1, s.ni, 1, s.nj):
                raise RuntimeError("call calibrate() (with a cross-rank max) before generating partial blocks")
            self.calibrate()
        return self._assemble(i0, i1, j0, j1, s.flow_scale, s.ntr if ntr is None else ntr, with_tau=with_tau)


def make_case(name: str, device="cpu", **overrides) -> Generator:
    import dataclasses
    spec = dataclasses.replace(CASES[name], **overrides)
    return Generator(spec, device=device)
