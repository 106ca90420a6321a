"""FMS's own known-answer pattern for halo updates: global(i,j,k) = k + i*1e-3 + j*1e-6
(src/shared/mpp/test_mpp_domains.F90:5628-5634) with the expected halos written as test_mpp_domains does for the
'Simple', 'Cyclic' and 'Folded-north' domain types (:5641-5685, fill_folded_north_halo :3749-3766).
Checked for the C oracle's block update and for the python filler used while generating the golden vectors."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from mom5_b200.domain import XUPDATE, YUPDATE, Decomposition
from oracle import oracle as orc

NX, NY, NZ, H = 24, 18, 3, 2


def expected_global(kind):
    """(nz, ny+2h, nx+2h) array following the Fortran reference fill; np.nan where no update may happen"""
    g = np.zeros((NZ, NY + 2 * H, NX + 2 * H))
    A = lambda i, j: (slice(None), j - 1 + H, i - 1 + H)   # 1-based access helper (scalars)
    for k in range(1, NZ + 1):
        for j in range(1, NY + 1):
            for i in range(1, NX + 1):
                g[k - 1, j - 1 + H, i - 1 + H] = k + i * 1e-3 + j * 1e-6
    if kind == "Simple":
        return g
    def sl(a, b):   # Fortran a:b (1-based, inclusive) -> python slice in halo'd coordinates
        return slice(a - 1 + H, b + H)
    # west / east (both Cyclic and Folded-north): global(1-whalo:0,1:ny) = global(nx-whalo+1:nx,1:ny) etc.
    g[:, sl(1, NY), sl(1 - H, 0)] = g[:, sl(1, NY), sl(NX - H + 1, NX)]
    g[:, sl(1, NY), sl(NX + 1, NX + H)] = g[:, sl(1, NY), sl(1, H)]
    if kind == "Cyclic":
        g[:, sl(1 - H, 0), sl(1 - H, NX + H)] = g[:, sl(NY - H + 1, NY), sl(1 - H, NX + H)]
        g[:, sl(NY + 1, NY + H), sl(1 - H, NX + H)] = g[:, sl(1, H), sl(1 - H, NX + H)]
    elif kind == "Folded-north":
        # fill_folded_north_halo(global, 0,0,0,0, 1): m1 = m2 = 0
        # data(1-whalo:0, ny+1:ny+nhalo) = data(whalo:1:-1, ny:ny-nhalo+1:-1)
        g[:, sl(NY + 1, NY + H), sl(1 - H, 0)] = g[:, sl(NY - H + 1, NY), sl(1, H)][:, ::-1, ::-1]
        # data(1:nx, ny+1:ny+nhalo) = data(nx:1:-1, ny:ny-nhalo+1:-1)
        g[:, sl(NY + 1, NY + H), sl(1, NX)] = g[:, sl(NY - H + 1, NY), sl(1, NX)][:, ::-1, ::-1]
        # data(nx+1:nx+ehalo, ny+1:ny+nhalo) = data(nx:nx-ehalo+1:-1, ny:ny-nhalo+1:-1)
        g[:, sl(NY + 1, NY + H), sl(NX + 1, NX + H)] = g[:, sl(NY - H + 1, NY), sl(NX - H + 1, NX)][:, ::-1, ::-1]
    return g


KINDS = {"Simple": dict(), "Cyclic": dict(cyclic_x=True, cyclic_y=True), "Folded-north": dict(cyclic_x=True, tripolar=True)}


@pytest.fixture(params=[2, 4, 1])
def halo_width(request):
    """the KAT at the three halo widths of the path: 2 (MDFL / quicker scratch), 4 (MDPPM scratch), 1 (data-domain fields)"""
    global H
    old, H = H, request.param
    yield request.param
    H = old


@pytest.mark.parametrize("kind", list(KINDS))
@pytest.mark.parametrize("layout", [(1, 1), (2, 2), (3, 2), (4, 1), (1, 3)])
def test_oracle_update_matches_fms_kat(kind, layout, halo_width):
    dec = Decomposition(NX, NY, layout[0], layout[1], **KINDS[kind])
    exp = expected_global(kind)
    fields = []
    for r in range(dec.nranks):
        i0, i1, j0, j1 = dec.extent(r)
        f = np.zeros((NZ, j1 - j0 + 1 + 2 * H, i1 - i0 + 1 + 2 * H))
        f[:, H:-H, H:-H] = exp[:, j0 - 1 + H:j1 + H, i0 - 1 + H:i1 + H]
        fields.append(f)
    ib, ie = (C.c_int * dec.px)(*dec.ibeg), (C.c_int * dec.px)(*dec.iend)
    jb, je = (C.c_int * dec.py)(*dec.jbeg), (C.c_int * dec.py)(*dec.jend)
    lay = orc.OrcLayout(NX, NY, dec.px, dec.py, ib, ie, jb, je, int(dec.cyclic_x), int(dec.cyclic_y), int(dec.tripolar))
    orc.lib().orc_update_halo(C.byref(lay), orc._pp(fields), NZ, H, XUPDATE | YUPDATE)
    for r in range(dec.nranks):
        i0, i1, j0, j1 = dec.extent(r)
        want = exp[:, j0 - 1:j1 + 2 * H, i0 - 1:i1 + 2 * H]   # the block with its halo cut from the global answer
        assert np.array_equal(fields[r], want), f"{kind} layout {layout} rank {r}"


@pytest.mark.parametrize("kind", list(KINDS))
def test_golden_generator_filler_matches_fms_kat(kind, halo_width):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from f90interp import FArray
    from gen_from_reference import Halo
    dec = Decomposition(NX, NY, 1, 1, **KINDS[kind])
    exp = expected_global(kind)
    f = np.zeros_like(exp)
    f[:, H:-H, H:-H] = exp[:, H:-H, H:-H]
    fa = FArray(data=f, lo=[1 - H, 1 - H, 1])
    Halo(dec, 1, NX, 1, NY).scalar(fa, XUPDATE | YUPDATE)
    assert np.array_equal(fa.a, exp)


def test_xupdate_yupdate_do_not_touch_corners():
    """mpp_do_update.h:57-78: XUPDATE = E/W only, YUPDATE = N/S only; corners only when both are set"""
    dec = Decomposition(NX, NY, 2, 2, cyclic_x=True, cyclic_y=True)
    for flags, touched in ((XUPDATE, "ew"), (YUPDATE, "ns")):
        fields = []
        for r in range(4):
            ni, nj = dec.local_size(r)
            f = np.full((1, nj + 2 * H, ni + 2 * H), -9.0)
            f[:, H:-H, H:-H] = 1.0
            fields.append(f)
        ib, ie = (C.c_int * 2)(*dec.ibeg), (C.c_int * 2)(*dec.iend)
        jb, je = (C.c_int * 2)(*dec.jbeg), (C.c_int * 2)(*dec.jend)
        lay = orc.OrcLayout(NX, NY, 2, 2, ib, ie, jb, je, 1, 1, 0)
        orc.lib().orc_update_halo(C.byref(lay), orc._pp(fields), 1, H, flags)
        f = fields[0][0]
        assert (f[:H, :H] == -9).all() and (f[-H:, -H:] == -9).all()          # corners untouched
        assert ((f[H:-H, :H] == 1).all()) == (touched == "ew")
        assert ((f[:H, H:-H] == 1).all()) == (touched == "ns")


def _cgrid_pattern():
    """x / y flux components on a halo-1 global array, FMS pattern on the compute domain (test_mpp_domains.F90:5908-5919)"""
    g = np.zeros((NZ, NY + 2, NX + 2))
    for k in range(1, NZ + 1):
        for j in range(1, NY + 1):
            for i in range(1, NX + 1):
                g[k - 1, j, i] = k + i * 1e-3 + j * 1e-6
    return g


def _cgrid_expected_fold_row(g2):
    """'redundant points must be equal and opposite' (test_mpp_domains.F90:5965): global2(nx/2+1:nx, ny) = -global2(nx/2:1:-1, ny)"""
    row = g2[:, NY, 1:NX + 1].copy()
    row[:, NX // 2:] = -row[:, NX // 2 - 1::-1]
    return row


def test_generator_cgrid_ne_filler_matches_fms_kat():
    """mpp_update_domains(flux_x, flux_y, gridtype=CGRID_NE) on the folded-north domain, as the quicker path uses it
    (OTA:2640): fold-line rule for the NORTH-position component and the cyclic west / east columns (fill_folded_north_halo's
    first two statements, :3758-3759) against FMS's own known answer."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from f90interp import FArray
    from gen_from_reference import Halo
    dec = Decomposition(NX, NY, 1, 1, cyclic_x=True, tripolar=True)
    g1, g2 = _cgrid_pattern(), _cgrid_pattern()
    want_row = _cgrid_expected_fold_row(g2)
    fx, fy = FArray(data=g1.copy(), lo=[0, 0, 1]), FArray(data=g2.copy(), lo=[0, 0, 1])
    Halo(dec, 1, NX, 1, NY).cgrid_ne(fx, fy)
    assert np.array_equal(fy.a[:, NY, 1:NX + 1], want_row)
    assert np.array_equal(fy.a[:, 1:NY, 1:NX + 1], g2[:, 1:NY, 1:NX + 1])          # nothing else on the compute domain moves
    assert np.array_equal(fx.a[:, 1:NY + 1, 1:NX + 1], g1[:, 1:NY + 1, 1:NX + 1])
    assert np.array_equal(fx.a[:, 1:NY + 1, 0], g1[:, 1:NY + 1, NX])                # west halo  = east edge
    assert np.array_equal(fx.a[:, 1:NY + 1, NX + 1], g1[:, 1:NY + 1, 1])            # east halo  = west edge


@pytest.mark.parametrize("layout", [(1, 1), (2, 1), (3, 2), (4, 1), (6, 3)])
def test_oracle_fold_line_fix_matches_fms_kat(layout):
    """the oracle's multi-block fold-line fix (orc_fold_fix_flux; the CUDA library's fold_line_fix is checked against it on
    the GPU) reproduces FMS's known answer for every layout, including mirror images owned by another block"""
    dec = Decomposition(NX, NY, layout[0], layout[1], cyclic_x=True, tripolar=True)
    g2 = _cgrid_pattern()
    want_row = _cgrid_expected_fold_row(g2)
    fx, fy = [], []
    for r in range(dec.nranks):
        i0, i1, j0, j1 = dec.extent(r)
        f = np.zeros((NZ, j1 - j0 + 3, i1 - i0 + 3))
        f[:, 1:-1, 1:-1] = g2[:, j0:j1 + 1, i0:i1 + 1]
        fy.append(f)
        fx.append(f.copy())
    ib, ie = (C.c_int * dec.px)(*dec.ibeg), (C.c_int * dec.px)(*dec.iend)
    jb, je = (C.c_int * dec.py)(*dec.jbeg), (C.c_int * dec.py)(*dec.jend)
    lay = orc.OrcLayout(NX, NY, dec.px, dec.py, ib, ie, jb, je, 1, 0, 1)
    orc.lib().orc_fold_fix_flux(C.byref(lay), orc._pp(fx), orc._pp(fy), C.c_int(NZ))
    for r in range(dec.nranks):
        i0, i1, j0, j1 = dec.extent(r)
        want = g2[:, j0:j1 + 1, i0:i1 + 1].copy()
        if j1 == NY:
            want[:, -1, :] = want_row[:, i0 - 1:i1]
        assert np.array_equal(fy[r][:, 1:-1, 1:-1], want), f"layout {layout} rank {r}"
