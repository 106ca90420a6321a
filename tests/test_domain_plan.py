"""Host-side decomposition logic: layout / extents / halo source map / exchange plans.
Three independent implementations must agree: mom5_b200/domain.py (product, Python), the C++ plan builder
inside libmom5adv.so (product, reached through its debug hooks -- no GPU needed) and the C oracle."""
import ctypes as C

import numpy as np
import pytest

from mom5_b200 import _lib
from mom5_b200.domain import XUPDATE, YUPDATE, Decomposition, compute_extent, define_layout
from oracle import oracle as orc


def test_define_layout_known_values():
    # mpp_define_layout2D: 8 ranks on 3600 x 2700 -> (2,4)
    assert define_layout(3600, 2700, 8) == (2, 4)
    assert define_layout(3600, 2700, 1) == (1, 1)
    assert define_layout(360, 300, 2) == (2, 1)
    assert define_layout(1440, 1080, 4) == (2, 2)
    assert define_layout(1, 100, 4) == (1, 4)
    L = orc.lib()
    for ni, nj, n in [(3600, 2700, 8), (360, 300, 24), (1440, 1080, 960), (24, 35, 6), (100, 7, 5)]:
        out = (C.c_int * 2)()
        L.orc_define_layout(ni, nj, n, out)
        assert (out[0], out[1]) == define_layout(ni, nj, n)


def test_compute_extent_mirror_symmetry():
    # the comment in mpp_compute_extent: nx=18, n=4 -> 4554 or 5445 are solutions, 4455 is not
    b, e = compute_extent(1, 18, 4)
    sizes = [e[d] - b[d] + 1 for d in range(4)]
    assert sizes in ([4, 5, 5, 4], [5, 4, 4, 5])
    assert sizes == sizes[::-1]


@pytest.mark.parametrize("npts,n", [(18, 4), (360, 7), (2700, 8), (3600, 2), (35, 3), (24, 5), (300, 16), (11, 2), (10, 3), (75, 1)])
def test_compute_extent_three_implementations(npts, n):
    b, e = compute_extent(1, npts, n)
    assert b[0] == 1 and e[-1] == npts and all(b[d + 1] == e[d] + 1 for d in range(n - 1))
    ob, oe = (C.c_int * n)(), (C.c_int * n)()
    assert orc.lib().orc_compute_extent(1, npts, n, ob, oe) == 0
    assert list(ob) == b and list(oe) == e
    lb, le = (C.c_int * n)(), (C.c_int * n)()
    assert _lib.load().mom5adv_debug_extent(1, npts, n, lb, le) == 0
    assert list(lb) == b and list(le) == e


def _lib_plan(dec, rank, flags, halo=2):
    out = (C.c_int * (7 * 256))()
    n = _lib.load().mom5adv_debug_plan(dec.ni_g, dec.nj_g, dec.px, dec.py, int(dec.cyclic_x), int(dec.cyclic_y),
                                       int(dec.tripolar), rank, flags, halo, out, 256)
    assert n >= 0
    rows = [tuple(out[7 * r:7 * r + 7]) for r in range(n)]
    return [r[1:] for r in rows if r[0] == 1], [r[1:] for r in rows if r[0] == 0]


CONFIGS = [dict(ni_g=40, nj_g=30, px=2, py=2, cyclic_x=True, tripolar=True),
           dict(ni_g=40, nj_g=30, px=1, py=4, cyclic_x=True, tripolar=True),
           dict(ni_g=40, nj_g=30, px=4, py=2, cyclic_x=True, tripolar=True),
           dict(ni_g=33, nj_g=27, px=3, py=2),
           dict(ni_g=32, nj_g=24, px=2, py=3, cyclic_x=True, cyclic_y=True),
           dict(ni_g=32, nj_g=24, px=1, py=1, cyclic_x=True, cyclic_y=True),
           dict(ni_g=20, nj_g=14, px=1, py=1, cyclic_x=True, tripolar=True),
           dict(ni_g=36, nj_g=20, px=2, py=1, cyclic_x=True, tripolar=True)]


@pytest.mark.parametrize("halo", [1, 2, 4])    # data-domain fields, MDFL / quicker scratch, MDPPM scratch
@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("flags", [XUPDATE, YUPDATE, XUPDATE | YUPDATE])
def test_exchange_plan_python_equals_library(cfg, flags, halo):
    dec = Decomposition(**cfg)
    for rank in range(dec.nranks):
        sends, recvs = dec.exchange_plan(rank, flags, halo=halo)
        ls, lr = _lib_plan(dec, rank, flags, halo)
        assert [(m.peer, m.i0, m.i1, m.j0, m.j1, int(m.flip)) for m in sends] == ls
        assert [(m.peer, m.i0, m.i1, m.j0, m.j1, int(m.flip)) for m in recvs] == lr


def _apply_plans(dec, fields, flags, halo=2):
    """execute every rank's plan on numpy h2 arrays (what pack / send / recv / unpack do on the GPUs)"""
    mail = {}
    for r in range(dec.nranks):
        sends, _ = dec.exchange_plan(r, flags, halo)
        per_peer = {}
        for m in sends:
            blk = fields[r][:, m.j0 - 1 + halo:m.j1 + halo, m.i0 - 1 + halo:m.i1 + halo].copy()
            per_peer.setdefault(m.peer, []).append(blk)
        for p, lst in per_peer.items():
            mail[(r, p)] = lst
    for r in range(dec.nranks):
        _, recvs = dec.exchange_plan(r, flags, halo)
        cursor = {}
        for m in recvs:
            q = cursor.get(m.peer, 0)
            blk = mail[(m.peer, r)][q]
            cursor[m.peer] = q + 1
            if m.flip:
                blk = blk[:, ::-1, ::-1]
            fields[r][:, m.j0 - 1 + halo:m.j1 + halo, m.i0 - 1 + halo:m.i1 + halo] = blk


@pytest.mark.parametrize("halo", [1, 2, 4])
@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("flags", [XUPDATE, YUPDATE, XUPDATE | YUPDATE])
def test_plan_execution_equals_oracle_update(cfg, flags, halo):
    dec = Decomposition(**cfg)
    nk = 3
    rng = np.random.default_rng(5)
    a, b = [], []
    for r in range(dec.nranks):
        ni, nj = dec.local_size(r)
        f = rng.standard_normal((nk, nj + 2 * halo, ni + 2 * halo))
        a.append(f.copy())
        b.append(f.copy())
    _apply_plans(dec, a, flags, halo)
    ib, ie = (C.c_int * dec.px)(*dec.ibeg), (C.c_int * dec.px)(*dec.iend)
    jb, je = (C.c_int * dec.py)(*dec.jbeg), (C.c_int * dec.py)(*dec.jend)
    lay = orc.OrcLayout(dec.ni_g, dec.nj_g, dec.px, dec.py, ib, ie, jb, je, int(dec.cyclic_x), int(dec.cyclic_y), int(dec.tripolar))
    orc.lib().orc_update_halo(C.byref(lay), orc._pp(b), nk, halo, flags)
    for r in range(dec.nranks):
        assert np.array_equal(a[r], b[r]), f"rank {r}"
