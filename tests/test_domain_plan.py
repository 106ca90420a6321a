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


def test_overlap_edge_sets_cover_every_cell_a_strip_touches():
    """The comm/compute overlap of both Sweby drivers rests on two invariants (round-1 advisor finding: with ni_local = 513 the E/W pack
    could read a column of an interior z tile that had not run yet):
      (1) every cell a strip PACK reads was produced by a tile of the edge set, which runs before the exchange starts;
      (2) no tile that runs UNDER the exchange reads a cell the strip UNPACK writes.
    Checked on the library's own tile sets (mom5adv_debug_overlap_sets, the very function the drivers call) for every block
    width / height up to 1100 plus the bench shapes."""
    import ctypes as C
    from mom5_b200 import _lib
    L = _lib.load()
    out = (C.c_int * 12)()

    def sets(ni, nj, nk):
        assert L.mom5adv_debug_overlap_sets(ni, nj, nk, out) == 0
        return list(out)

    shapes = [(n, 80, 6) for n in range(2, 1101)] + [(520, n, 6) for n in range(2, 1101)] + \
             [(3600, 2700, 75), (1800, 675, 75), (1800, 1350, 75), (1800, 2700, 75), (513, 41, 6), (497, 80, 6), (720, 540, 50)]
    for ni, nj, nk in shapes:
        nzt, z_last, f_rows, njc_f, c_hi, nxt, x_last, y_rows, njc_y, y_last, zbx, xw = sets(ni, nj, nk)
        # ---- z tiles: tile t = data-domain columns t*zbx .. t*zbx + zbx-1; the E/W pack reads columns 1, 2, ni-1, ni
        assert nzt == ni // zbx + 1
        edge_z = {0} | set(range(nzt - z_last, nzt))
        for col in (1, min(2, ni), max(ni - 1, 1), ni):
            assert col // zbx in edge_z, (ni, col, "z edge set")
        # ---- fused pass: interior chunks 1..c_hi read rows js-2 .. je+2 with js = jc*R+1, je = (jc+1)*R: none may be a halo row
        for jc in range(1, c_hi + 1):
            js, je = jc * f_rows + 1, min((jc + 1) * f_rows, nj)
            assert js - 2 >= 1 and je + 2 <= nj, (nj, f_rows, jc, "fused interior chunk reads a halo row")
        assert c_hi <= njc_f - 1
        # ---- three-sweep driver, x: interior tiles 1 .. nxt-1-x_last read tm(31t-1 .. 31t+33)
        assert nxt == (ni + 30) // 31
        for t in range(1, nxt - x_last):
            assert xw * t - 1 >= 1 and xw * t + 33 <= ni, (ni, t, "x interior tile reads a halo column")
        # ---- three-sweep driver, y: interior chunks 1 .. njc_y-1-y_last read rows cR-1 .. (c+1)R+2
        for c in range(1, njc_y - y_last):
            assert c * y_rows - 1 >= 1 and (c + 1) * y_rows + 2 <= nj, (nj, y_rows, c, "y interior chunk reads a halo row")
