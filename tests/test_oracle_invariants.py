"""Reference-sanctioned invariants of the advection path, checked on the CPU oracle:
decomposition invariance (doc/testcase_ocean_solo.pdf section 1.1: changing the PE count does not change answers,
bitwise), tendency of a uniform tracer, conservation, monotonicity of the limited scheme, positivity of upwind."""
import numpy as np
import pytest
import torch

from mom5_b200.synthetic import make_case
from oracle.oracle import Oracle, gather, split_blocks
from tests.util import assert_bit_equal


def _global(case, **over):
    g = make_case(case, **over)
    return g, g.block(with_tau=True)


def _sweby(g, gb, px, py, Tlist=None):
    dec = g.s.decomposition(px, py)
    blocks = split_blocks(dec, gb)
    o = Oracle(dec, blocks)
    T = [[t.numpy() for t in b.T] for b in blocks] if Tlist is None else Tlist(blocks)
    th = [[np.zeros_like(t.numpy()) for t in b.th_tendency] for b in blocks]
    out = o.sweby_all(T, th, g.s.dtime)
    ntr = len(T[0])
    return o, dec, [gather(dec, [out["adv"][b][n] for b in range(dec.nranks)]) for n in range(ntr)]


@pytest.mark.parametrize("case", ["mini_tripolar", "mini_walls", "mini_torus"])
def test_decomposition_invariance_sweby_all(case):
    g, gb = _global(case)
    _, _, ref = _sweby(g, gb, 1, 1)
    for (px, py) in [(2, 2), (2, 4), (1, 8), (4, 1), (3, 2)]:
        _, _, got = _sweby(g, gb, px, py)
        for n in range(len(ref)):
            assert_bit_equal(got[n], ref[n], f"{case} layout {px}x{py} tracer {n}")


@pytest.mark.parametrize("case", ["mini_tripolar", "mini_walls"])
def test_decomposition_invariance_other_schemes(case):
    g, gb = _global(case)
    ref = {}
    for (px, py) in [(1, 1), (2, 2), (1, 4), (4, 2)]:
        dec = g.s.decomposition(px, py)
        blocks = split_blocks(dec, gb)
        o = Oracle(dec, blocks)
        Tm1 = [b.T[0].numpy() for b in blocks]
        Tt = [b.T_tau[0].numpy() for b in blocks]
        tl = [b.tmask_limit[0].numpy() for b in blocks]
        res = dict(
            mdfl=o.mdfl_sweby(Tm1, g.s.dtime, 1.0)["wrk1"], dst=o.mdfl_sweby(Tm1, g.s.dtime, 0.0)["wrk1"],
            mdflt=o.sweby_test(Tm1, g.s.dtime, 1.0)["wrk1"], dstt=o.sweby_test(Tm1, g.s.dtime, 0.0)["wrk1"],
            ppm1=o.mdppm(Tm1, g.s.dtime, 1)["wrk1"], ppm2=o.mdppm(Tm1, g.s.dtime, 2)["wrk1"], ppm3=o.mdppm(Tm1, g.s.dtime, 3)["wrk1"],
            qh=o.horz_quicker(Tm1, Tt, tl, False)["wrk1"], qhl=o.horz_quicker(Tm1, Tt, tl, True)["wrk1"],
            qv=o.vert_quicker(Tm1, Tt, tl)["wrk1"], uh=o.horz_upwind(Tm1)["wrk1"], uv=o.vert_upwind(Tm1)["wrk1"])
        for k, v in res.items():
            gv = gather(dec, v)
            if (px, py) == (1, 1):
                ref[k] = gv
            else:
                assert_bit_equal(gv, ref[k], f"{case} {k} layout {px}x{py}")


@pytest.mark.parametrize("case", ["mini_tripolar", "mini_walls", "mini_torus"])
def test_uniform_tracer_has_round_off_tendency(case):
    """T = const: every flux divergence is cancelled by the T*(mass divergence) terms -> tendency ~ round-off"""
    g, gb = _global(case)
    c = 3.75

    def Tl(blocks):
        return [[np.full_like(b.T[0].numpy(), c)] for b in blocks]

    o, dec, adv = _sweby(g, gb, 1, 1, Tl)
    scale = c * gb.rho_dzt.max().item() / g.s.dtime      # size of rho_dzt*T/dtime
    assert np.abs(adv[0]).max() < 1e-12 * scale


@pytest.mark.parametrize("case", ["mini_tripolar", "mini_walls", "mini_torus"])
def test_global_content_conserved(case):
    """sum over the closed / periodic domain of dat * advective tendency vanishes to round-off
    (total_tracer, ocean_tracer_diag.F90:2405-2408, is unchanged by advection)"""
    g, gb = _global(case)
    o, dec, adv = _sweby(g, gb, 1, 1)
    dat = gb.grid2d["dat"][1:-1, 1:-1].numpy()
    for n in range(len(adv)):
        net = float((adv[n] * dat[None]).sum())
        gross = float(np.abs(adv[n] * dat[None]).sum())
        assert abs(net) < 1e-11 * gross, (n, net, gross)


def test_square_pulse_on_torus_stays_monotone():
    """doc/testcase_ocean_solo.pdf section 2: the limited scheme keeps the [0,1] square pulse within bounds"""
    g, gb = _global("mini_torus", ni=48, nj=16, nk=4, ntr=3, rho_noise=0.0)   # uniform thickness: exactly non-divergent volume flow
    dec = g.s.decomposition(1, 1)
    o = Oracle(dec, [gb])
    rho = gb.rho_dzt.numpy()
    T = gb.T[2].numpy().copy()          # tracer 3 = square pulse
    assert T.min() == 0.0 and T.max() == 1.0
    for step in range(12):
        th = [[np.zeros_like(T)]]
        o.sweby_all([[T]], th, g.s.dtime)
        # consumer (ocean_tracer.F90:2341-2350) with steady rho_dzt
        T = (rho * T + g.s.dtime * th[0][0]) / rho
        o.update([T], 1, 3)              # halo-1 update of field(taup1) (ocean_model.F90:1903-1911)
        c = T[:, 1:-1, 1:-1]
        assert c.min() > -1e-12 and c.max() < 1.0 + 1e-12, (step, c.min(), c.max())


def test_dst_linear_equals_unlimited_psi():
    """sweby_limiter = 0: psi = d0 + d1*theta (OTA:3874-3884) -> differs from the limited scheme at extrema"""
    g, gb = _global("mini_walls")
    dec = g.s.decomposition(1, 1)
    o = Oracle(dec, [gb])
    T = [gb.T[0].numpy()]
    a = o.mdfl_sweby(T, g.s.dtime, 1.0)["wrk1"][0]
    b = o.mdfl_sweby(T, g.s.dtime, 0.0)["wrk1"][0]
    assert np.isfinite(a).all() and np.isfinite(b).all() and not np.array_equal(a, b)


def test_upwind_keeps_a_nonnegative_tracer_nonnegative():
    g, gb = _global("mini_walls", cfl=0.3)
    dec = g.s.decomposition(1, 1)
    o = Oracle(dec, [gb])
    T = np.abs(gb.T[1].numpy() - 35.0)          # non-negative, rough
    h = o.horz_upwind([T])["wrk1"][0]
    v = o.vert_upwind([T])["wrk1"][0]
    rho = gb.rho_dzt.numpy()
    Tn = (rho * T + g.s.dtime * (h + v)) / rho
    assert Tn[:, 1:-1, 1:-1].min() > -1e-12


def test_sweby_all_vs_single_tracer_variant_agree_to_round_off():
    """a1 and a2 use different association orders: equal to ~1e-13 relative, not bitwise"""
    g, gb = _global("mini_tripolar")
    dec = g.s.decomposition(1, 1)
    o = Oracle(dec, [gb])
    th = [[np.zeros_like(t.numpy()) for t in gb.T]]
    a1 = o.sweby_all([[t.numpy() for t in gb.T]], th, g.s.dtime)["adv"][0][0]
    a2 = o.mdfl_sweby([gb.T[0].numpy()], g.s.dtime, 1.0)["wrk1"][0]
    assert np.abs(a1 - a2).max() <= 1e-9 * np.abs(a1).max()


@pytest.mark.parametrize("case", ["mini_tripolar", "mini_torus"])
def test_mass_weighted_variant_conserves_and_is_monotone(case):
    """advect_tracer_mdfl_sweby_test carries tracer MASS through the sweeps: the global content change implied by the
    tendency is round-off (fluxes telescope), and with the limiter on a [0,1] field stays within [0,1]."""
    g, gb = _global(case)
    dec = g.s.decomposition(1, 1)
    o = Oracle(dec, [gb])
    T = gb.T[0].numpy()
    T01 = (T - T.min()) / (T.max() - T.min())
    out = o.sweby_test([T01], g.s.dtime, 1.0)
    wrk1 = out["wrk1"][0][:, 1:-1, 1:-1]
    dat = gb.grid2d["dat"].numpy()[1:-1, 1:-1]
    m = gb.tmask.numpy()[:, 1:-1, 1:-1]
    total = float((wrk1 * dat[None]).sum())
    scale = float(np.abs(wrk1 * dat[None]).sum())
    assert abs(total) <= 1e-11 * scale
    tr = out["tracer"][0][:, 2:-2, 2:-2]
    assert float((tr * m).min()) >= -1e-12 and float((tr * m).max()) <= 1.0 + 1e-12


def test_bench_parity_checksums_are_the_oracles_and_layout_invariant():
    """tests/golden/parity_chksums.json (what bench.py's parity_check and the multi-GPU runs compare device checksums with) is what the
    oracle produces -- as one block AND as 2 x 4 blocks (mpp_chksum is invariant under the PE count, mpp_chksum_int.h:20-38)."""
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("gen_parity_chksums", os.path.join(here, "golden", "gen_parity_chksums.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = json.load(open(os.path.join(here, "golden", "parity_chksums.json")))["global_1deg_ntr3"]
    for layout in ((1, 1), (2, 4)):
        got = mod.compute(layout)
        assert got["th"] == gold["th"] and got["adv"] == gold["adv"], layout
