"""The C oracle must reproduce, bit for bit, the golden vectors obtained by executing the reference's own
Fortran text (tests/golden/gen_from_reference.py -> tests/golden/*.npz).  CPU only."""
import numpy as np
import pytest

from oracle.oracle import Oracle, split_blocks
from tests.util import GOLDEN_NAMES, assert_bit_equal, load_golden


def _oracle(b, px=1, py=1):
    dec = b.spec.decomposition(px, py)
    return dec, Oracle(dec, split_blocks(dec, b))


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_mdfl_init_mask(name):
    b, gold, _ = load_golden(name)
    _, o = _oracle(b)
    o.mdfl_init()
    assert_bit_equal(o.blocks[0].tmask_h2, gold["tmask_mdfl"], "tmask_mdfl")


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_sweby_all(name):
    b, gold, cites = load_golden(name)
    assert "advect_tracer_sweby_all=OTA:" in cites
    _, o = _oracle(b)
    ntr = len(b.T)
    T = [[t.numpy() for t in b.T]]
    th = [[t.numpy().copy() for t in b.th_tendency]]
    out = o.sweby_all(T, th, b.spec.dtime, diag=True)
    diag_map = {"zflux_adv": "flux_z", "xflux_adv": "flux_x", "yflux_adv": "flux_y", "advection_z": "adv_z",
                "advection_x": "adv_x", "advection_y": "adv_y", "sweby_advect": "adv"}
    ncmp = 0
    for n in range(1, ntr + 1):
        assert_bit_equal(th[0][n - 1], gold[f"sweby_all.th_tendency.{n}"], f"th_tendency[{n}]")
        assert_bit_equal(out["adv"][0][n - 1], gold[f"sweby_all.wrk1.{n}"], f"T_prog({n})%wrk1")
        # tracer_mdfl_all: compare the compute domain (halo content after the last update is stale by design)
        assert_bit_equal(out["tm"][0][n - 1][:, 2:-2, 2:-2], gold[f"sweby_all.tm.{n}"][:, 2:-2, 2:-2], f"tm[{n}]")
        for dname, oname in diag_map.items():
            assert_bit_equal(out[oname][0][n - 1], gold[f"sweby_all.diag.{dname}.{n}"], f"{dname}[{n}]")
            ncmp += 1
        # z-integrated fluxes (OTA:4317-4326, 4449-4458): sum k=1..nk in order, compute domain
        for dname, oname in (("xflux_adv_int_z", "flux_x"), ("yflux_adv_int_z", "flux_y")):
            f = out[oname][0][n - 1]
            acc = np.zeros_like(f[0])
            for k in range(f.shape[0]):
                acc[1:-1, 1:-1] = acc[1:-1, 1:-1] + f[k, 1:-1, 1:-1]
            assert_bit_equal(acc, gold[f"sweby_all.diag.{dname}.{n}"], f"{dname}[{n}]")
    assert ncmp == 7 * ntr


@pytest.mark.parametrize("name", GOLDEN_NAMES)
@pytest.mark.parametrize("tag,sl", [("mdfl_sweby", 1.0), ("dst_linear", 0.0)])
def test_mdfl_sweby(name, tag, sl):
    b, gold, _ = load_golden(name)
    _, o = _oracle(b)
    n = int(gold[f"{tag}.tracer"])
    out = o.mdfl_sweby([b.T[n - 1].numpy()], b.spec.dtime, sl)
    assert_bit_equal(out["wrk1"][0], gold[f"{tag}.horz.wrk1"], "wrk1")
    th = b.th_tendency[n - 1].numpy().copy()
    o.L.orc_accumulate  # dispatcher tail
    import ctypes as C
    from oracle.oracle import _ptr
    o.L.orc_accumulate(C.byref(o.blocks[0].c), _ptr(out["wrk1"][0]), _ptr(th))
    assert_bit_equal(th, gold[f"{tag}.horz.th_tendency"], "th_tendency")
    assert_bit_equal(out["flux_x"][0], gold[f"{tag}.flux_x"], "flux_x")
    assert_bit_equal(out["flux_y"][0], gold[f"{tag}.flux_y"], "flux_y")
    assert_bit_equal(out["flux_z"][0][:, 1:-1, 1:-1], gold[f"{tag}.flux_z"][:, 1:-1, 1:-1], "flux_z")
    # vert_advect_tracer is a no-op for the 3-D schemes (OTA:2147-2155): wrk1 zeroed, th unchanged
    assert not gold[f"{tag}.vert.wrk1"].any()
    assert_bit_equal(gold[f"{tag}.vert.th_tendency"], gold[f"{tag}.horz.th_tendency"], "vert no-op")


@pytest.mark.parametrize("name", GOLDEN_NAMES)
@pytest.mark.parametrize("tag,sl", [("mdfl_sweby_test", 1.0), ("dst_linear_test", 0.0)])
def test_mdfl_sweby_test_variant(name, tag, sl):
    """advect_tracer_mdfl_sweby_test (OTA:3469-3746): mass-weighted CFL, sign(1e-30,Rj), three running fields"""
    import ctypes as C
    from oracle.oracle import _ptr
    b, gold, _ = load_golden(name)
    _, o = _oracle(b)
    n = int(gold[f"{tag}.tracer"])
    out = o.sweby_test([b.T[n - 1].numpy()], b.spec.dtime, sl)
    assert_bit_equal(out["wrk1"][0], gold[f"{tag}.horz.wrk1"], "wrk1")
    th = b.th_tendency[n - 1].numpy().copy()
    o.L.orc_accumulate(C.byref(o.blocks[0].c), _ptr(out["wrk1"][0]), _ptr(th))
    assert_bit_equal(th, gold[f"{tag}.horz.th_tendency"], "th_tendency")
    assert_bit_equal(out["flux_x"][0], gold[f"{tag}.flux_x"], "flux_x")
    assert_bit_equal(out["flux_y"][0], gold[f"{tag}.flux_y"], "flux_y")
    assert_bit_equal(out["flux_z"][0][:, 1:-1, 1:-1], gold[f"{tag}.flux_z"][:, 1:-1, 1:-1], "flux_z")
    for nm, key in (("tracer", "tracer_mdfl"), ("tracermass", "tracermass_mdfl"), ("mass", "mass_mdfl")):
        assert_bit_equal(out[nm][0][:, 2:-2, 2:-2], gold[f"{tag}.{key}"][:, 2:-2, 2:-2], key)
    assert not gold[f"{tag}.vert.wrk1"].any()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
@pytest.mark.parametrize("tag,limiter", [("mdppm_cw84", 1), ("mdppm_ifc", 2), ("mdppm_sh", 3)])
def test_mdppm(name, tag, limiter):
    """advect_tracer_mdppm (OTA:5990-6494) with the three edge-value limiters (OTA:6510-6657), halo 4"""
    import ctypes as C
    from oracle.oracle import _ptr
    b, gold, _ = load_golden(name)
    _, o = _oracle(b)
    o.mdppm_init()
    assert_bit_equal(o.m4[0], gold["tmask_mdppm"], "tmask_mdppm")
    n = int(gold[f"{tag}.tracer"])
    out = o.mdppm([b.T[n - 1].numpy()], b.spec.dtime, limiter)
    assert_bit_equal(out["wrk1"][0], gold[f"{tag}.horz.wrk1"], "wrk1")
    th = b.th_tendency[n - 1].numpy().copy()
    o.L.orc_accumulate(C.byref(o.blocks[0].c), _ptr(out["wrk1"][0]), _ptr(th))
    assert_bit_equal(th, gold[f"{tag}.horz.th_tendency"], "th_tendency")
    assert_bit_equal(out["flux_x"][0], gold[f"{tag}.flux_x"], "flux_x")
    assert_bit_equal(out["flux_y"][0], gold[f"{tag}.flux_y"], "flux_y")
    assert_bit_equal(out["flux_z"][0][:, 1:-1, 1:-1], gold[f"{tag}.flux_z"][:, 1:-1, 1:-1], "flux_z")
    assert_bit_equal(out["tracer"][0][:, 4:-4, 4:-4], gold[f"{tag}.tracer_mdppm"][:, 4:-4, 4:-4], "tracer_mdppm")
    assert not gold[f"{tag}.vert.wrk1"].any()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
@pytest.mark.parametrize("tag", ["upwind", "quicker", "quicker_lim", "mdfl_sweby", "dst_linear", "mdfl_sweby_test", "dst_linear_test",
                                 "mdppm_cw84", "mdppm_ifc", "mdppm_sh"])
def test_compute_adv_diss(name, tag):
    """compute_adv_diss (OTA:7547-7712): the scheme applied to the squared tracer, then the dissipation formula"""
    b, gold, _ = load_golden(name)
    _, o = _oracle(b)
    n = int(gold[f"{tag}.tracer"])
    out = o.adv_diss(tag.replace("_lim", ""), [b.T_tau[n - 1].numpy()], [b.tmask_limit[n - 1].numpy()], tag == "quicker_lim",
                     [gold[f"{tag}.advect_tendency"]], [b.rho_dzt.numpy() * 1.01], b.spec.dtime, float(gold[f"{tag}.conversion"]))
    assert_bit_equal(out["t2_tendency"][0], gold[f"{tag}.adv_diss.t2_tendency"], "advection of the squared tracer")
    assert_bit_equal(out["diss"][0], gold[f"{tag}.adv_diss"], "adv_diss")
    assert np.abs(gold[f"{tag}.adv_diss"]).max() > 0


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_quicker_init(name):
    b, gold, _ = load_golden(name)
    _, o = _oracle(b)
    o.quicker_init()
    blk = o.blocks[0]
    for nm in ("quick_x", "quick_y", "curv_xp", "curv_xn", "curv_yp", "curv_yn", "quick_z", "curv_zp", "curv_zn"):
        assert_bit_equal(getattr(blk, nm), gold[f"quicker_init.{nm}"], nm)
    assert_bit_equal(blk.dxt_h2, gold["quicker_init.dxt_quick"], "dxt_quick")
    assert_bit_equal(blk.dyt_h2, gold["quicker_init.dyt_quick"], "dyt_quick")
    assert_bit_equal(blk.tmask_h2, gold["quicker_init.tmask_quick"], "tmask_quick")


@pytest.mark.parametrize("name", GOLDEN_NAMES)
@pytest.mark.parametrize("tag", ["quicker", "quicker_lim", "upwind"])
def test_horz_vert_arms(name, tag):
    b, gold, _ = load_golden(name)
    _, o = _oracle(b)
    n = int(gold[f"{tag}.tracer"])
    Tm1, Tt, tl = b.T[n - 1].numpy(), b.T_tau[n - 1].numpy(), b.tmask_limit[n - 1].numpy()
    if tag == "upwind":
        h = o.horz_upwind([Tm1])
        v = o.vert_upwind([Tm1])
    else:
        h = o.horz_quicker([Tm1], [Tt], [tl], limit_with_upwind=(tag == "quicker_lim"))
        v = o.vert_quicker([Tm1], [Tt], [tl])
    assert_bit_equal(h["wrk1"][0], gold[f"{tag}.horz.wrk1"], "horz wrk1")
    assert_bit_equal(v["wrk1"][0], gold[f"{tag}.vert.wrk1"], "vert wrk1")
    # fluxes: compare where the reference defines them (OTA:2260,2271 / 2572,2591)
    fx, fy, fz = h["flux_x"][0], h["flux_y"][0], v["flux_z"][0]
    assert_bit_equal(fx[:, 1:-1, 0:-1], gold[f"{tag}.flux_x"][:, 1:-1, 0:-1], "flux_x")
    assert_bit_equal(fy[:, 0:-1, 1:-1], gold[f"{tag}.flux_y"][:, 0:-1, 1:-1], "flux_y")
    assert_bit_equal(fz[:, 1:-1, 1:-1], gold[f"{tag}.flux_z"][:, 1:-1, 1:-1], "flux_z")
    import ctypes as C
    from oracle.oracle import _ptr
    th = b.th_tendency[n - 1].numpy().copy()
    o.L.orc_accumulate(C.byref(o.blocks[0].c), _ptr(h["wrk1"][0]), _ptr(th))
    assert_bit_equal(th, gold[f"{tag}.horz.th_tendency"], "th after horz")
    o.L.orc_accumulate(C.byref(o.blocks[0].c), _ptr(v["wrk1"][0]), _ptr(th))
    assert_bit_equal(th, gold[f"{tag}.vert.th_tendency"], "th after vert")


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_continuity(name):
    """diverge_t + wrho_bt recurrence (ocean_advection_velocity.F90:660-669, BDX_ET / BDY_NT) vs the reference text"""
    b, gold, _ = load_golden(name)
    _, o = _oracle(b)
    w0 = np.zeros_like(gold["continuity.wrho_bt"])
    w0[0] = gold["continuity.wrho_bt"][0]                  # wrho_bt(:,:,0) = -(pme + river) is the caller's
    w, div = o.continuity(0, b.uhrho_et.numpy(), b.vhrho_nt.numpy(), w0, gold["continuity.in.tend"], gold["continuity.in.src"])
    assert_bit_equal(div, gold["continuity.diverge_t"], "diverge_t")
    assert_bit_equal(w, gold["continuity.wrho_bt"], "wrho_bt")
