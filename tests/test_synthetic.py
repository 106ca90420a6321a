"""The synthetic input generator: reproducible, pointwise (any block == the same block cut from the global one)."""
import numpy as np
import torch

from mom5_b200.synthetic import make_case, splitmix64_uniform, splitmix64_uniform_py


def test_splitmix_matches_pure_python():
    idx = torch.tensor([0, 1, 2, 12345, 2 ** 40 + 7, 2 ** 62 + 11], dtype=torch.int64)
    for stream in (0, 5, 1293):
        u = splitmix64_uniform(idx, stream)
        for q, i in enumerate(idx.tolist()):
            assert u[q].item() == splitmix64_uniform_py(i, stream)
    assert (splitmix64_uniform(torch.arange(10000), 3) < 1.0).all()


def test_block_generation_equals_slicing_the_global_block():
    g = make_case("mini_tripolar")
    gb = g.block()
    for (i0, i1, j0, j1) in [(1, 20, 1, 15), (21, 40, 16, 30), (11, 30, 9, 22)]:
        direct = g.block(i0, i1, j0, j1)
        cut = gb.sub_block(i0, i1, j0, j1)
        for nm in ("tmask", "rho_dzt", "uhrho_et", "vhrho_nt"):
            assert torch.equal(getattr(direct, nm), getattr(cut, nm)), nm
        # w: identical on everything the schemes read (compute domain); the halo ring of w is never read
        assert torch.equal(direct.wrho_bt[:, 1:-1, 1:-1], cut.wrho_bt[:, 1:-1, 1:-1])
        for n in range(len(gb.T)):
            assert torch.equal(direct.T[n], cut.T[n]) and torch.equal(direct.th_tendency[n], cut.th_tendency[n])
        for k in gb.grid2d:
            assert torch.equal(direct.grid2d[k], cut.grid2d[k]), k


def test_inputs_are_sane():
    for name in ("mini_tripolar", "mini_walls", "mini_torus", "box1"):
        g = make_case(name)
        b = g.block()
        assert (b.rho_dzt > 0).all(), "rho_dzt > 0 everywhere, land and halos included (ocean_thickness.F90:1218)"
        assert set(np.unique(b.tmask.numpy())) <= {0.0, 1.0}
        for t in b.T + [b.uhrho_et, b.vhrho_nt, b.wrho_bt]:
            assert torch.isfinite(t).all()
        dt = g.s.dtime
        cz = (b.wrho_bt[1:, 1:-1, 1:-1] * dt / b.rho_dzt[:, 1:-1, 1:-1]).abs().max().item()
        cx = (b.uhrho_et[:, 1:-1, :-1] * dt * 2 / ((b.rho_dzt[:, 1:-1, :-1] + b.rho_dzt[:, 1:-1, 1:]) * b.grid2d["dxte"][1:-1, :-1])).abs().max().item()
        assert max(cx, cz) <= g.s.cfl * 1.0001
    g = make_case("mini_tripolar")
    b = g.block()
    land = 1.0 - b.tmask[0, 1:-1, 1:-1].mean().item()
    assert 0.1 < land < 0.6
    # bottom mass flux vanishes to round-off: the face profiles have zero vertical sum (conservation needs it)
    kmt = b.tmask.sum(0).long()
    wbot = torch.gather(b.wrho_bt, 0, kmt[None])[0][1:-1, 1:-1]
    assert wbot.abs().max().item() < 1e-9 * b.wrho_bt.abs().max().item()
    # fold consistency of the top-row northward transport: V(i,nj) = -V(ni+1-i,nj)
    V = (b.grid2d["dxtn"] * b.vhrho_nt)[:, -2, 1:-1]
    assert torch.equal(V, -V.flip(-1))
