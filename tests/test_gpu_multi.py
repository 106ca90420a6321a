"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): N processes, one per GPU, NCCL halo exchange inside
libmom5adv.so.  Every rank generates only its block; the gathered th_tendency / adv_tendency must be bit-identical
to the single-domain oracle for every layout (the reference's PE-count invariance)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, over, px, py, outdir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from mom5_b200.api import ADVECT_MDFL_SWEBY, ADVECT_MDFL_SWEBY_TEST, ADVECT_MDPPM, ADVECT_QUICKER, Communicator, TracerAdvect
    from mom5_b200.synthetic import make_case
    g = make_case(case, **over)
    dec = g.s.decomposition(px, py)
    i0, i1, j0, j1 = dec.extent(rank)

    def rmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g.calibrate(i0, i1, j0, j1, reduce_max=rmax)
    b = g.block(i0, i1, j0, j1, with_tau=True)
    comm = Communicator.create_from_torch_distributed()
    adv = TracerAdvect(b, dec=dec, rank=rank, ntracers_max=len(b.T), comm=comm)
    T = [t.cuda() for t in b.T]
    th = [t.cuda().clone() for t in b.th_tendency]
    out = [torch.empty_like(t) for t in T]
    u, v, w, rho = b.uhrho_et.cuda(), b.vhrho_nt.cuda(), b.wrho_bt.cuda(), b.rho_dzt.cuda()
    adv.advect_tracer_sweby_all(T, th, out, u, v, w, rho, g.s.dtime)
    # single-tracer arms that exchange halos: mdfl_sweby (X then Y) and quicker (full update + fold line)
    th1 = b.th_tendency[0].cuda().clone()
    w1 = torch.empty_like(th1)
    adv.horz_advect_tracer(ADVECT_MDFL_SWEBY, T[0], th1, w1, u, v, g.s.dtime, wrho_bt=w, rho_dzt=rho)
    res = dict(ext=np.array([i0, i1, j0, j1]), scale=np.array(g.s.flow_scale), mdfl=w1.cpu().numpy())
    th2 = b.th_tendency[0].cuda().clone()   # quicker: full halo-2 update + (tripolar) the fold-line fix across ranks
    w2 = torch.empty_like(th2)
    adv.horz_advect_tracer(ADVECT_QUICKER, T[0], th2, w2, u, v, T_tau=b.T_tau[0].cuda(), tmask_limit=b.tmask_limit[0].cuda())
    res["quicker"] = w2.cpu().numpy()
    th3 = b.th_tendency[0].cuda().clone()   # mass-weighted variant: three fields exchanged per update
    w3 = torch.empty_like(th3)
    adv.horz_advect_tracer(ADVECT_MDFL_SWEBY_TEST, T[0], th3, w3, u, v, g.s.dtime, wrho_bt=w, rho_dzt=rho)
    res["sweby_test"] = w3.cpu().numpy()
    th4 = b.th_tendency[0].cuda().clone()   # MDPPM: halo-4 strips
    w4 = torch.empty_like(th4)
    adv.set_ppm_limiters(3)
    adv.horz_advect_tracer(ADVECT_MDPPM, T[0], th4, w4, u, v, g.s.dtime, wrho_bt=w, rho_dzt=rho)
    res["mdppm"] = w4.cpu().numpy()
    torch.cuda.synchronize()
    for n in range(len(T)):
        res[f"th{n}"] = th[n].cpu().numpy()
        res[f"adv{n}"] = out[n].cpu().numpy()
    np.savez(os.path.join(outdir, f"r{rank}.npz"), **res)
    dist.barrier()
    adv.close()
    comm.destroy()
    dist.destroy_process_group()


def _check(tmp_path, case, over, px, py, fuse="1"):
    import torch.multiprocessing as mp
    os.environ["MOM5ADV_FUSE"] = fuse     # read by mom5adv_init in the spawned workers (they inherit the environment)
    from mom5_b200.synthetic import make_case
    from oracle.oracle import Oracle
    world = px * py
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mp.spawn(_worker, args=(world, _free_port(), case, over, px, py, str(tmp_path)), nprocs=world, join=True)
    g = make_case(case, **over)
    gb = g.block(with_tau=True)
    o = Oracle(g.s.decomposition(1, 1), [gb])
    th = [[t.numpy().copy() for t in gb.th_tendency]]
    ref = o.sweby_all([[t.numpy() for t in gb.T]], th, g.s.dtime)
    mdfl = o.mdfl_sweby([gb.T[0].numpy()], g.s.dtime, 1.0)["wrk1"][0]
    ppm = o.mdppm([gb.T[0].numpy()], g.s.dtime, 3)["wrk1"][0]
    stest = o.sweby_test([gb.T[0].numpy()], g.s.dtime, 1.0)["wrk1"][0]
    quick = o.horz_quicker([gb.T[0].numpy()], [gb.T_tau[0].numpy()], [gb.tmask_limit[0].numpy()], False)["wrk1"][0]
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert float(z["scale"]) == g.s.flow_scale
        i0, i1, j0, j1 = z["ext"]
        cmp = [(f"th{n}", th[0][n]) for n in range(len(gb.T))] + [(f"adv{n}", ref["adv"][0][n]) for n in range(len(gb.T))]
        cmp.append(("mdfl", mdfl))
        cmp.append(("sweby_test", stest))
        cmp.append(("mdppm", ppm))
        if "quicker" in z.files:
            cmp.append(("quicker", quick))
        for nm, full in cmp:
            got = z[nm][:, 1:-1, 1:-1]
            want = full[:, j0:j1 + 1, i0:i1 + 1]
            assert np.array_equal(got.view(np.int64), want.view(np.int64)), (case, px, py, r, nm)


@pytest.mark.parametrize("case,over,px,py", [
    ("mini_tripolar", {}, 2, 1), ("mini_tripolar", {}, 1, 2), ("mini_walls", {}, 2, 1), ("mini_torus", {}, 1, 2),
    ("global_1deg", dict(ntr=5), 2, 1), ("global_1deg", dict(ntr=5), 1, 2),
    # wide / tall enough for the comm-compute overlap branches of both drivers (>= 4 z tiles per rank; interior j-chunks)
    ("global_1deg", dict(ni=1040, nj=80, nk=6, ntr=3), 2, 1), ("global_1deg", dict(ni=1040, nj=80, nk=6, ntr=3), 1, 2)])
def test_two_gpus(tmp_path, case, over, px, py):
    _check(tmp_path, case, over, px, py)


@pytest.mark.parametrize("case,over,px,py", [("mini_tripolar", {}, 2, 1), ("global_1deg", dict(ni=1040, nj=80, nk=6, ntr=3), 1, 2)])
def test_two_gpus_three_sweep_driver(tmp_path, case, over, px, py):
    """the second Sweby driver (MOM5ADV_FUSE=0: z, x, y as separate sweeps), incl. its comm-compute overlap branch"""
    try:
        _check(tmp_path, case, over, px, py, fuse="0")
    finally:
        os.environ.pop("MOM5ADV_FUSE", None)


@pytest.mark.parametrize("case,over,px,py", [("mini_tripolar", {}, 2, 2), ("mini_tripolar", {}, 1, 4), ("global_1deg", dict(ntr=3), 2, 2)])
def test_four_gpus(tmp_path, case, over, px, py):
    _check(tmp_path, case, over, px, py)


def test_eight_gpus(tmp_path):
    _check(tmp_path, "global_1deg", dict(ntr=3), 2, 4)
