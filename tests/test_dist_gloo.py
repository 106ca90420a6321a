"""world_size-2 (and 4) multi-process test of the N>1 host logic on CPU with the gloo backend:
every rank generates ONLY its own block, runs the per-block sweeps, and exchanges halos with
mom5_b200.exchange.halo_exchange (the message plan the CUDA library executes over NCCL).  The gathered result
must be bit-identical to the single-domain answer (the reference's PE-count invariance)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, px, py, outdir):
    import ctypes as C
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mom5_b200.domain import XUPDATE, YUPDATE
    from tests.gloo_exchange import halo_exchange
    from mom5_b200.synthetic import make_case
    from oracle.oracle import Block, _pp, _ptr, lib
    L = lib()
    g = make_case(case)
    dec = g.s.decomposition(px, py)
    i0, i1, j0, j1 = dec.extent(rank)
    # flow scale: cross-rank max, as bench.py does it
    def rmax(x):
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    g.calibrate(i0, i1, j0, j1, reduce_max=rmax)
    b = g.block(i0, i1, j0, j1)
    blk = Block(b)
    ntr = len(b.T)
    # mdfl_init: mask with a full halo-2 update
    L.orc_mdfl_init_mask(C.byref(blk.c))
    tm_mask = torch.from_numpy(blk.tmask_h2)
    halo_exchange(dec, rank, [tm_mask], XUPDATE | YUPDATE)
    T = [t.numpy() for t in b.T]
    th = [t.numpy().copy() for t in b.th_tendency]
    tm = [blk.h2() for _ in range(ntr)]
    adv = [blk.d1() for _ in range(ntr)]
    u, v, w, rho = (x.numpy() for x in (b.uhrho_et, b.vhrho_nt, b.wrho_bt, b.rho_dzt))
    dt = C.c_double(g.s.dtime)
    nul = C.POINTER(C.POINTER(C.c_double))()
    L.orc_sweby_all_z(C.byref(blk.c), ntr, dt, _pp(T), _ptr(w), _ptr(rho), _pp(tm), nul, nul)
    halo_exchange(dec, rank, [torch.from_numpy(t) for t in tm], XUPDATE)
    L.orc_sweby_all_x(C.byref(blk.c), ntr, dt, _pp(T), _ptr(u), _ptr(rho), _pp(tm), nul, nul)
    halo_exchange(dec, rank, [torch.from_numpy(t) for t in tm], YUPDATE)
    L.orc_sweby_all_y(C.byref(blk.c), ntr, dt, _pp(T), _ptr(u), _ptr(v), _ptr(w), _ptr(rho), _pp(tm), _pp(th), _pp(adv), nul, nul)
    np.savez(os.path.join(outdir, f"r{rank}.npz"), ext=np.array([i0, i1, j0, j1]), scale=np.array(g.s.flow_scale),
             **{f"th{n}": th[n] for n in range(ntr)}, **{f"adv{n}": adv[n] for n in range(ntr)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,px,py", [("mini_tripolar", 2, 1), ("mini_tripolar", 1, 2), ("mini_walls", 2, 1),
                                         ("mini_torus", 1, 2), ("mini_tripolar", 2, 2)])
def test_two_rank_gloo_sweby_all_is_bit_identical_to_single_domain(tmp_path, case, px, py):
    from mom5_b200.synthetic import make_case
    from oracle.oracle import Oracle
    world = px * py
    mp.spawn(_worker, args=(world, _free_port(), case, px, py, str(tmp_path)), nprocs=world, join=True)
    g = make_case(case)
    gb = g.block()
    o = Oracle(g.s.decomposition(1, 1), [gb])
    th = [[t.numpy().copy() for t in gb.th_tendency]]
    ref = o.sweby_all([[t.numpy() for t in gb.T]], th, g.s.dtime)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert float(z["scale"]) == g.s.flow_scale
        i0, i1, j0, j1 = z["ext"]
        for n in range(len(gb.T)):
            for nm, full in ((f"th{n}", th[0][n]), (f"adv{n}", ref["adv"][0][n])):
                got = z[nm][:, 1:-1, 1:-1]
                want = full[:, j0:j1 + 1, i0:i1 + 1]
                assert np.array_equal(got.view(np.int64), want.view(np.int64)), (case, px, py, r, nm)


def _worker_f3(rank, world, port, case, px, py, outdir):
    """SURVEY section 8f row 3 on N ranks: advect_tracer_mdfl_sweby_test (three halo-2 fields per exchange) and
    advect_tracer_mdppm (halo-4 mask and tracer), every rank holding only its block"""
    import ctypes as C
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mom5_b200.domain import XUPDATE, YUPDATE
    from tests.gloo_exchange import halo_exchange
    from mom5_b200.synthetic import make_case
    from oracle.oracle import Block, _ptr, lib
    L = lib()
    g = make_case(case)
    dec = g.s.decomposition(px, py)
    i0, i1, j0, j1 = dec.extent(rank)

    def rmax(x):
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g.calibrate(i0, i1, j0, j1, reduce_max=rmax)
    b = g.block(i0, i1, j0, j1)
    blk = Block(b)
    T = b.T[0].numpy()
    u, v, w, rho = (x.numpy() for x in (b.uhrho_et, b.vhrho_nt, b.wrho_bt, b.rho_dzt))
    dt, one = C.c_double(g.s.dtime), C.c_double(1.0)
    ex = lambda fields, flags, halo: halo_exchange(dec, rank, [torch.from_numpy(f) for f in fields], flags, halo=halo)
    # ---- sweby_test ----
    L.orc_mdfl_init_mask(C.byref(blk.c))
    ex([blk.tmask_h2], XUPDATE | YUPDATE, 2)
    tr, tms, ms = blk.h2(), blk.h2(), blk.h2()
    fx, fy, fz, w_st = blk.d1(), blk.d1(), blk.d1(), blk.d1()
    L.orc_sweby_test_z(C.byref(blk.c), dt, one, _ptr(T), _ptr(w), _ptr(rho), _ptr(tr), _ptr(tms), _ptr(ms), _ptr(fz))
    ex([tr, tms, ms], XUPDATE, 2)
    L.orc_sweby_test_x(C.byref(blk.c), dt, one, _ptr(u), _ptr(tr), _ptr(tms), _ptr(ms), _ptr(fx))
    ex([tr, tms, ms], YUPDATE, 2)
    L.orc_sweby_test_y(C.byref(blk.c), dt, one, _ptr(T), _ptr(v), _ptr(rho), _ptr(tr), _ptr(tms), _ptr(ms), _ptr(fy), _ptr(w_st))
    # ---- mdppm, limiter 3 ----
    m4 = blk.h4()
    m4[:, 4:-4, 4:-4] = blk.tmask[:, 1:-1, 1:-1]
    ex([m4], XUPDATE | YUPDATE, 4)
    t4 = blk.h4()
    fx, fy, fz, w_pp = blk.d1(), blk.d1(), blk.d1(), blk.d1()
    lim = C.c_int(3)
    L.orc_mdppm_z(C.byref(blk.c), dt, lim, _ptr(T), _ptr(w), _ptr(rho), _ptr(m4), _ptr(t4), _ptr(fz))
    ex([t4], XUPDATE, 4)
    L.orc_mdppm_x(C.byref(blk.c), dt, lim, _ptr(T), _ptr(u), _ptr(rho), _ptr(m4), _ptr(t4), _ptr(fx))
    ex([t4], YUPDATE, 4)
    L.orc_mdppm_y(C.byref(blk.c), dt, lim, _ptr(T), _ptr(u), _ptr(v), _ptr(w), _ptr(rho), _ptr(m4), _ptr(t4), _ptr(fy), _ptr(w_pp))
    np.savez(os.path.join(outdir, f"r{rank}.npz"), ext=np.array([i0, i1, j0, j1]), sweby_test=w_st, mdppm=w_pp)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,px,py", [("mini_tripolar", 2, 1), ("mini_tripolar", 1, 2), ("mini_torus", 2, 2)])
def test_gloo_ranks_sweby_test_and_mdppm_are_bit_identical_to_single_domain(tmp_path, case, px, py):
    from mom5_b200.synthetic import make_case
    from oracle.oracle import Oracle
    world = px * py
    mp.spawn(_worker_f3, args=(world, _free_port(), case, px, py, str(tmp_path)), nprocs=world, join=True)
    g = make_case(case)
    gb = g.block()
    o = Oracle(g.s.decomposition(1, 1), [gb])
    ref = dict(sweby_test=o.sweby_test([gb.T[0].numpy()], g.s.dtime, 1.0)["wrk1"][0],
               mdppm=o.mdppm([gb.T[0].numpy()], g.s.dtime, 3)["wrk1"][0])
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        i0, i1, j0, j1 = z["ext"]
        for nm, full in ref.items():
            got = z[nm][:, 1:-1, 1:-1]
            want = full[:, j0:j1 + 1, i0:i1 + 1]
            assert np.array_equal(got.view(np.int64), want.view(np.int64)), (case, px, py, r, nm)
