"""compute-sanitizer driver (run on the GPU box):  compute-sanitizer --tool memcheck|racecheck|initcheck python tests/sanitize_run.py

Small cases through every Sweby driver / staging flavour and the other dispatcher arms, tracer counts 1..5 (template instances NT = 1..4 and
the 3+2 split), each compared with the CPU oracle so that a sanitizer-clean run is also a correct one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mom5_b200.api import ADVECT_MDFL_SWEBY, ADVECT_MDPPM, ADVECT_QUICKER, ADVECT_UPWIND, TracerAdvect
from mom5_b200.synthetic import make_case
from oracle.oracle import Oracle


def one(case, over, fuse, tma):
    os.environ["MOM5ADV_FUSE"], os.environ["MOM5ADV_TMA"] = fuse, tma
    g = make_case(case, **over)
    b = g.block(with_tau=True)
    o = Oracle(g.s.decomposition(1, 1), [b])
    th_ref = [[t.numpy().copy() for t in b.th_tendency]]
    ref = o.sweby_all([[t.numpy() for t in b.T]], th_ref, g.s.dtime)
    adv = TracerAdvect(b, ntracers_max=len(b.T), limit_with_upwind=True)
    T = [t.cuda() for t in b.T]
    th = [t.cuda().clone() for t in b.th_tendency]
    out = [torch.empty_like(t) for t in T]
    u, v, w, rho = b.uhrho_et.cuda(), b.vhrho_nt.cuda(), b.wrho_bt.cuda(), b.rho_dzt.cuda()
    adv.advect_tracer_sweby_all(T, th, out, u, v, w, rho, g.s.dtime)
    torch.cuda.synchronize()
    bad = 0
    for n in range(len(T)):
        bad += int((th[n].cpu().numpy().view(np.int64) != th_ref[0][n].view(np.int64)).sum())
        bad += int((out[n].cpu().numpy().view(np.int64) != ref["adv"][0][n].view(np.int64)).sum())
    # the single-tracer arms
    for scheme in (ADVECT_MDFL_SWEBY, ADVECT_QUICKER, ADVECT_UPWIND, ADVECT_MDPPM):
        t1 = b.th_tendency[0].cuda().clone()
        w1 = torch.empty_like(t1)
        adv.horz_advect_tracer(scheme, T[0], t1, w1, u, v, g.s.dtime, T_tau=b.T_tau[0].cuda(), tmask_limit=b.tmask_limit[0].cuda(), wrho_bt=w, rho_dzt=rho)
        if scheme in (ADVECT_QUICKER, ADVECT_UPWIND):
            adv.vert_advect_tracer(scheme, T[0], t1, w1, w, T_tau=b.T_tau[0].cuda(), tmask_limit=b.tmask_limit[0].cuda())
    # host-pointer entry (banded / tracer pipelines)
    hth = [t.numpy().copy() for t in b.th_tendency]
    hout = [np.empty_like(t) for t in hth]
    adv.advect_tracer_sweby_all([t.numpy() for t in b.T], hth, hout, b.uhrho_et.numpy(), b.vhrho_nt.numpy(), b.wrho_bt.numpy(), b.rho_dzt.numpy(), g.s.dtime)
    for n in range(len(T)):
        bad += int((hth[n].view(np.int64) != th_ref[0][n].view(np.int64)).sum())
    torch.cuda.synchronize()
    adv.close()
    print(f"SANITIZE case={case} over={over} fuse={fuse} tma={tma} ntr={len(T)} mismatches={bad}", flush=True)
    return bad


if __name__ == "__main__":
    total = 0
    for fuse, tma in (("1", "3"), ("1", "0"), ("0", "3")):
        for case, over in (("mini_tripolar", {}), ("mini_tripolar", dict(ntr=1)), ("mini_tripolar", dict(ntr=4)), ("mini_tripolar", dict(ntr=5)),
                           ("mini_walls", {}), ("mini_torus", {}), ("global_1deg", dict(ni=136, nj=40, nk=8, ntr=2))):
            total += one(case, over, fuse, tma)
    print("SANITIZE total mismatches", total)
    sys.exit(1 if total else 0)
