"""diagnostic: one tiny sweby_all through the selected staging, compared with the oracle (run under compute-sanitizer on the GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mom5_b200.api import TracerAdvect
from mom5_b200.synthetic import make_case
from oracle.oracle import Oracle
case = sys.argv[1] if len(sys.argv) > 1 else "mini_tripolar"
g = make_case(case)
b = g.block()
o = Oracle(g.s.decomposition(1, 1), [b])
th_ref = [[t.numpy().copy() for t in b.th_tendency]]
ref = o.sweby_all([[t.numpy() for t in b.T]], th_ref, g.s.dtime)
adv = TracerAdvect(b, ntracers_max=len(b.T))
T = [t.cuda() for t in b.T]
th = [t.cuda().clone() for t in b.th_tendency]
out = [torch.empty_like(t) for t in T]
adv.advect_tracer_sweby_all(T, th, out, b.uhrho_et.cuda(), b.vhrho_nt.cuda(), b.wrho_bt.cuda(), b.rho_dzt.cuda(), g.s.dtime)
torch.cuda.synchronize()
bad = 0
for n in range(len(T)):
    bad += int((th[n].cpu().numpy().view(np.int64) != th_ref[0][n].view(np.int64)).sum())
    bad += int((out[n].cpu().numpy().view(np.int64) != ref["adv"][0][n].view(np.int64)).sum())
print("DIAG", case, "TMA", os.environ.get("MOM5ADV_TMA"), "FUSE", os.environ.get("MOM5ADV_FUSE"), "mismatching elements:", bad, adv.kernel_launches())
adv.close()
