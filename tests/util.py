"""Shared helpers for the test-suite (test infrastructure)."""
from __future__ import annotations

import os

import numpy as np
import torch

from mom5_b200.synthetic import BlockInputs, CaseSpec

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_NAMES = ["g_tripolar", "g_walls", "g_torus", "g_walls_rough"]


def bits(a) -> np.ndarray:
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def assert_bit_equal(a, b, what=""):
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    b = b.detach().cpu().numpy() if hasattr(b, "detach") else np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    ne = bits(a) != bits(b)
    if ne.any():
        idx = np.argwhere(ne)
        q = tuple(idx[0])
        raise AssertionError(f"{what}: {ne.sum()} of {ne.size} elements differ bitwise; first at {q}: "
                             f"{a[q]!r} vs {b[q]!r} (max abs diff {np.abs(a - b).max():.3e})")


def load_golden(name):
    """-> (BlockInputs for the single global domain, dict of golden outputs, cite string)"""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = lambda k: z["in." + k]
    ni, nj, nk, ntr = int(g("ni")), int(g("nj")), int(g("nk")), int(g("ntr"))
    spec = CaseSpec(name, ni, nj, nk, ntr, cyclic_x=bool(g("cyclic_x")), cyclic_y=bool(g("cyclic_y")),
                    tripolar=bool(g("tripolar")), dtime=float(g("dtime")), flow_scale=1.0)
    t = lambda a: torch.from_numpy(np.array(a, dtype=np.float64, copy=True))
    grid2d = {k: t(g("grid." + k)) for k in ("dat", "datr", "dxt", "dyt", "dxte", "dyte", "dxtn", "dytn")}
    b = BlockInputs(spec, 1, ni, 1, nj, grid2d, t(g("dzt")), t(g("tmask")), t(g("rho_dzt")), t(g("uhrho_et")),
                    t(g("vhrho_nt")), t(g("wrho_bt")))
    for n in range(1, ntr + 1):
        b.T.append(t(g(f"T.{n}")))
        b.T_tau.append(t(g(f"T_tau.{n}")))
        b.th_tendency.append(t(g(f"th0.{n}")))
        b.tmask_limit.append(t(g(f"tmask_limit.{n}")))
    out = {k[4:]: z[k] for k in z.files if k.startswith("out.")}
    return b, out, str(z["__cites__"])
