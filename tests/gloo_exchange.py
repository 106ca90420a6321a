"""Test infrastructure (CPU gloo ranks): halo exchange of halo-`halo` fields over torch.distributed point-to-point ops, driven by
``Decomposition.exchange_plan`` -- the same message lists the CUDA library hands to ncclSend/ncclRecv
(mom5_b200/csrc/capi.cu:halo_update).  Backend-agnostic: gloo on CPU tensors (used by the world_size-2 CPU tests
of the multi-rank logic) or nccl on CUDA tensors.

Replaces, for this path only, FMS mpp_update_domains(XUPDATE / YUPDATE) (OTA:4213-4240, 4302-4343).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist

from mom5_b200.domain import Decomposition


def _rect(f: torch.Tensor, m, halo: int) -> torch.Tensor:
    return f[..., m.j0 - 1 + halo:m.j1 + halo, m.i0 - 1 + halo:m.i1 + halo]


def halo_exchange(dec: Decomposition, rank: int, fields: Sequence[torch.Tensor], flags: int, halo: int = 2) -> None:
    """In-place halo update of every tensor in ``fields`` (shape (nk, nj+2*halo, ni+2*halo)).
    All fields travel in ONE message per peer (the reference's `complete=` aggregation, OTA:4218-4219)."""
    sends, recvs = dec.exchange_plan(rank, flags, halo)
    # local (self) strips: cyclic wrap / fold onto this rank
    self_s = [m for m in sends if m.peer == rank]
    self_r = [m for m in recvs if m.peer == rank]
    assert len(self_s) == len(self_r)
    staged = [[_rect(f, s, halo).clone() for f in fields] for s in self_s]
    peers_s = sorted({m.peer for m in sends if m.peer != rank})
    peers_r = sorted({m.peer for m in recvs if m.peer != rank})
    ops, rbufs = [], {}
    for p in peers_s:
        buf = torch.cat([_rect(f, m, halo).reshape(-1) for m in sends if m.peer == p for f in fields])
        ops.append(dist.P2POp(dist.isend, buf.contiguous(), p))
    for p in peers_r:
        n = sum(m.count2d for m in recvs if m.peer == p) * sum(f.shape[0] for f in fields)
        rbufs[p] = torch.empty(n, dtype=fields[0].dtype, device=fields[0].device)
        ops.append(dist.P2POp(dist.irecv, rbufs[p], p))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for r, blks in zip(self_r, staged):
        for f, blk in zip(fields, blks):
            _rect(f, r, halo).copy_(blk.flip(-1, -2) if r.flip else blk)
    for p in peers_r:
        off = 0
        for m in recvs:
            if m.peer != p:
                continue
            for f in fields:
                dst = _rect(f, m, halo)
                n = dst.numel()
                blk = rbufs[p][off:off + n].reshape(dst.shape)
                dst.copy_(blk.flip(-1, -2) if m.flip else blk)
                off += n
