"""2-D horizontal domain decomposition for the tracer-advection path.

Host-side mirror of the FMS pieces this path needs (k is never split):

* ``define_layout``   -- mpp_define_layout2D   (src/shared/mpp/include/mpp_domains_define.inc:28-55)
* ``compute_extent``  -- mpp_compute_extent    (mpp_domains_define.inc:187-273), mirror-symmetric split
* ``Decomposition``   -- set_ocean_domain's cyclic / tripolar-fold setup (src/mom5/ocean_core/ocean_domains.F90:176-226)
* ``map_source``      -- where a halo point's value lives: interior neighbour, cyclic wrap, folded north edge
                         ``(i, nj+m) <- (ni+1-i, nj+1-m)`` (mpp_domains_define.inc:4865-4885), or nowhere (solid wall)
* ``exchange_plan``   -- the strips one rank sends / receives for an XUPDATE (E/W) or YUPDATE (N/S) of a
                         halo-2 field (mpp_do_update.h:57-78): the message list handed to NCCL send/recv.

Pure Python / numpy; no GPU, no oracle.  All indices are 1-based global indices, inclusive, as in the reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

XUPDATE = 1
YUPDATE = 2


def _nint(x: float) -> int:
    """Fortran NINT: round half away from zero."""
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def define_layout(ni_g: int, nj_g: int, ndivs: int) -> Tuple[int, int]:
    """mpp_define_layout2D: divide ``ndivs`` ranks in the domain aspect ratio."""
    idiv = _nint(math.sqrt(float(ndivs * ni_g) / nj_g))
    idiv = max(idiv, 1)
    while ndivs % idiv != 0:
        idiv -= 1
    return idiv, ndivs // idiv


def compute_extent(isg: int, ieg: int, ndivs: int) -> Tuple[List[int], List[int]]:
    """mpp_compute_extent without a user extent: mirror-symmetric uneven split."""
    npts = ieg - isg + 1
    even_n, even_p = ndivs % 2 == 0, npts % 2 == 0
    symmetrize = (even_n and even_p) or (not even_n and not even_p) or (not even_n and even_p and ndivs < npts // 2)
    ibegin, iend = [0] * ndivs, [0] * ndivs
    is_, ie, imax, ndmax = isg, 0, ieg, ndivs
    for ndiv in range(ndivs):
        if ndiv < (ndivs - 1) // 2 + 1:
            ie = is_ + int(math.ceil(float(imax - is_ + 1) / (ndmax - ndiv))) - 1
            ndmirror = (ndivs - 1) - ndiv
            if ndmirror > ndiv and symmetrize:
                ibegin[ndmirror] = max(isg + ieg - ie, ie + 1)
                iend[ndmirror] = max(isg + ieg - is_, ie + 1)
                imax = ibegin[ndmirror] - 1
                ndmax -= 1
        else:
            if symmetrize:
                is_, ie = ibegin[ndiv], iend[ndiv]
            else:
                ie = is_ + int(math.ceil(float(imax - is_ + 1) / (ndmax - ndiv))) - 1
        ibegin[ndiv], iend[ndiv] = is_, ie
        if ie < is_:
            raise ValueError("compute_extent: domain extents must be positive definite")
        if ndiv == ndivs - 1 and iend[ndiv] != ieg:
            raise ValueError("compute_extent: domain extents do not span space completely")
        is_ = ie + 1
    return ibegin, iend


@dataclass
class Message:
    """One halo strip: ``peer`` rank, and the rectangle (1-based LOCAL indices of the halo-`halo` field,
    compute domain = 1..ni) read on the sender / written on the receiver.  ``flip`` = folded north edge:
    the receiver's rectangle is filled in reversed i AND reversed j order relative to the sender's."""
    peer: int
    i0: int
    i1: int
    j0: int
    j1: int
    flip: bool = False

    @property
    def count2d(self) -> int:
        return (self.i1 - self.i0 + 1) * (self.j1 - self.j0 + 1)


@dataclass
class Decomposition:
    ni_g: int
    nj_g: int
    px: int
    py: int
    cyclic_x: bool = False
    cyclic_y: bool = False
    tripolar: bool = False  # folded north edge
    ibeg: List[int] = field(default_factory=list)
    iend: List[int] = field(default_factory=list)
    jbeg: List[int] = field(default_factory=list)
    jend: List[int] = field(default_factory=list)

    def __post_init__(self):
        if not self.ibeg:
            self.ibeg, self.iend = compute_extent(1, self.ni_g, self.px)
        if not self.jbeg:
            self.jbeg, self.jend = compute_extent(1, self.nj_g, self.py)
        if self.tripolar and self.cyclic_y:
            raise ValueError("tripolar fold and cyclic_y are mutually exclusive")

    # ---- ranks <-> blocks (rank = ix + px*iy, x fastest, as in the FMS pelist order) ----
    @property
    def nranks(self) -> int:
        return self.px * self.py

    def coords(self, rank: int) -> Tuple[int, int]:
        return rank % self.px, rank // self.px

    def rank_of(self, ix: int, iy: int) -> int:
        return ix + self.px * iy

    def extent(self, rank: int) -> Tuple[int, int, int, int]:
        ix, iy = self.coords(rank)
        return self.ibeg[ix], self.iend[ix], self.jbeg[iy], self.jend[iy]

    def local_size(self, rank: int) -> Tuple[int, int]:
        i0, i1, j0, j1 = self.extent(rank)
        return i1 - i0 + 1, j1 - j0 + 1

    # ---- where does the value of global point (ig, jg) live? ----
    def map_source(self, ig: int, jg: int) -> Optional[Tuple[int, int]]:
        if jg > self.nj_g:
            if self.tripolar:
                jg = 2 * self.nj_g + 1 - jg
                ig = self.ni_g + 1 - ig
            elif self.cyclic_y:
                jg -= self.nj_g
            else:
                return None
        elif jg < 1:
            if self.cyclic_y:
                jg += self.nj_g
            else:
                return None
        if ig < 1:
            if not self.cyclic_x:
                return None
            ig += self.ni_g
        elif ig > self.ni_g:
            if not self.cyclic_x:
                return None
            ig -= self.ni_g
        if not (1 <= ig <= self.ni_g and 1 <= jg <= self.nj_g):
            return None
        return ig, jg

    def map_source_arrays(self, ig: np.ndarray, jg: np.ndarray):
        """Vectorised map_source. Returns (igs, jgs, valid); invalid points are clamped into range."""
        ig = np.array(ig, dtype=np.int64, copy=True)
        jg = np.array(jg, dtype=np.int64, copy=True)
        ig, jg = np.broadcast_arrays(ig, jg)
        ig, jg = ig.copy(), jg.copy()
        valid = np.ones(ig.shape, dtype=bool)
        north = jg > self.nj_g
        if self.tripolar:
            ig = np.where(north, self.ni_g + 1 - ig, ig)
            jg = np.where(north, 2 * self.nj_g + 1 - jg, jg)
        elif self.cyclic_y:
            jg = np.where(north, jg - self.nj_g, jg)
        else:
            valid &= ~north
        south = jg < 1
        if self.cyclic_y:
            jg = np.where(south, jg + self.nj_g, jg)
        else:
            valid &= ~south
        if self.cyclic_x:
            ig = np.where(ig < 1, ig + self.ni_g, ig)
            ig = np.where(ig > self.ni_g, ig - self.ni_g, ig)
        valid &= (ig >= 1) & (ig <= self.ni_g) & (jg >= 1) & (jg <= self.nj_g)
        return np.clip(ig, 1, self.ni_g), np.clip(jg, 1, self.nj_g), valid

    def owner(self, ig: int, jg: int) -> int:
        ix = next(d for d in range(self.px) if self.ibeg[d] <= ig <= self.iend[d])
        iy = next(d for d in range(self.py) if self.jbeg[d] <= jg <= self.jend[d])
        return self.rank_of(ix, iy)

    # ---- message lists for one update of a halo-`halo` field ----
    def exchange_plan(self, rank: int, flags: int, halo: int = 2):
        """Return (sends, recvs) for ``rank``.

        recvs[n] is filled from the peer's sends entry that names ``rank`` with the same strip; both lists are
        ordered identically on the two sides (by direction, then by source rank), so a grouped
        send/recv matches them pairwise.  Self-messages (peer == rank: cyclic wrap or fold onto the same rank)
        are included; the caller serves them with a local copy kernel.
        """
        sends: List[Message] = []
        recvs: List[Message] = []
        for r in range(self.nranks):
            for m_recv, m_send in self._recv_strips(r, flags, halo):
                if r == rank:
                    recvs.append(m_recv)
                if m_recv.peer == rank:
                    sends.append(Message(peer=r, i0=m_send.i0, i1=m_send.i1, j0=m_send.j0, j1=m_send.j1, flip=m_send.flip))
        return sends, recvs

    def _recv_strips(self, rank: int, flags: int, halo: int):
        """Yield (recv Message on `rank`, matching send Message on the peer) for every maximal rectangle of
        `rank`'s halo that has a single source rank, direction by direction."""
        i0g, i1g, j0g, j1g = self.extent(rank)
        ni, nj = i1g - i0g + 1, j1g - j0g + 1
        regions = []
        if flags & XUPDATE:
            regions += [(1 - halo, 0, 1, nj), (ni + 1, ni + halo, 1, nj)]  # W, E
        if flags & YUPDATE:
            regions += [(1, ni, 1 - halo, 0), (1, ni, nj + 1, nj + halo)]  # S, N
        if (flags & XUPDATE) and (flags & YUPDATE):
            regions += [(1 - halo, 0, 1 - halo, 0), (ni + 1, ni + halo, 1 - halo, 0),
                        (1 - halo, 0, nj + 1, nj + halo), (ni + 1, ni + halo, nj + 1, nj + halo)]
        for (ia, ib, ja, jb) in regions:
            # split the region into runs of constant (owner, orientation)
            il = np.arange(ia, ib + 1)
            jl = np.arange(ja, jb + 1)
            IG, JG = np.meshgrid(il + i0g - 1, jl + j0g - 1, indexing="xy")  # [j, i]
            igs, jgs, valid = self.map_source_arrays(IG, JG)
            if not valid.any():
                continue
            # the mapping is separable and monotone along each axis within a region -> rectangles by
            # splitting at owner changes along i (row 0) and along j (column 0)
            own_x = np.array([next(d for d in range(self.px) if self.ibeg[d] <= g <= self.iend[d]) for g in igs[0, :]])
            own_y = np.array([next(d for d in range(self.py) if self.jbeg[d] <= g <= self.jend[d]) for g in jgs[:, 0]])
            vx, vy = valid[0, :], valid[:, 0]
            flip = bool(self.tripolar and (JG[0, 0] > self.nj_g))
            for (a, b) in _runs(own_x, vx):
                for (c, d) in _runs(own_y, vy):
                    peer = self.rank_of(int(own_x[a]), int(own_y[c]))
                    p_i0g, _, p_j0g, _ = self.extent(peer)
                    si = sorted((int(igs[0, a]) - p_i0g + 1, int(igs[0, b]) - p_i0g + 1))
                    sj = sorted((int(jgs[c, 0]) - p_j0g + 1, int(jgs[d, 0]) - p_j0g + 1))
                    yield (Message(peer=peer, i0=int(il[a]), i1=int(il[b]), j0=int(jl[c]), j1=int(jl[d]), flip=flip),
                           Message(peer=rank, i0=si[0], i1=si[1], j0=sj[0], j1=sj[1], flip=flip))


def _runs(owner: np.ndarray, valid: np.ndarray):
    """Maximal index runs [a, b] with constant owner and valid==True."""
    out, n, a = [], len(owner), None
    for q in range(n):
        if valid[q] and a is None:
            a = q
        if a is not None and (q == n - 1 or not valid[q + 1] or owner[q + 1] != owner[a]):
            if valid[q]:
                out.append((a, q))
            a = None
    return out
