// sweby_z_tma.cuh -- the z sweep of the MDFL Sweby scheme with its k-columns staged by the TMA engine.
//
// Reference: advect_tracer_sweby_all z sweep OTA:4150-4211 (VAR_ALL), advect_tracer_mdfl_sweby OTA:3843-3911 (VAR_ONE).
// Same arithmetic, thread mapping (one thread per (i,j) column marching down k) and register window as k_sweby_z
// (sweby_kernels.cuh); what changes is how the operands reach the SM:
//   - a block owns ZBX consecutive columns of one row j.  Its operands -- T(k+2) per tracer, w(k), rho_dzt(k) -- are fetched
//     as BOXES of ZBX columns x ZT_KC levels, one cp.async.bulk.tensor (SASS UTMALDG) per operand and box, issued by one
//     elected thread into a ZT_SLOTS-deep ring of shared-memory slots: NT+2 instructions per ZT_KC levels and block instead
//     of NT+2 LDGSTS (+ their 64-bit address arithmetic) per level and THREAD;
//   - completion is tracked per slot by an mbarrier (`full`, transaction bytes); warps hand a slot back through a second
//     mbarrier (`empty`, one arrival per warp): no block-wide barrier in the loop, warps may drift a slot apart;
//   - the land/sea mask of a column arrives as a BIT STRING (mom5adv_init packs it: bit b = tmask(i,j,clamp(b,1,nk)),
//     b = 0..nk+2), three 32-bit words for 75 levels, so the loop carries no per-level mask load at all.  In the LDGSTS kernel
//     of round 1 the per-level nibble byte was a plain LDG whose scoreboard wait exposed one DRAM latency per level (ncu:
//     85% of the kernel's long-scoreboard stall samples sat on the first instruction after that wait).
// T(min(k+2,nk)): the box holds T(k+2); below the last level the tensor map returns zeros and the clamped value is the
// register copy of T(k+1) = T(nk) instead.
#pragma once

#include "sweby_kernels.cuh"
#include "tma.cuh"

#ifndef ZT_KC
#define ZT_KC 3        // levels per box; a multiple of 3 makes the rotation of the 3-deep register windows free when unrolled
#endif
#ifndef ZT_SLOTS
#define ZT_SLOTS 2     // boxes in flight / being computed on
#endif
#ifndef ZTMINB
#define ZTMINB 4
#endif

template <int NT>
struct ZMaps {
    CUtensorMap T[NT], w, rho;   // 3-D maps over (isd:ied, jsd:jed, levels), box ZBX x 1 x ZT_KC
};

template <int NT>
struct ZTLayout {
    static constexpr int NF = NT + 2;
    static constexpr size_t SLOT = (size_t)NF * ZT_KC * ZBX * sizeof(double);
    static constexpr size_t BYTES = ZT_SLOTS * SLOT + 2 * ZT_SLOTS * sizeof(uint64_t);
};

template <int NT, int VAR, bool DIAG>
__global__ void __launch_bounds__(ZBX, ZTMINB)
k_sweby_z_tma(const Geom g, const SwebyArgs<NT> a, const __grid_constant__ ZMaps<NT> maps, const unsigned *__restrict__ zbits, const int nzw)
{
    constexpr int NF = NT + 2, KC = ZT_KC, D = ZT_SLOTS;
    constexpr unsigned BOX_BYTES = KC * ZBX * sizeof(double);
    // slots: [D][NF][KC][ZBX] doubles (tensor-map destinations: 128-byte aligned); then the barriers
    extern __shared__ __align__(128) double sm[];
    uint64_t *const full = reinterpret_cast<uint64_t *>(sm + (size_t)D * NF * KC * ZBX);
    uint64_t *const empty = full + D;
    const int tx = threadIdx.x, lane = tx & 31, warp = tx >> 5;
    // tile t covers the data-domain columns t*ZBX .. t*ZBX + ZBX-1: the box origin is EVEN, i.e. on a 16-byte boundary, as the
    // tensor-map copy of FP64 data requires (an odd first coordinate is an illegal instruction on B200, tests/cuda/tma_probe.cu);
    // column 0 -- the west halo -- is an idle lane of tile 0
    const int i0 = (a.tile_first + (int)blockIdx.x * a.tile_step) * ZBX;
    const int j = a.row_first + (int)blockIdx.y;
    const int ncol = min(ZBX, g.ni - i0 + 1);            // columns i0 .. min(i0 + ZBX-1, ni): > 0 by construction of the grid
    const int nwarps = (ncol + 31) >> 5;
    if (tx == 0) {
#pragma unroll
        for (int s = 0; s < D; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], nwarps); }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp >= nwarps) return;                          // whole warps outside the domain leave; `empty` counts the others
    const bool col_ok = (tx < ncol) && (i0 + tx >= 1);
    const int i = max(i0 + min(tx, ncol - 1), 1);        // clamped column for the direct loads of idle lanes
    const int ks = blockIdx.z * a.kc + 1;
    const int ke = min(ks + a.kc - 1, g.nk);
    const int k0 = ks > 1 ? ks - 1 : 1;                  // first face evaluated (warm-up face below the surface chunk)
    const int nbox = (ke - k0 + KC) / KC;                // boxes of KC levels starting at k0

    // ---- producer (thread 0): box b (levels k0 + b*KC ...) into slot b % D ----
    auto issue = [&](int b) {
        const int s = b % D, kb = k0 + b * KC;           // first level of the box (1-based)
        double *dst = sm + (size_t)s * NF * KC * ZBX;
        mbar_expect_tx(&full[s], NF * BOX_BYTES);
#pragma unroll
        for (int n = 0; n < NT; n++) tma_load_3d(dst + (size_t)n * KC * ZBX, &maps.T[n], i0, j, kb + 1, &full[s]);   // T(kb+2): index kb+1
        tma_load_3d(dst + (size_t)NT * KC * ZBX, &maps.w, i0, j, kb, &full[s]);                                     // w(kb): (..,0:nk)
        tma_load_3d(dst + (size_t)(NT + 1) * KC * ZBX, &maps.rho, i0, j, kb - 1, &full[s]);                         // rho(kb)
    };
    if (warp == 0) {
        if (elect_one()) {
#pragma unroll
            for (int b = 0; b < D; b++)
                if (b < nbox) issue(b);
        }
        __syncwarp();
    }

    typedef unsigned ofs_t;                              // 32-bit element offsets (every array < 2^32 elements, mom5adv_init)
    const ofs_t c2 = (ofs_t)d2(g, i, j);
    const ofs_t slab = (ofs_t)g.slab;
    ofs_t qd = (ofs_t)d3(g, i, j, k0);                   // data-domain offset of level k
    ofs_t qt = (ofs_t)t3(g, i, j, k0);                   // h2 offset of level k
    const ofs_t qkm = (k0 > 1) ? qd - slab : qd, qkp = (k0 < g.nk) ? qd + slab : qd;

    ZBits zb;
    zb.init(zbits + c2, slab, nzw, k0);
    ZLevel<NT> L;
    L.dat = a.dat[c2]; L.datr = a.datr[c2]; L.dtime = a.dtime; L.sl = a.sl;
    L.nb = zb.nib();
#pragma unroll
    for (int n = 0; n < NT; n++) {
        const double Tkm = a.T[n][qkm];
        L.Tk[n] = a.T[n][qd];
        L.Tp1[n] = a.T[n][qkp];
        L.Rm1[n] = (Tkm - L.Tk[n]) * nib_and(L.nb, 3u);      // ((T(km1)-T(k))*m(km1))*m(k); +0 at k = 1 (km1 clamps)
        L.R0[n] = (L.Tk[n] - L.Tp1[n]) * nib_and(L.nb, 6u);  // ((T(k)-T(kp1))*m(k))*m(kp1)
        L.ftp[n] = 0.0;
    }
    L.wkm1 = 0.0;

    int k = k0;
#pragma unroll 1
    for (int b = 0; b < nbox; b++) {
        const int s = b % D;
        const unsigned ph = (unsigned)(b / D) & 1u;
        mbar_wait(&full[s], ph);
        const double *S = sm + (size_t)s * NF * KC * ZBX + tx;
#pragma unroll
        for (int q = 0; q < KC; q++) {
            // Levels beyond ke (the tail of the last box) run the arithmetic on whatever the box holds -- zeros below the
            // bottom -- and store nothing: an unconditional body lets the compiler rotate the register windows for free.
            const bool in_chunk = (k <= ke);              // uniform over the block
            const bool has_p2 = (k + 2 <= g.nk);
#pragma unroll
            for (int n = 0; n < NT; n++) {
                const double v = S[(n * KC + q) * ZBX];
                L.Tp2[n] = has_p2 ? v : L.Tp1[n];         // T(min(k+2,nk)) = T(nk) = T(min(k+1,nk)) once k+2 > nk
            }
            L.wk = S[(NT * KC + q) * ZBX];
            L.r = S[((NT + 1) * KC + q) * ZBX];
            L.nb = zb.nib();
            if (z_level<NT, VAR, false>(L) && col_ok && in_chunk) {
                ZLevel<NT> X = L;
                z_level_exact<NT, VAR>(&X);
#pragma unroll
                for (int n = 0; n < NT; n++) { L.fbt[n] = X.fbt[n]; L.Rp1[n] = X.Rp1[n]; L.t[n] = X.t[n]; L.wz[n] = X.wz[n]; }
            }
            const bool live = (k >= ks) && in_chunk && col_ok;
#pragma unroll
            for (int n = 0; n < NT; n++) {
                if (live) {
                    a.tm_in[n][qt] = L.t[n];
                    if (DIAG && VAR == VAR_ALL && a.dadv[n]) a.dadv[n][qd] = L.wz[n];
                    if (DIAG && a.flux[n]) a.flux[n][qd] = L.fbt[n];
                }
                L.ftp[n] = L.fbt[n];
                L.Rm1[n] = L.R0[n];
                L.R0[n] = L.Rp1[n];
                L.Tk[n] = L.Tp1[n];
                L.Tp1[n] = L.Tp2[n];
            }
            L.wkm1 = L.wk;
            zb.next(zbits + c2, slab, nzw);
            qd += slab;
            qt += (ofs_t)g.tslab;
            k++;
        }
        // ---- hand the slot back (every lane's reads of it have been consumed by the arithmetic above) and refill it ----
        if (b + D < nbox) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (warp == 0) {                              // warp-uniform; one elected lane refills the slot
                if (elect_one()) {
                    mbar_wait(&empty[s], ph);
                    issue(b + D);
                }
                __syncwarp();
            }
        }
    }
}
