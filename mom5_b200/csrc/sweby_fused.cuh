// sweby_fused.cuh -- the x and y sweeps of the MDFL Sweby scheme in ONE pass over HBM.
//
// Reference: advect_tracer_sweby_all   OTA:4251-4432  (VAR_ALL)
//            advect_tracer_mdfl_sweby  OTA:3916-4056  (VAR_ONE)
//
// The reference materialises the running tracer tm between the x and the y sweep (it has to: an mpp_update_domains sits
// between them).  Here a warp marches north over a chunk of j for 32 consecutive east faces (= 31 cells, as in
// k_sweby_x) at one level k and, every iteration,
//   (1) PRODUCES one row of the x-updated tm (row jf+2) from the z-updated tm, with the very same x_face / x_cell
//       arithmetic as the stand-alone x sweep, and
//   (2) CONSUMES it in the y sweep's rolling window (tm(jf-1..jf+2)) for north face jf and cell (i,jf) with the same
//       y_level arithmetic as the stand-alone y sweep.
// The x-updated tm never goes to HBM: per cell the pass reads tm(z), T, th, u, v, w, rho and writes th, adv -- the
// traffic of the stand-alone y sweep -- and the whole stand-alone x sweep (24*ntr + 16 B per cell) disappears.
//
// Rows outside 1..nj (the width-2 north/south halo of the x-updated tm: neighbour ranks, cyclic wrap, tripolar fold)
// cannot be produced locally -- the caller's arrays have halo 1 only.  They are read from the h2 scratch `tm_out`,
// which the driver fills beforehand: stand-alone x sweep over the four edge rows 1, 2, nj-1, nj, then the usual
// north/south strip exchange.  A chunk starts with three warm-up rows (x only), so the x arithmetic is redone for
// 4 rows per chunk.
//
// Staging: per-thread cp.async one iteration ahead, as in the stand-alone sweeps.  Operands both phases need
// (T, u, rho, dyte, datr of a row: x at row jf+2, y at row jf) are staged ONCE into a 4-slot ring indexed by row & 3.
#pragma once

#include "sweby_kernels.cuh"

#ifndef FWARPS
#define FWARPS 4     // warps per block = consecutive 31-cell x tiles at the same k
#endif
#ifndef FMINB
#define FMINB 3     // 168 registers: no spills; 4 blocks (128 registers, 224 KB of shared memory, no L1 left) measured 35% slower
#endif
#ifndef FUNROLL
#define FUNROLL 1
#endif
constexpr int kFusedUnroll = FUNROLL;
#define RROW 34      // ring row: 32 lanes (+1 for rho's east neighbour, +1 pad)

// UPD: the epilogue also performs update_advection_only's time update (ocean_tracer.F90:2618-2649):
//   th_tendency = 0 + T_prog%wrk1;  field(taup1) = (rho_dzt(taum1)*field(taum1) + dtime*th_tendency)*rho_dztr
// th_tendency is then not read, and neither it nor wrk1 has to be written (null pointers skip the stores).
template <int NT, bool UPD = false>
struct FusedLayout {
    // per-warp shared memory, in doubles
    static constexpr int NR = NT + 4;                 // ring fields: T[NT], u, rho, dyte, datr
    static constexpr int R_U = NT, R_RHO = NT + 1, R_DYTE = NT + 2, R_DATR = NT + 3;
    static constexpr int NX = NT + 1;                 // x-only fields: tm(z)[NT], dxte
    static constexpr int X_DXTE = NT;
    static constexpr int NY = NT + 5 + (UPD ? 2 : 0); // y-only fields: th[NT], v, w(k), w(k-1), dxtn, dytn (+ rho_dzt(taum1), rho_dztr)
    static constexpr int Y_V = NT, Y_WK = NT + 1, Y_WM = NT + 2, Y_DXTN = NT + 3, Y_DYTN = NT + 4, Y_RM1 = NT + 5, Y_RR = NT + 6;
    static constexpr int RING = 4 * NR * RROW, XS = 2 * NX * XROW, YS = 2 * NY * 32;
    static constexpr int PER_WARP = RING + XS + YS;
    static constexpr size_t BYTES = (size_t)PER_WARP * FWARPS * sizeof(double);
};

template <int NT, int VAR, bool DIAG, bool UPD = false>
__global__ void __launch_bounds__(32 * FWARPS, FMINB) k_sweby_xy(const Geom g, const SwebyArgs<NT> a, const int nxb, const int nxt)
{
    typedef FusedLayout<NT, UPD> LY;
    extern __shared__ double fsm[];
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    double *const ring = fsm + (size_t)wy * LY::PER_WARP;   // [4][NR][RROW]
    double *const xs = ring + LY::RING;                     // [2][NX][XROW]
    double *const ys = xs + LY::XS;                         // [2][NY][32]
    // linear block id, k fastest (2-D metrics and w(k-1) of concurrently resident blocks hit in L2)
    const int lin = blockIdx.x;
    const int k = lin % g.nk + 1;
    const int rest = lin / g.nk;
    const int tile = (rest % nxb) * FWARPS + wy;
    const int jc = a.tile_first + (rest / nxb) * a.tile_step;
    if (tile >= nxt) return;                             // whole warp leaves together; warps are independent
    const int iw = tile * 31;                            // first east-face index of this warp
    const int i = iw + lane;                             // east-face index 0..ni; lanes >= 1 also own cell i
    const bool face_ok = (i <= g.ni);
    const bool cell_ok = face_ok && (lane >= 1);
    const int ic = min(i, g.ni);
    const int lw = lane ? lane - 1 : 0;                  // slot of the west neighbour's face operands
    const int js = jc * a.kc + 1;
    const int je = min(js + a.kc - 1, g.nj);
    typedef unsigned ofs_t;   // 32-bit element offsets (see k_sweby_z)
    const ofs_t nxd = (ofs_t)g.nxd, tp = (ofs_t)g.tpitch, wofs = (ofs_t)g.slab;
    const bool has_km1 = (k > 1);
    // offsets of row 0 (row r adds r*nxd / r*tp; rows -1 and nj+2 are never dereferenced through the d-offsets)
    const ofs_t q0 = (ofs_t)d3(g, ic, 0, k), c0 = (ofs_t)d2(g, ic, 0);
    const ofs_t qr0 = (ofs_t)d3(g, min(iw + lane, g.ni + 1), 0, k), qrb0 = (ofs_t)d3(g, min(iw + 32, g.ni + 1), 0, k);
    const ofs_t tqa0 = (ofs_t)t3(g, min(iw - 1 + lane, g.ni + 2), 0, k), tqb0 = (ofs_t)t3(g, min(iw + 31 + lane, g.ni + 2), 0, k);
    const ofs_t tqc0 = (ofs_t)t3(g, ic, 0, k);

    // ---- staging ----
    auto stage_row = [&](int r) {     // operands of row r (0 <= r <= nj+1): ring slot r & 3, x-only slot r & 1
        double *R = ring + (r & 3) * (LY::NR * RROW), *X = xs + (r & 1) * (LY::NX * XROW);
        const ofs_t ro = (ofs_t)r * nxd, q = q0 + ro, c2 = c0 + ro, rt = (ofs_t)r * tp;
#pragma unroll
        for (int n = 0; n < NT; n++) {
            cp_async8(&X[n * XROW + lane], a.tm_in[n] + tqa0 + rt);
            if (lane < 3) cp_async8(&X[n * XROW + 32 + lane], a.tm_in[n] + tqb0 + rt);
            cp_async8(&R[n * RROW + lane], a.T[n] + q);
        }
        cp_async8(&R[LY::R_U * RROW + lane], a.u + q);
        cp_async8(&R[LY::R_RHO * RROW + lane], a.rho + qr0 + ro);
        if (lane == 0) cp_async8(&R[LY::R_RHO * RROW + 32], a.rho + qrb0 + ro);
        cp_async8(&R[LY::R_DYTE * RROW + lane], a.dyte + c2);
        cp_async8(&R[LY::R_DATR * RROW + lane], a.datr + c2);
        cp_async8(&X[LY::X_DXTE * XROW + lane], a.dxte + c2);
    };
    auto stage_face = [&](int jf) {   // y-only operands of north face / cell row jf (0 <= jf <= nj): slot jf & 1
        double *Y = ys + (jf & 1) * (LY::NY * 32);
        const ofs_t ro = (ofs_t)jf * nxd, q = q0 + ro, c2 = c0 + ro;
#pragma unroll
        for (int n = 0; n < NT; n++)
            if (!UPD && a.accumulate) cp_async8(&Y[n * 32 + lane], a.th[n] + q);
        if (UPD) {
            cp_async8(&Y[LY::Y_RM1 * 32 + lane], a.rho_m1 + q);
            cp_async8(&Y[LY::Y_RR * 32 + lane], a.rho_r + q);
        }
        cp_async8(&Y[LY::Y_V * 32 + lane], a.v + q);
        cp_async8(&Y[LY::Y_WK * 32 + lane], a.w + q + wofs);          // w3(k) = d3(k) + slab ; w3(k-1) = d3(k)
        if (has_km1) cp_async8(&Y[LY::Y_WM * 32 + lane], a.w + q);
        cp_async8(&Y[LY::Y_DXTN * 32 + lane], a.dxtn + c2);
        cp_async8(&Y[LY::Y_DYTN * 32 + lane], a.dytn + c2);
    };
    auto row_ok = [&](int r) { return r >= 0 && r <= g.nj + 1; };
    auto row_x = [&](int r) { return r >= 1 && r <= g.nj; };       // rows the x sweep is evaluated on

    const int jf0 = js - 4;                                         // first (warm-up) iteration: produces row js-2
    unsigned nbx = 0, nby = 0;
    if (row_ok(jf0 + 2)) stage_row(jf0 + 2);
    if (row_x(jf0 + 2)) nbx = a.nib[q0 + (ofs_t)(jf0 + 2) * nxd];
    cp_async_commit();
    const unsigned nby_first = a.nib2[q0 + (ofs_t)(js - 1) * nxd];  // y nibble of the first face (initial differences)

    XFace<NT> F;
    XCell<NT> C;
    YLevel<NT> L;
    F.dtime = a.dtime; F.sl = a.sl; C.dtime = a.dtime;
    L.dtime = a.dtime; L.sl = a.sl;
    L.rho0 = 0.0;
#pragma unroll
    for (int n = 0; n < NT; n++) { L.t0[n] = 0.0; L.t1[n] = 0.0; L.Rm1[n] = 0.0; L.R0[n] = 0.0; L.fprev[n] = 0.0; }

#pragma unroll kFusedUnroll
    for (int jf = jf0; jf <= je; jf++) {
        const int r = jf + 2;                                       // row produced by this iteration
        __syncwarp();                                               // everyone is done with the slots about to be refilled
        unsigned nbx_n = 0, nby_n = 0;
        if (jf < je) {
            if (row_ok(r + 1)) stage_row(r + 1);
            if (row_x(r + 1)) nbx_n = a.nib[q0 + (ofs_t)(r + 1) * nxd];
            if (jf + 1 >= js - 1) {
                stage_face(jf + 1);
                nby_n = a.nib2[q0 + (ofs_t)(jf + 1) * nxd];
            }
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();

        // ---------------- x: produce tm(i, r, k) -> L.t2 ----------------
        if (row_x(r)) {
            const double *R = ring + (r & 3) * (LY::NR * RROW), *X = xs + (r & 1) * (LY::NX * XROW);
            F.nb = nbx;
            F.dyte = R[LY::R_DYTE * RROW + lane];
            F.dxte = X[LY::X_DXTE * XROW + lane];
            F.uu = R[LY::R_U * RROW + lane];
            F.rho_i = R[LY::R_RHO * RROW + lane];
            F.rho_e = R[LY::R_RHO * RROW + lane + 1];
#pragma unroll
            for (int n = 0; n < NT; n++) {
                F.tm1[n] = X[n * XROW + lane]; F.t0[n] = X[n * XROW + lane + 1];
                F.t1[n] = X[n * XROW + lane + 2]; F.t2[n] = X[n * XROW + lane + 3];
            }
            if (x_face<NT, VAR, false>(F)) {
                XFace<NT> Z = F;
                x_face_exact<NT, VAR>(&Z);
                F.mf = Z.mf;
#pragma unroll
                for (int n = 0; n < NT; n++) F.f[n] = Z.f[n];
            }
            C.m_i = nib_and(nbx, 2u); C.rho_i = F.rho_i; C.mf = F.mf;
            C.datr = R[LY::R_DATR * RROW + lane];
            C.mfw = __shfl_up_sync(0xffffffffu, F.mf, 1);
            const bool own_row = DIAG && (r >= js) && (r <= je);    // diagnostics are written by the chunk that owns the row
            const ofs_t qd = q0 + (ofs_t)r * nxd;
#pragma unroll
            for (int n = 0; n < NT; n++) {
                C.f[n] = F.f[n];
                C.fw[n] = __shfl_up_sync(0xffffffffu, F.f[n], 1);
                C.Tc[n] = R[n * RROW + lane];
                C.t0[n] = F.t0[n];
                if (DIAG && own_row && face_ok && a.flux[n]) a.flux[n][qd] = F.f[n];
            }
            if (x_cell<NT, VAR, false>(C)) {
                XCell<NT> Z = C;
                x_cell_exact<NT, VAR>(&Z);
#pragma unroll
                for (int n = 0; n < NT; n++) { C.t[n] = Z.t[n]; C.wx[n] = Z.wx[n]; }
            }
#pragma unroll
            for (int n = 0; n < NT; n++) {
                L.t2[n] = C.t[n];
                if (DIAG && VAR == VAR_ALL && own_row && cell_ok && a.dadv[n]) a.dadv[n][qd] = C.wx[n];
            }
        } else {
            // halo row of the x-updated tm: filled by the driver (edge-row x sweep + north/south strip update)
#pragma unroll
            for (int n = 0; n < NT; n++) L.t2[n] = a.tm_out[n][tqc0 + (ofs_t)r * tp];
        }

        if (jf >= js - 1) {
            // ---------------- y: north face jf and (when live) cell (i, jf, k) ----------------
            const double *R = ring + (jf & 3) * (LY::NR * RROW), *R1 = ring + ((jf + 1) & 3) * (LY::NR * RROW);
            const double *Y = ys + (jf & 1) * (LY::NY * 32);
            if (jf == js - 1) L.rho0 = R[LY::R_RHO * RROW + lane];
            L.nb = nby;
            L.live = (jf >= js);
            L.vv = Y[LY::Y_V * 32 + lane];
            L.rho1 = R1[LY::R_RHO * RROW + lane];
            L.dxtn = Y[LY::Y_DXTN * 32 + lane];
            L.dytn = Y[LY::Y_DYTN * 32 + lane];
            L.datr = R[LY::R_DATR * RROW + lane];
            L.wk = Y[LY::Y_WK * 32 + lane];
            L.wkm1 = has_km1 ? Y[LY::Y_WM * 32 + lane] : 0.0;
            L.dyte_w = R[LY::R_DYTE * RROW + lw]; L.u_w = R[LY::R_U * RROW + lw];
            L.dyte_c = R[LY::R_DYTE * RROW + lane]; L.u_c = R[LY::R_U * RROW + lane];
#pragma unroll
            for (int n = 0; n < NT; n++) L.Tc[n] = R[n * RROW + lane];
            if (y_level<NT, VAR, false>(L)) {
                YLevel<NT> Z = L;
                y_level_exact<NT, VAR>(&Z);
#pragma unroll
                for (int n = 0; n < NT; n++) { L.f[n] = Z.f[n]; L.Rp1[n] = Z.Rp1[n]; L.adv[n] = Z.adv[n]; L.wy[n] = Z.wy[n]; }
            }
            const ofs_t q = q0 + (ofs_t)jf * nxd;
#pragma unroll
            for (int n = 0; n < NT; n++) {
                if (DIAG && cell_ok && a.flux2[n]) a.flux2[n][q] = L.f[n];
                if (L.live && cell_ok) {
                    if (UPD) {
                        const double thv = 0.0 + L.adv[n];             // th_tendency = 0.0, then += wrk1
                        if (a.adv[n]) a.adv[n][q] = L.adv[n];
                        if (a.th[n]) a.th[n][q] = thv;
                        a.Tnew[n][q] = ((Y[LY::Y_RM1 * 32 + lane] * L.Tc[n]) + (a.dtime * thv)) * Y[LY::Y_RR * 32 + lane];
                    } else {
                        a.adv[n][q] = L.adv[n];
                        if (a.accumulate) a.th[n][q] = Y[n * 32 + lane] + L.adv[n];
                    }
                    if (DIAG && VAR == VAR_ALL && a.dadv2[n]) a.dadv2[n][q] = L.wy[n];
                }
                L.fprev[n] = L.f[n];
                L.Rm1[n] = L.R0[n];
                L.R0[n] = L.Rp1[n];
            }
            L.rho0 = L.rho1;
        } else {
            // warm-up: build the differences of the first face (js-1) as the stand-alone y sweep does, from ITS nibble:
            // row js-1 arrives -> ((tm(j)-tm(j-1))*m(j))*m(j-1);  row js arrives -> ((tm(j+1)-tm(j))*m(j+1))*m(j)
            const double mm = nib_and(nby_first, (jf == js - 3) ? 3u : 6u);
#pragma unroll
            for (int n = 0; n < NT; n++) {
                L.Rm1[n] = L.R0[n];
                L.R0[n] = (L.t2[n] - L.t1[n]) * mm;
            }
        }
#pragma unroll
        for (int n = 0; n < NT; n++) { L.t0[n] = L.t1[n]; L.t1[n] = L.t2[n]; }
        nbx = nbx_n;
        nby = nby_n;
    }
}
