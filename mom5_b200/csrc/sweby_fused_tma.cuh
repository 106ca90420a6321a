// sweby_fused_tma.cuh -- the fused x + y pass of the MDFL Sweby scheme (see sweby_fused.cuh for the algorithm) with every
// operand row staged by the TMA engine.
//
// Reference: advect_tracer_sweby_all OTA:4251-4432 (VAR_ALL), advect_tracer_mdfl_sweby OTA:3916-4056 (VAR_ONE).
//
// Same produce/consume march, same x_face / x_cell / y_level arithmetic and the same register window as k_sweby_xy.  What
// changes is the staging.  A block is FWARPS warps = FWARPS consecutive 31-cell x tiles of one level k, i.e. one CONTIGUOUS
// stretch of 31*FWARPS + 4 columns; per row the block needs one 128-element segment of every operand.  Instead of one
// LDGSTS per thread, operand and row (23 per thread and iteration, each with its 64-bit address arithmetic), ONE thread
// issues one tensor-map copy (cp.async.bulk.tensor, SASS UTMALDG) per operand row for the whole block -- 18 per iteration,
// the two mask-nibble rows (u8, rows padded to a multiple of 16 bytes) included:
//   ring   (4 slots, row r & 3): T[NT], u, rho, (dyte, datr)          row r+1   -> x phase at r+1, y phase at r+1 and r+2
//   x-only (2 slots, row r & 1): tm(z)[NT], dxte, x nibbles           row r+1
//   y-only (2 slots, row jf & 1): th[NT], v, (w(k-1), w(k)), (dxtn, dytn), y nibbles [, rho_dzt(taum1), rho_dztr]   face jf+1
// Pairs in parentheses are two planes / levels of one array and arrive as ONE box.  Everything issued in iteration `it` is
// consumed in iteration it+1 ("fill it+1") and signals the mbarrier full[(it+1) & 1]; a warp reports the end of iteration `it`
// on done[it & 1] and the producer of iteration it+1 waits for that before it overwrites anything.  The producer role rotates over the
// warps (warp it % nact), so no warp carries the issue cost alone.
// (Staging TWO iterations ahead -- 5/3-slot rings, 72 KB per block -- was measured too: the warps then stop spinning on `full`
// (with a distance of one ~17% of the executed instructions are try_wait / branch / yield), but the pass is not a cycle faster,
// profiles/ab_r02_k_*.json: on a power-capped B200 the step time follows the energy per cell-update)
// No thread computes a global address for staging, no register holds one, and no plain load is left in the loop (the
// nibble LDGs of the LDGSTS kernel cost a scoreboard wait per iteration).
// Requirements (checked by the driver, which otherwise launches k_sweby_xy): even ni+2, 16-byte aligned bases.
#pragma once

#include <type_traits>

#include "sweby_fused.cuh"
#include "tma.cuh"

#ifndef FTMINB
#define FTMINB 3
#endif
#define FT_RW 128                    // staged row: 128 doubles (31*FWARPS + 4 = 128 columns for FWARPS = 4)
#define FT_NW 144                    // staged nibble row: 128 + 15 (alignment shift) bytes, rounded up to a multiple of 16
static_assert(31 * FWARPS + 4 <= FT_RW, "block stretch must fit the staged row");

template <int NT, bool UPD>
struct FusedMaps {
    CUtensorMap T[NT], tm[NT], th[NT];     // th unused for UPD
    CUtensorMap u, v, w, rho;              // w: box of two levels (k-1, k)
    CUtensorMap met_ring, dxte, met_y;     // library-owned: (dyte, datr) planes; dxte; (dxtn, dytn) planes
    CUtensorMap rho_m1, rho_r;             // UPD only
    CUtensorMap nibx, niby;                // u8 mask nibbles of the x / y sweeps, rows padded to a multiple of 16 bytes
};

template <int NT, bool UPD>
struct FusedTmaLayout {
    // rows of FT_RW doubles
    static constexpr int NR = NT + 4;                 // ring: T[NT], u, rho, dyte, datr
    static constexpr int R_U = NT, R_RHO = NT + 1, R_DYTE = NT + 2, R_DATR = NT + 3;
    static constexpr int NX = NT + 1;                 // x-only: tm[NT], dxte
    static constexpr int X_DXTE = NT;
    static constexpr int NTH = UPD ? 0 : NT;          // th_tendency rows (the time-update flavour does not read th_tendency)
    static constexpr int NY = NTH + 5 + (UPD ? 2 : 0); // y-only: th[NTH], v, w(k-1), w(k), dxtn, dytn (+ rho_m1, rho_r)
    static constexpr int Y_V = NTH, Y_WM = NTH + 1, Y_WK = NTH + 2, Y_DXTN = NTH + 3, Y_DYTN = NTH + 4, Y_RM1 = NTH + 5, Y_RR = NTH + 6;
    static constexpr int ROWS = 4 * NR + 2 * NX + 2 * NY;
    static constexpr size_t NIB_OFF = (size_t)ROWS * FT_RW * sizeof(double);
    static constexpr size_t BAR_OFF = NIB_OFF + 4 * 256;            // nibx[2], niby[2]: 144-byte rows in 256-byte slots (128-byte aligned boxes)
    static constexpr size_t BYTES = BAR_OFF + 4 * sizeof(uint64_t); // full[2], done[2]
};

template <int NT, int VAR, bool DIAG, bool UPD = false>
__global__ void __launch_bounds__(32 * FWARPS, FTMINB)
k_sweby_xy_tma(const Geom g, const SwebyArgs<NT> a, const __grid_constant__ FusedMaps<NT, UPD> maps, const int nxb, const int nxt)
{
    typedef FusedTmaLayout<NT, UPD> LY;
    extern __shared__ __align__(128) double fsm[];
    double *const ring = fsm;                               // [4][NR][FT_RW]
    double *const xs = ring + 4 * LY::NR * FT_RW;           // [2][NX][FT_RW]
    double *const ys = xs + 2 * LY::NX * FT_RW;             // [2][NY][FT_RW]
    uint8_t *const nbx_s = reinterpret_cast<uint8_t *>(fsm) + LY::NIB_OFF;   // [2][FT_NW]
    uint8_t *const nby_s = nbx_s + 2 * 256;                                  // [2][256]
    uint64_t *const full = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(fsm) + LY::BAR_OFF);
    uint64_t *const done = full + 2;
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    // linear block id, k fastest (2-D metrics and w(k-1) of concurrently resident blocks hit in L2)
    const int lin = blockIdx.x;
    const int k = lin % g.nk + 1;
    const int rest = lin / g.nk;
    const int tile0 = (rest % nxb) * FWARPS;
    const int jc = a.tile_first + (rest / nxb) * a.tile_step;
    const int nact = min(FWARPS, nxt - tile0);            // active warps (x tiles inside the domain), >= 1
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        mbar_init(&done[0], nact); mbar_init(&done[1], nact);
        mbar_fence_init();
    }
    __syncthreads();
    if (wy >= nact) return;                               // whole warps leave; `done` counts the others
    const int iw0 = tile0 * 31;                           // first east-face index of the block = first column of its stretch
    const int col = 31 * wy + lane;                       // this thread's column inside the stretch (face iw0 + col)
    const int i = iw0 + col;                              // east-face index 0..ni; lanes >= 1 also own cell i
    const bool face_ok = (i <= g.ni);
    const bool cell_ok = face_ok && (lane >= 1);
    const int ic = min(i, g.ni);
    const int colw = lane ? col - 1 : col;                // column of the west neighbour's face operands
    const int js = jc * a.kc + 1;
    const int je = min(js + a.kc - 1, g.nj);
    typedef unsigned ofs_t;   // 32-bit element offsets (see k_sweby_z)
    const ofs_t nxd = (ofs_t)g.nxd, tp = (ofs_t)g.tpitch;
    const bool has_km1 = (k > 1);
    const ofs_t q0 = (ofs_t)d3(g, ic, 0, k);              // own element in row 0 (stores, halo-row loads)
    const ofs_t tqc0 = (ofs_t)t3(g, ic, 0, k);
    const int nsh = iw0 & 15;                             // a nibble row is fetched from the 16-byte boundary at or below the stretch
    const bool acc = !UPD && a.accumulate;

    // ---- producer: everything iteration `it` stages (consumed in iteration it + 1) ----
    constexpr unsigned ROWB = FT_RW * sizeof(double);
    auto issue = [&](int jf, uint64_t *bar) {
        const int r1 = jf + 3;                            // row staged for the x phase of the next iteration
        const int f1 = jf + 1;                            // face row staged for the y phase of the next iteration
        const bool do_face = (f1 >= js - 1);
        const int kz = k - 1;                             // level index of the data-domain arrays
        double *R = ring + (r1 & 3) * (LY::NR * FT_RW), *X = xs + (r1 & 1) * (LY::NX * FT_RW), *Y = ys + (f1 & 1) * (LY::NY * FT_RW);
        const unsigned rows_x = 2 * NT + 5;               // T, tm [NT each], u, rho, dyte, datr, dxte
        const unsigned rows_y = (acc ? NT : 0) + 5 + (UPD ? 2 : 0);
        mbar_expect_tx(bar, rows_x * ROWB + FT_NW + (do_face ? rows_y * ROWB + FT_NW : 0u));
#pragma unroll
        for (int n = 0; n < NT; n++) {
            tma_load_3d(X + n * FT_RW, &maps.tm[n], iw0 - 1 + TOFF, r1 + 1, kz, bar);     // h2 row index j+1, column i+TOFF
            tma_load_3d(R + n * FT_RW, &maps.T[n], iw0, r1, kz, bar);
        }
        tma_load_3d(R + LY::R_U * FT_RW, &maps.u, iw0, r1, kz, bar);
        tma_load_3d(R + LY::R_RHO * FT_RW, &maps.rho, iw0, r1, kz, bar);
        tma_load_3d(R + LY::R_DYTE * FT_RW, &maps.met_ring, iw0, r1, 0, bar);              // two planes: dyte, datr
        tma_load_3d(X + LY::X_DXTE * FT_RW, &maps.dxte, iw0, r1, 0, bar);
        tma_load_3d(nbx_s + (r1 & 1) * 256, &maps.nibx, iw0 - nsh, r1, kz, bar);
        if (do_face) {
            if (acc) {
#pragma unroll
                for (int n = 0; n < NT; n++) tma_load_3d(Y + n * FT_RW, &maps.th[n], iw0, f1, kz, bar);
            }
            if (UPD) {
                tma_load_3d(Y + LY::Y_RM1 * FT_RW, &maps.rho_m1, iw0, f1, kz, bar);
                tma_load_3d(Y + LY::Y_RR * FT_RW, &maps.rho_r, iw0, f1, kz, bar);
            }
            tma_load_3d(Y + LY::Y_V * FT_RW, &maps.v, iw0, f1, kz, bar);
            tma_load_3d(Y + LY::Y_WM * FT_RW, &maps.w, iw0, f1, k - 1, bar);               // two levels: w(k-1), w(k) of (..,0:nk)
            tma_load_3d(Y + LY::Y_DXTN * FT_RW, &maps.met_y, iw0, f1, 0, bar);             // two planes: dxtn, dytn
            tma_load_3d(nby_s + (f1 & 1) * 256, &maps.niby, iw0 - nsh, f1, kz, bar);
        }
    };
    auto row_x = [&](int r) { return r >= 1 && r <= g.nj; };       // rows the x sweep is evaluated on

    const int jf0 = js - 4;                                         // first (warm-up) iteration: produces row js-2
    // prologue: what iteration 0 consumes (row jf0+2; the face row jf0 is never used: jf0 < js-1)
    if (wy == 0) {
        if (elect_one()) issue(jf0 - 1, &full[0]);
    }
    const unsigned nby_first = a.nib2[q0 + (ofs_t)(js - 1) * nxd];  // y nibble of the first face (initial differences)

    XFace<NT> F;
    XCell<NT> C;
    YLevel<NT> L;
    F.dtime = a.dtime; F.sl = a.sl; C.dtime = a.dtime;
    L.dtime = a.dtime; L.sl = a.sl;
    L.rho0 = 0.0;
#pragma unroll
    for (int n = 0; n < NT; n++) { L.t0[n] = 0.0; L.t1[n] = 0.0; L.Rm1[n] = 0.0; L.R0[n] = 0.0; L.fprev[n] = 0.0; }

    int it = 0;                                                     // iteration counter; jf = jf0 + it
    int prod = 0;                                                   // producer warp of this iteration = it % nact
    // One iteration of the march.  GEN = true is the general form; GEN = false is the steady state of a chunk that touches no
    // halo row of the x-updated tracer (rows js-2 .. je+2 inside 1..nj) once its warm-up is over: the x phase always runs, the y
    // phase always runs on a live cell -- the loop body is then free of data-independent branches (ncu: branch_resolving was
    // 0.45 stall cycles per issued instruction with the general body).
    auto body = [&](const int jf, auto gen_tag) {
        constexpr bool GEN = decltype(gen_tag)::value;
        const int r = jf + 2;                                       // row produced by this iteration
        // ---- producer of this iteration: stage row r+1 and face jf+1 once every warp has finished iteration it-1 ----
        if (jf < je && wy == prod) {                                // warp-uniform
            if (elect_one()) {
                if (it >= 1) mbar_wait(&done[(it - 1) & 1], (unsigned)((it - 1) >> 1) & 1u);
                issue(jf, &full[(it + 1) & 1]);
            }
            __syncwarp();
        }
        if (++prod == nact) prod = 0;
        // ---- operands of this iteration: fill number `it` (issued during iteration it-1; the prologue is fill 0) ----
        mbar_wait(&full[it & 1], (unsigned)(it >> 1) & 1u);

        // ---------------- x: produce tm(i, r, k) -> L.t2 ----------------
        if (!GEN || row_x(r)) {
            const double *R = ring + (r & 3) * (LY::NR * FT_RW), *X = xs + (r & 1) * (LY::NX * FT_RW);
            const unsigned nbx = nbx_s[(r & 1) * 256 + col + nsh];
            F.nb = nbx;
            F.dyte = R[LY::R_DYTE * FT_RW + col];
            F.dxte = X[LY::X_DXTE * FT_RW + col];
            F.uu = R[LY::R_U * FT_RW + col];
            F.rho_i = R[LY::R_RHO * FT_RW + col];
            F.rho_e = R[LY::R_RHO * FT_RW + col + 1];
#pragma unroll
            for (int n = 0; n < NT; n++) {
                F.tm1[n] = X[n * FT_RW + col]; F.t0[n] = X[n * FT_RW + col + 1];
                F.t1[n] = X[n * FT_RW + col + 2]; F.t2[n] = X[n * FT_RW + col + 3];
            }
            if (x_face<NT, VAR, false>(F) && face_ok) {
                XFace<NT> Z = F;
                x_face_exact<NT, VAR>(&Z);
                F.mf = Z.mf;
#pragma unroll
                for (int n = 0; n < NT; n++) F.f[n] = Z.f[n];
            }
            C.m_i = nib_and(nbx, 2u); C.rho_i = F.rho_i; C.mf = F.mf;
            C.datr = R[LY::R_DATR * FT_RW + col];
            C.mfw = __shfl_up_sync(0xffffffffu, F.mf, 1);
            const bool own_row = DIAG && (r >= js) && (r <= je);    // diagnostics are written by the chunk that owns the row
            const ofs_t qd = q0 + (ofs_t)r * nxd;
#pragma unroll
            for (int n = 0; n < NT; n++) {
                C.f[n] = F.f[n];
                C.fw[n] = __shfl_up_sync(0xffffffffu, F.f[n], 1);
                C.Tc[n] = R[n * FT_RW + col];
                C.t0[n] = F.t0[n];
                if (DIAG && own_row && face_ok && a.flux[n]) a.flux[n][qd] = F.f[n];
            }
            if (x_cell<NT, VAR, false>(C) && cell_ok) {
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

        if (!GEN || jf >= js - 1) {
            // ---------------- y: north face jf and (when live) cell (i, jf, k) ----------------
            const double *R = ring + (jf & 3) * (LY::NR * FT_RW), *R1 = ring + ((jf + 1) & 3) * (LY::NR * FT_RW);
            const double *Y = ys + (jf & 1) * (LY::NY * FT_RW);
            if (GEN && jf == js - 1) L.rho0 = R[LY::R_RHO * FT_RW + col];
            L.nb = nby_s[(jf & 1) * 256 + col + nsh];
            L.live = GEN ? (jf >= js) : 1;
            L.vv = Y[LY::Y_V * FT_RW + col];
            L.rho1 = R1[LY::R_RHO * FT_RW + col];
            L.dxtn = Y[LY::Y_DXTN * FT_RW + col];
            L.dytn = Y[LY::Y_DYTN * FT_RW + col];
            L.datr = R[LY::R_DATR * FT_RW + col];
            L.wk = Y[LY::Y_WK * FT_RW + col];
            L.wkm1 = has_km1 ? Y[LY::Y_WM * FT_RW + col] : 0.0;
            L.dyte_w = R[LY::R_DYTE * FT_RW + colw]; L.u_w = R[LY::R_U * FT_RW + colw];
            L.dyte_c = R[LY::R_DYTE * FT_RW + col]; L.u_c = R[LY::R_U * FT_RW + col];
#pragma unroll
            for (int n = 0; n < NT; n++) L.Tc[n] = R[n * FT_RW + col];
            if (y_level<NT, VAR, false>(L) && cell_ok) {
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
                        a.Tnew[n][q] = ((Y[LY::Y_RM1 * FT_RW + col] * L.Tc[n]) + (a.dtime * thv)) * Y[LY::Y_RR * FT_RW + col];
                    } else {
                        a.adv[n][q] = L.adv[n];
                        if (a.accumulate) a.th[n][q] = Y[n * FT_RW + col] + L.adv[n];
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
        // ---- end of iteration: every value read from the staging slots has been consumed ----
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[it & 1]);
        it++;
    };
    int jf = jf0;
#ifdef FT_PEEL   // measured on B200 (profiles/ab_r02_j_*.json): the specialised steady-state body costs 10 registers and is 2% SLOWER
    const bool steady = (js - 2 >= 1) && (je + 2 <= g.nj);          // uniform over the block
#else
    const bool steady = false;                                      // general body everywhere
#endif
    const int j_gen_end = steady ? js - 1 : je;                     // warm-up rows and the first (face-only) y iteration, or everything
    for (; jf <= j_gen_end; jf++) body(jf, std::true_type{});
#pragma unroll kFusedUnroll
    for (; jf <= je; jf++) body(jf, std::false_type{});
}
