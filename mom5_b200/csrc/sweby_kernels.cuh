// sweby_kernels.cuh -- the three directional sweeps of the MDFL Sweby scheme as FP64 stencil kernels.
//
// Reference: advect_tracer_sweby_all   OTA:4104-4511  (VAR_ALL)
//            advect_tracer_mdfl_sweby  OTA:3806-4066  (VAR_ONE; different association order in the updates)
// OTA = src/mom5/ocean_tracers/ocean_tracer_advect.F90 of the reference tree.
//
// Thread mapping (lanes always run along i, the contiguous index, so every global access is coalesced):
//   z sweep: one thread per (i,j) column marching down k; the k-carried quantities of the reference
//            (ftp, wkm1) and the three limiter differences R(k-1), R(k), R(k+1) live in registers, so every
//            z-face flux and every difference is computed exactly once.  The next level's operands are loaded
//            into registers one iteration ahead (software pipelining).
//   x sweep: one thread per (i,j) east face, looping over k with the 2-D metrics in registers.  A warp owns
//            32 consecutive faces = 31 cells; the west-face flux and mass flux come from the neighbouring lane by
//            warp shuffle, so warps are independent (no shared memory, no block barrier).
//   y sweep: one thread per (i,k) marching north over a chunk of j with a rolling register window
//            (tm(j-1..j+2), R(j-1..j+1), flux(j-1)); blocks are ordered k-fastest so that the 2-D metrics and
//            w(k-1) of concurrently resident blocks hit in L2.
// x and y issue L2 prefetches for the next iteration's lines (no register cost).
// All NT tracers of a group are advanced by the same thread so the tracer-independent face coefficients
// (cfl, d0, d1, (1-cfl)/(1e-30+cfl), mf+-|mf|, mask products, reciprocals of rho_dzt and dtime) are computed once.
// Masks come as one byte per cell and sweep direction holding m(-1), m(0), m(+1), m(+2) along that direction
// (built once in mom5adv_init from tmask_mdfl with its halo-2 update, OTA:1668-1675).
#pragma once

#include "mom5adv_internal.cuh"

#define MAXNT 4

template <int NT>
struct SwebyArgs {
    const double *T[NT];        // T(taum1), data-domain layout
    double *tm_in[NT];          // h2 scratch read by the sweep (x, y) / written (z)
    double *tm_out[NT];         // h2 scratch written by the x sweep
    double *th[NT];             // y: th_tendency (+=)
    double *adv[NT];            // y: T_prog(n)%wrk1 / Tracer%wrk1
    double *flux[NT];           // optional diagnostics (data-domain layout) or nullptr
    double *dadv[NT];           // optional per-direction tendency diagnostics or nullptr
    const double *u, *v, *w, *rho;
    const uint8_t *nib;         // mask nibbles of this sweep's direction, data-domain layout
    const double *dat, *datr, *dxte, *dyte, *dxtn, *dytn;
    double dtime, sl;
    int kc;                     // z/x: levels per k-chunk;  y: rows per j-chunk
    int accumulate;             // y: th += adv
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// =================================================================================================
// z sweep  (OTA:4150-4211 / 3843-3911)
// =================================================================================================
#define ZBX 128

template <int NT, int VAR, bool DIAG>
__global__ void __launch_bounds__(ZBX, 4) k_sweby_z(const Geom g, const SwebyArgs<NT> a)
{
    const int i = blockIdx.x * ZBX + threadIdx.x + 1;
    const int j = blockIdx.y + 1;
    if (i > g.ni) return;
    const int ks = blockIdx.z * a.kc + 1;
    const int ke = min(ks + a.kc - 1, g.nk);
    const int k0 = ks > 1 ? ks - 1 : 1;  // first face evaluated (warm-up face when the chunk starts below the surface)
    const double dtime = a.dtime;
    const size_t c2 = d2(g, i, j);
    const double dat = a.dat[c2], datr = a.datr[c2];
    const size_t slab = (size_t)g.slab;

    size_t qd = d3(g, i, j, k0);                                  // data-domain offset of level k
    size_t qt = t3(g, i, j, k0);                                  // h2 offset of level k
    size_t q2 = qd + slab * (size_t)(min(k0 + 2, g.nk) - k0);     // data-domain offset of level min(k+2, nk)
    const size_t qkm = (k0 > 1) ? qd - slab : qd, qkp = (k0 < g.nk) ? qd + slab : qd;

    unsigned nb = a.nib[qd];
    double Tk[NT], Tp1[NT], Rm1[NT], R0[NT], ftp[NT], Tp2[NT];
#pragma unroll
    for (int n = 0; n < NT; n++) {
        const double Tkm = a.T[n][qkm];
        Tk[n] = a.T[n][qd];
        Tp1[n] = a.T[n][qkp];
        Tp2[n] = a.T[n][q2];
        Rm1[n] = (Tkm - Tk[n]) * nib_and(nb, 3u);      // ((T(km1)-T(k))*m(km1))*m(k); +0 at k = 1 (km1 clamps)
        R0[n] = (Tk[n] - Tp1[n]) * nib_and(nb, 6u);    // ((T(k)-T(kp1))*m(k))*m(kp1)
        ftp[n] = 0.0;
    }
    double wkm1 = 0.0;
    double wk = a.w[qd + slab];                         // w3(k) = d3(k) + slab
    double r = a.rho[qd];

    for (int k = k0; k <= ke; k++) {
        // ---- software pipeline: operands of level k+1 ----
        const bool more = (k < ke);
        const size_t qd_n = qd + slab;
        const size_t q2_n = (k + 3 <= g.nk) ? q2 + slab : q2;
        unsigned nb_n = 0;
        double wk_n = 0.0, r_n = 1.0, Tp2_n[NT];
#pragma unroll
        for (int n = 0; n < NT; n++) Tp2_n[n] = 0.0;
        if (more) {
            nb_n = a.nib[qd_n];
            wk_n = a.w[qd_n + slab];
            r_n = a.rho[qd_n];
#pragma unroll
            for (int n = 0; n < NT; n++) Tp2_n[n] = a.T[n][q2_n];
        }
        // ---- level k ----
        const Rcp rr = make_rcp(r);
        const FaceCoef c = make_coef(dat * wk, fabs(div_rcp(wk * dtime, rr)), nib_and(nb, 6u));
        const double mm12 = nib_and(nb, 12u);           // m(kp1)*m(kp2)
        const bool live = (k >= ks);
#pragma unroll
        for (int n = 0; n < NT; n++) {
            const double Rp1 = (Tp1[n] - Tp2[n]) * mm12;
            const double fbt = sweby_flux<VAR>(c, Rm1[n], R0[n], Rp1, Tp1[n], Tk[n], a.sl);
            if (live) {
                double t;
                if (VAR == VAR_ALL) {  // OTA:4191-4195
                    const double wz = (datr * (fbt - ftp[n])) + (Tk[n] * (wkm1 - wk));
                    t = Tk[n] + div_rcp(wz * dtime, rr);
                    if (DIAG && a.dadv[n]) a.dadv[n][qd] = wz;
                } else {               // OTA:3892-3896
                    t = Tk[n] + (div_rcp(dtime, rr) * ((datr * (fbt - ftp[n])) + (Tk[n] * (wkm1 - wk))));
                }
                a.tm_in[n][qt] = t;
                if (DIAG && a.flux[n]) a.flux[n][qd] = fbt;
            }
            ftp[n] = fbt;
            Rm1[n] = R0[n];
            R0[n] = Rp1;
            Tk[n] = Tp1[n];
            Tp1[n] = Tp2[n];
            Tp2[n] = Tp2_n[n];
        }
        wkm1 = wk;
        wk = wk_n;
        r = r_n;
        nb = nb_n;
        qd = qd_n;
        q2 = q2_n;
        qt += (size_t)g.tslab;
    }
}

// =================================================================================================
// x sweep  (OTA:4251-4299 / 3916-3969)
// =================================================================================================
#define XWARPS 4   // warps per block, stacked along j

template <int NT, int VAR, bool DIAG>
__global__ void __launch_bounds__(32 * XWARPS, 4) k_sweby_x(const Geom g, const SwebyArgs<NT> a)
{
    const int lane = threadIdx.x;
    const int i = blockIdx.x * 31 + lane;        // east-face index 0..ni; lanes >= 1 also update cell i
    const int j = blockIdx.y * XWARPS + threadIdx.y + 1;
    if (j > g.nj) return;                        // whole warp leaves together
    const bool face_ok = (i <= g.ni);
    const bool cell_ok = face_ok && (lane >= 1);
    const int ic = face_ok ? i : g.ni;           // clamped index for loads of idle lanes
    const int ks = blockIdx.z * a.kc + 1;
    const int ke = min(ks + a.kc - 1, g.nk);
    const double dtime = a.dtime;
    const size_t c2 = d2(g, ic, j);
    const double dyte = a.dyte[c2], dxte = a.dxte[c2], datr = a.datr[c2];
    const size_t slab = (size_t)g.slab, tslab = (size_t)g.tslab;
    size_t q = d3(g, ic, j, ks);
    size_t tq = t3(g, ic, j, ks);
    for (int k = ks; k <= ke; k++, q += slab, tq += tslab) {
        if (k < ke) {   // next level's lines -> L2
            prefetch_l2(a.u + q + slab);
            prefetch_l2(a.rho + q + slab);
#pragma unroll
            for (int n = 0; n < NT; n++) {
                prefetch_l2(a.tm_in[n] + tq + tslab);
                prefetch_l2(a.T[n] + q + slab);
            }
        }
        const unsigned nb = a.nib[q];
        const double m_i = nib_and(nb, 2u);
        const double uu = a.u[q];
        const double rho_i = a.rho[q];
        const double rho_e = a.rho[q + 1];
        const double mf = dyte * uu;
        const FaceCoef c = make_coef(mf, fabs(div_rcp((uu * dtime) * 2.0, make_rcp((rho_i + rho_e) * dxte))), nib_and(nb, 6u));
        const double mm01 = nib_and(nb, 3u), mm23 = nib_and(nb, 12u);
        const Rcp rr = make_rcp(rho_i);
        const double mfw = __shfl_up_sync(0xffffffffu, mf, 1);
#pragma unroll
        for (int n = 0; n < NT; n++) {
            const double tm1 = a.tm_in[n][tq - 1], t0 = a.tm_in[n][tq], t1 = a.tm_in[n][tq + 1], t2 = a.tm_in[n][tq + 2];
            const double Rjp = (t2 - t1) * mm23;      // ((tm(i+2)-tm(i+1))*m(i+2))*m(i+1)
            const double Rj = (t1 - t0) * c.mm;       // ((tm(i+1)-tm(i))*m(i+1))*m(i)
            const double Rjm = (t0 - tm1) * mm01;     // ((tm(i)-tm(i-1))*m(i))*m(i-1)
            const double f = sweby_flux<VAR>(c, Rjp, Rj, Rjm, t0, t1, a.sl);
            const double fw = __shfl_up_sync(0xffffffffu, f, 1);
            if (DIAG && face_ok && a.flux[n]) a.flux[n][q] = f;
            if (cell_ok) {
                const double Tc = a.T[n][q];
                double t;
                if (VAR == VAR_ALL) {  // OTA:4288-4295
                    const double wx = (m_i * datr) * ((fw - f) + (Tc * (mf - mfw)));
                    t = t0 + div_rcp(wx * dtime, rr);
                    if (DIAG && a.dadv[n]) a.dadv[n][q] = wx;
                } else {               // OTA:3960-3965
                    t = t0 + (div_rcp((dtime * m_i) * datr, rr) * ((fw - f) + (Tc * (mf - mfw))));
                }
                a.tm_out[n][tq] = t;
            }
        }
    }
}

// =================================================================================================
// y sweep + total tendency  (OTA:4362-4432 / 3980-4056)
// =================================================================================================
#define YBX 128

template <int NT, int VAR, bool DIAG>
__global__ void __launch_bounds__(YBX, 4) k_sweby_y(const Geom g, const SwebyArgs<NT> a, const int nxt)
{
    // linear block id, k fastest
    const int lin = blockIdx.x;
    const int k = lin % g.nk + 1;
    const int rest = lin / g.nk;
    const int xt = rest % nxt, jc = rest / nxt;
    const int i = xt * YBX + threadIdx.x + 1;
    if (i > g.ni) return;
    const int js = jc * a.kc + 1;
    const int je = min(js + a.kc - 1, g.nj);
    const double dtime = a.dtime;
    const Rcp rdt = make_rcp(dtime);
    const size_t nxd = (size_t)g.nxd, tp = (size_t)g.tpitch;

    // state at face jf = js-1
    size_t q = d3(g, i, js - 1, k);        // data-domain offset of (i, jf, k)
    size_t c2 = d2(g, i, js - 1);
    size_t tq = t3(g, i, js - 1, k);
    const size_t wofs = (size_t)g.slab;    // w3(k) = d3(k) + slab ; w3(k-1) = d3(k)
    unsigned nb = a.nib[q];
    double t0[NT], t1[NT], Rm1[NT], R0[NT], fprev[NT];
#pragma unroll
    for (int n = 0; n < NT; n++) {
        const double tmm = a.tm_in[n][tq - tp];
        t0[n] = a.tm_in[n][tq];
        t1[n] = a.tm_in[n][tq + tp];
        Rm1[n] = (t0[n] - tmm) * nib_and(nb, 3u);     // ((tm(j)-tm(j-1))*m(j))*m(j-1)
        R0[n] = (t1[n] - t0[n]) * nib_and(nb, 6u);    // ((tm(j+1)-tm(j))*m(j+1))*m(j)
        fprev[n] = 0.0;
    }
    double rho0 = a.rho[q];

    for (int jf = js - 1; jf <= je; jf++, q += nxd, c2 += nxd, tq += tp) {
        if (jf < je) {   // next row's lines -> L2
            prefetch_l2(a.v + q + nxd);
            prefetch_l2(a.rho + q + 2 * nxd);
            prefetch_l2(a.u + q + nxd);
            prefetch_l2(a.w + q + nxd + wofs);
#pragma unroll
            for (int n = 0; n < NT; n++) {
                prefetch_l2(a.tm_in[n] + tq + 3 * tp);
                prefetch_l2(a.T[n] + q + nxd);
                if (a.accumulate) prefetch_l2(a.th[n] + q + nxd);
            }
        }
        const double vv = a.v[q];
        const double rho1 = a.rho[q + nxd];
        const double mf = a.dxtn[c2] * vv;
        const FaceCoef c = make_coef(mf, fabs(div_rcp((vv * dtime) * 2.0, make_rcp((rho0 + rho1) * a.dytn[c2]))), nib_and(nb, 6u));
        const double mm23 = nib_and(nb, 12u), m0 = nib_and(nb, 2u);
        const bool live = (jf >= js);
        double wdiv = 0.0, datr = 0.0;
        Rcp rr;
        rr.b = 1.0; rr.y = 1.0;
        if (live) {   // T*( (w(k)-wkm1) + datr*(dyte(i-1)*u(i-1) - dyte(i)*u(i)) )  (OTA:4402-4406)
            datr = a.datr[c2];
            const double wk = a.w[q + wofs];
            const double wkm1 = (k == 1) ? 0.0 : a.w[q];
            wdiv = (wk - wkm1) + (datr * ((a.dyte[c2 - 1] * a.u[q - 1]) - (a.dyte[c2] * a.u[q])));
            rr = make_rcp(rho0);
        }
#pragma unroll
        for (int n = 0; n < NT; n++) {
            const double t2 = a.tm_in[n][tq + 2 * tp];
            const double Rp1 = (t2 - t1[n]) * mm23;   // ((tm(j+2)-tm(j+1))*m(j+2))*m(j+1)
            const double f = sweby_flux<VAR>(c, Rp1, R0[n], Rm1[n], t0[n], t1[n], a.sl);
            if (DIAG && a.flux[n]) a.flux[n][q] = f;
            if (live) {
                const double Tc = a.T[n][q];
                double t;
                if (VAR == VAR_ALL) {  // OTA:4401-4413
                    const double wy = ((m0 * datr) * (fprev[n] - f)) + (Tc * wdiv);
                    t = t0[n] + div_rcp(wy * dtime, rr);
                    if (DIAG && a.dadv[n]) a.dadv[n][q] = wy;
                } else {               // OTA:4025-4040
                    t = t0[n] + (div_rcp((dtime * m0) * datr, rr) * (fprev[n] - f));
                    t = t + (div_rcp(dtime * Tc, rr) * wdiv);
                }
                const double adv = div_rcp(rho0 * (t - Tc), rdt) * m0;
                a.adv[n][q] = adv;
                if (a.accumulate) a.th[n][q] = a.th[n][q] + adv;
            }
            fprev[n] = f;
            Rm1[n] = R0[n];
            R0[n] = Rp1;
            t0[n] = t1[n];
            t1[n] = t2;
        }
        rho0 = rho1;
        nb = a.nib[q + nxd];
    }
}

// zero the halo ring of NT data-domain arrays (T_prog(n)%wrk1 = 0 over the data domain, OTA:4142-4148,
// 1925-1931; the compute domain is overwritten by the y sweep)
template <int NT>
struct RingArgs {
    double *p[NT];
};
template <int NT>
__global__ void k_zero_ring(const Geom g, const RingArgs<NT> a)
{
    const int ring = 2 * g.nxd + 2 * g.nj;   // per level
    const long long tot = (long long)ring * g.nk;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(q / ring) + 1;
        const int r = (int)(q % ring);
        int i, j;
        if (r < g.nxd) { i = r; j = 0; }
        else if (r < 2 * g.nxd) { i = r - g.nxd; j = g.nyd - 1; }
        else { const int s = r - 2 * g.nxd; j = 1 + s / 2; i = (s & 1) ? g.nxd - 1 : 0; }
#pragma unroll
        for (int n = 0; n < NT; n++)
            if (a.p[n]) a.p[n][d3(g, i, j, k)] = 0.0;
    }
}

// mask nibbles from the halo-2 u8 mask: dir 0 = z (clamped levels), 1 = x, 2 = y; data-domain layout
__global__ void k_build_nibbles(const Geom g, const uint8_t *__restrict__ m, uint8_t *__restrict__ nz, uint8_t *__restrict__ nx,
                                uint8_t *__restrict__ ny)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // 0..ni+1
    const int j = blockIdx.y, k = blockIdx.z + 1;          // 0..nj+1
    if (i > g.ni + 1) return;
    const size_t q = d3(g, i, j, k);
    auto M = [&](int ii, int jj, int kk) -> unsigned { return m[m3(g, ii, jj, kk)] ? 1u : 0u; };
    const int km1 = max(k - 1, 1), kp1 = min(k + 1, g.nk), kp2 = min(k + 2, g.nk);
    nz[q] = (uint8_t)(M(i, j, km1) | (M(i, j, k) << 1) | (M(i, j, kp1) << 2) | (M(i, j, kp2) << 3));
    nx[q] = (i <= g.ni) ? (uint8_t)(M(i - 1, j, k) | (M(i, j, k) << 1) | (M(i + 1, j, k) << 2) | (M(i + 2, j, k) << 3)) : 0;
    ny[q] = (j <= g.nj) ? (uint8_t)(M(i, j - 1, k) | (M(i, j, k) << 1) | (M(i, j + 1, k) << 2) | (M(i, j + 2, k) << 3)) : 0;
}
