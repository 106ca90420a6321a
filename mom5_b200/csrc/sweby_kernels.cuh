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
//            warp shuffle and tm(i-1..i+2) from the warp's staged row, so warps are independent (no block barrier).
//   y sweep: one thread per (i,k) marching north over a chunk of j with a rolling register window
//            (tm(j-1..j+2), R(j-1..j+1), flux(j-1)); blocks are ordered k-fastest so that the 2-D metrics and
//            w(k-1) of concurrently resident blocks hit in L2.
// x and y stage the next iteration's operands global -> shared with per-thread cp.async (LDGSTS) one iteration
// ahead: no register cost, no block barrier (warps stage and consume their own rows; __syncwarp only).
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

// ---- per-thread asynchronous staging (LDGSTS): global -> shared one iteration ahead, no register cost ----
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// =================================================================================================
// x sweep  (OTA:4251-4299 / 3916-3969)
// =================================================================================================
#define XWARPS 4    // warps per block, stacked along j
#define XROW 36     // staged row: tm(i_w-1 .. i_w+33) = 35 values (+1 pad)

template <int NT, int VAR, bool DIAG>
__global__ void __launch_bounds__(32 * XWARPS, 4) k_sweby_x(const Geom g, const SwebyArgs<NT> a)
{
    constexpr int NF = 2 * NT + 2;                       // tm[NT], T[NT], u, rho
    __shared__ double sm[2][XWARPS][NF][XROW];
    const int lane = threadIdx.x, wy = threadIdx.y;
    const int iw = blockIdx.x * 31;                      // first east-face index of this warp
    const int i = iw + lane;                             // east-face index 0..ni; lanes >= 1 also update cell i
    const int j = blockIdx.y * XWARPS + wy + 1;
    if (j > g.nj) return;                                // whole warp leaves together
    const bool face_ok = (i <= g.ni);
    const bool cell_ok = face_ok && (lane >= 1);
    const int ic = min(i, g.ni);                         // clamped index for the loads of idle lanes
    const int ks = blockIdx.z * a.kc + 1;
    const int ke = min(ks + a.kc - 1, g.nk);
    const double dtime = a.dtime;
    const size_t c2 = d2(g, ic, j);
    const double dyte = a.dyte[c2], dxte = a.dxte[c2], datr = a.datr[c2];
    const size_t slab = (size_t)g.slab, tslab = (size_t)g.tslab;
    size_t q = d3(g, ic, j, ks);
    // staged elements: tm element e = iw-1+lane (slot lane) and, for lanes 0..2, iw+31+lane (slot 32+lane);
    // rho element iw+lane (slot lane) and, for lane 0, iw+32 (slot 32); T, u at ic (slot lane)
    size_t tqa = t3(g, min(iw - 1 + lane, g.ni + 2), j, ks);
    size_t tqb = t3(g, min(iw + 31 + lane, g.ni + 2), j, ks);
    size_t qr = d3(g, min(iw + lane, g.ni + 1), j, ks);
    size_t qrb = d3(g, min(iw + 32, g.ni + 1), j, ks);

    auto stage = [&](int st) {
#pragma unroll
        for (int n = 0; n < NT; n++) {
            cp_async8(&sm[st][wy][n][lane], a.tm_in[n] + tqa);
            if (lane < 3) cp_async8(&sm[st][wy][n][32 + lane], a.tm_in[n] + tqb);
            cp_async8(&sm[st][wy][NT + n][lane], a.T[n] + q);
        }
        cp_async8(&sm[st][wy][2 * NT][lane], a.u + q);
        cp_async8(&sm[st][wy][2 * NT + 1][lane], a.rho + qr);
        if (lane == 0) cp_async8(&sm[st][wy][2 * NT + 1][32], a.rho + qrb);
    };
    stage(0);
    cp_async_commit();
    unsigned nb = a.nib[q];
    int st = 0;
    for (int k = ks; k <= ke; k++, st ^= 1) {
        __syncwarp();                                    // everyone is done reading the stage we are about to refill
        unsigned nb_n = 0;
        const size_t q_cur = q, tq_cur = tqa + 1;        // tm(i) lives one element right of the lane's staged element
        if (k < ke) {
            q += slab; qr += slab; qrb += slab; tqa += tslab; tqb += tslab;
            stage(st ^ 1);
            nb_n = a.nib[q];
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        double(*S)[XROW] = sm[st][wy];
        const double m_i = nib_and(nb, 2u);
        const double uu = S[2 * NT][lane];
        const double rho_i = S[2 * NT + 1][lane];
        const double rho_e = S[2 * NT + 1][lane + 1];
        const double mf = dyte * uu;
        const FaceCoef c = make_coef(mf, fabs(div_rcp((uu * dtime) * 2.0, make_rcp((rho_i + rho_e) * dxte))), nib_and(nb, 6u));
        const double mm01 = nib_and(nb, 3u), mm23 = nib_and(nb, 12u);
        const Rcp rr = make_rcp(rho_i);
        const double mfw = __shfl_up_sync(0xffffffffu, mf, 1);
#pragma unroll
        for (int n = 0; n < NT; n++) {
            const double tm1 = S[n][lane], t0 = S[n][lane + 1], t1 = S[n][lane + 2], t2 = S[n][lane + 3];
            const double Rjp = (t2 - t1) * mm23;      // ((tm(i+2)-tm(i+1))*m(i+2))*m(i+1)
            const double Rj = (t1 - t0) * c.mm;       // ((tm(i+1)-tm(i))*m(i+1))*m(i)
            const double Rjm = (t0 - tm1) * mm01;     // ((tm(i)-tm(i-1))*m(i))*m(i-1)
            const double f = sweby_flux<VAR>(c, Rjp, Rj, Rjm, t0, t1, a.sl);
            const double fw = __shfl_up_sync(0xffffffffu, f, 1);
            if (DIAG && face_ok && a.flux[n]) a.flux[n][q_cur] = f;
            if (cell_ok) {
                const double Tc = S[NT + n][lane];
                double t;
                if (VAR == VAR_ALL) {  // OTA:4288-4295
                    const double wx = (m_i * datr) * ((fw - f) + (Tc * (mf - mfw)));
                    t = t0 + div_rcp(wx * dtime, rr);
                    if (DIAG && a.dadv[n]) a.dadv[n][q_cur] = wx;
                } else {               // OTA:3960-3965
                    t = t0 + (div_rcp((dtime * m_i) * datr, rr) * ((fw - f) + (Tc * (mf - mfw))));
                }
                a.tm_out[n][tq_cur] = t;
            }
        }
        nb = nb_n;
    }
}

// =================================================================================================
// y sweep + total tendency  (OTA:4362-4432 / 3980-4056)
// =================================================================================================
#define YWARPS 4
#define YROW 33     // staged row: element 0 = (i_w - 1), elements 1..32 = the lanes' own i

template <int NT, int VAR, bool DIAG>
__global__ void __launch_bounds__(32 * YWARPS, 4) k_sweby_y(const Geom g, const SwebyArgs<NT> a, const int nxt)
{
    constexpr int NF = 3 * NT + 9;   // tm, T, th [NT each]; v, rho, u, wk, wkm1, dxtn, dytn, datr, dyte
    constexpr int F_T = NT, F_TH = 2 * NT, F_V = 3 * NT, F_RHO = F_V + 1, F_U = F_V + 2, F_WK = F_V + 3, F_WM = F_V + 4,
                  F_DXTN = F_V + 5, F_DYTN = F_V + 6, F_DATR = F_V + 7, F_DYTE = F_V + 8;
    __shared__ double sm[2][YWARPS][NF][YROW];
    // linear block id, k fastest
    const int lin = blockIdx.x;
    const int k = lin % g.nk + 1;
    const int rest = lin / g.nk;
    const int xt = rest % nxt, jc = rest / nxt;
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int i_raw = xt * (32 * YWARPS) + threadIdx.x + 1;
    if (i_raw - lane > g.ni) return;                     // the whole warp is outside: leave together
    const bool ok = (i_raw <= g.ni);
    const int i = min(i_raw, g.ni);
    const int js = jc * a.kc + 1;
    const int je = min(js + a.kc - 1, g.nj);
    const double dtime = a.dtime;
    const Rcp rdt = make_rcp(dtime);
    const size_t nxd = (size_t)g.nxd, tp = (size_t)g.tpitch;
    const size_t wofs = (size_t)g.slab;    // w3(k) = d3(k) + slab ; w3(k-1) = d3(k)
    const bool has_km1 = (k > 1);

    size_t q = d3(g, i, js - 1, k);        // data-domain offset of (i, jf, k)
    size_t c2 = d2(g, i, js - 1);
    size_t tq = t3(g, i, js - 1, k);

    auto stage = [&](int st, size_t qq, size_t cc, size_t tt) {   // operands of the iteration at face row jf (offsets of that row)
        double(*S)[YROW] = sm[st][wy];
#pragma unroll
        for (int n = 0; n < NT; n++) {
            cp_async8(&S[n][lane + 1], a.tm_in[n] + tt + 2 * tp);
            cp_async8(&S[F_T + n][lane + 1], a.T[n] + qq);
            if (a.accumulate) cp_async8(&S[F_TH + n][lane + 1], a.th[n] + qq);
        }
        cp_async8(&S[F_V][lane + 1], a.v + qq);
        cp_async8(&S[F_RHO][lane + 1], a.rho + qq + nxd);
        cp_async8(&S[F_U][lane + 1], a.u + qq);
        cp_async8(&S[F_WK][lane + 1], a.w + qq + wofs);
        if (has_km1) cp_async8(&S[F_WM][lane + 1], a.w + qq);
        cp_async8(&S[F_DXTN][lane + 1], a.dxtn + cc);
        cp_async8(&S[F_DYTN][lane + 1], a.dytn + cc);
        cp_async8(&S[F_DATR][lane + 1], a.datr + cc);
        cp_async8(&S[F_DYTE][lane + 1], a.dyte + cc);
        if (lane == 0) {
            cp_async8(&S[F_U][0], a.u + qq - 1);
            cp_async8(&S[F_DYTE][0], a.dyte + cc - 1);
        }
    };
    stage(0, q, c2, tq);
    cp_async_commit();

    // state at face jf = js-1
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

    int st = 0;
    for (int jf = js - 1; jf <= je; jf++, q += nxd, c2 += nxd, tq += tp, st ^= 1) {
        __syncwarp();
        unsigned nb_n = 0;
        if (jf < je) {
            stage(st ^ 1, q + nxd, c2 + nxd, tq + tp);
            nb_n = a.nib[q + nxd];
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        double(*S)[YROW] = sm[st][wy];
        const double vv = S[F_V][lane + 1];
        const double rho1 = S[F_RHO][lane + 1];
        const double mf = S[F_DXTN][lane + 1] * vv;
        const FaceCoef c = make_coef(mf, fabs(div_rcp((vv * dtime) * 2.0, make_rcp((rho0 + rho1) * S[F_DYTN][lane + 1]))), nib_and(nb, 6u));
        const double mm23 = nib_and(nb, 12u), m0 = nib_and(nb, 2u);
        const bool live = (jf >= js);
        double wdiv = 0.0, datr = 0.0;
        Rcp rr;
        rr.b = 1.0; rr.y = 1.0;
        if (live) {   // T*( (w(k)-wkm1) + datr*(dyte(i-1)*u(i-1) - dyte(i)*u(i)) )  (OTA:4402-4406)
            datr = S[F_DATR][lane + 1];
            const double wk = S[F_WK][lane + 1];
            const double wkm1 = has_km1 ? S[F_WM][lane + 1] : 0.0;
            wdiv = (wk - wkm1) + (datr * ((S[F_DYTE][lane] * S[F_U][lane]) - (S[F_DYTE][lane + 1] * S[F_U][lane + 1])));
            rr = make_rcp(rho0);
        }
#pragma unroll
        for (int n = 0; n < NT; n++) {
            const double t2 = S[n][lane + 1];
            const double Rp1 = (t2 - t1[n]) * mm23;   // ((tm(j+2)-tm(j+1))*m(j+2))*m(j+1)
            const double f = sweby_flux<VAR>(c, Rp1, R0[n], Rm1[n], t0[n], t1[n], a.sl);
            if (DIAG && ok && a.flux[n]) a.flux[n][q] = f;
            if (live) {
                const double Tc = S[F_T + n][lane + 1];
                double t;
                if (VAR == VAR_ALL) {  // OTA:4401-4413
                    const double wy_ = ((m0 * datr) * (fprev[n] - f)) + (Tc * wdiv);
                    t = t0[n] + div_rcp(wy_ * dtime, rr);
                    if (DIAG && ok && a.dadv[n]) a.dadv[n][q] = wy_;
                } else {               // OTA:4025-4040
                    t = t0[n] + (div_rcp((dtime * m0) * datr, rr) * (fprev[n] - f));
                    t = t + (div_rcp(dtime * Tc, rr) * wdiv);
                }
                const double adv = div_rcp(rho0 * (t - Tc), rdt) * m0;
                if (ok) {
                    a.adv[n][q] = adv;
                    if (a.accumulate) a.th[n][q] = S[F_TH + n][lane + 1] + adv;
                }
            }
            fprev[n] = f;
            Rm1[n] = R0[n];
            R0[n] = Rp1;
            t0[n] = t1[n];
            t1[n] = t2;
        }
        rho0 = rho1;
        nb = nb_n;
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
