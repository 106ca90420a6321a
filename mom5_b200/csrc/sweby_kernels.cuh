// sweby_kernels.cuh -- the three directional sweeps of the MDFL Sweby scheme as FP64 stencil kernels.
//
// Reference: advect_tracer_sweby_all   OTA:4104-4511  (VAR_ALL)
//            advect_tracer_mdfl_sweby  OTA:3806-4066  (VAR_ONE; different association order in the updates)
// OTA = src/mom5/ocean_tracers/ocean_tracer_advect.F90 of the reference tree.
//
// Thread mapping (lanes always run along i, the contiguous index, so every global access is coalesced):
//   z sweep: one thread per (i,j) column marching down k; the k-carried quantities of the reference
//            (ftp, wkm1) and the three limiter differences R(k-1), R(k), R(k+1) live in registers, so every
//            z-face flux and every difference is computed exactly once.
//   x sweep: one thread per (i,j) east face, looping over k with the 2-D metrics in registers; face fluxes and
//            mass fluxes are handed to the neighbouring cell through (double-buffered) shared memory.
//            A block of XBX threads owns XBX faces = XBX-1 cells.
//   y sweep: one thread per (i,k) marching north over a chunk of j with a rolling register window
//            (tm(j-1..j+2), R(j-1..j+1), flux(j-1)); blocks are ordered k-fastest so that the 2-D metrics and
//            w(k-1) of concurrently resident blocks hit in L2.
// All NT tracers of a group are advanced by the same thread so the tracer-independent face coefficients
// (cfl, d0, d1, (1-cfl)/(1e-30+cfl), mf+-|mf|, mask products) are computed once per face.
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
    const uint8_t *mask;        // u8, halo 2
    const double *dat, *datr, *dxte, *dyte, *dxtn, *dytn;
    double dtime, sl;
    int kc;                     // z/x: levels per k-chunk;  y: rows per j-chunk
    int accumulate;             // y: th += adv
};

__device__ __forceinline__ double mk(const uint8_t *m, size_t q) { return m[q] ? 1.0 : 0.0; }

// =================================================================================================
// z sweep  (OTA:4150-4211 / 3843-3911)
// =================================================================================================
#define ZBX 128
#define ZBY 1

template <int NT, int VAR, bool DIAG>
__global__ void __launch_bounds__(ZBX *ZBY) k_sweby_z(const Geom g, const SwebyArgs<NT> a)
{
    const int i = blockIdx.x * ZBX + threadIdx.x + 1;
    const int j = blockIdx.y * ZBY + threadIdx.y + 1;
    if (i > g.ni || j > g.nj) return;
    const int ks = blockIdx.z * a.kc + 1;
    const int ke = min(ks + a.kc - 1, g.nk);
    const int k0 = ks > 1 ? ks - 1 : 1;  // first face evaluated (warm-up face when the chunk starts below the surface)
    const double dtime = a.dtime;
    const double dat = a.dat[d2(g, i, j)], datr = a.datr[d2(g, i, j)];

    const int km = max(k0 - 1, 1), kp = min(k0 + 1, g.nk);
    const double mkm = mk(a.mask, m3(g, i, j, km));
    double m0 = mk(a.mask, m3(g, i, j, k0)), m1 = mk(a.mask, m3(g, i, j, kp));
    double Tk[NT], Tp1[NT], Rm1[NT], R0[NT], ftp[NT];
#pragma unroll
    for (int n = 0; n < NT; n++) {
        const double Tkm = a.T[n][d3(g, i, j, km)];
        Tk[n] = a.T[n][d3(g, i, j, k0)];
        Tp1[n] = a.T[n][d3(g, i, j, kp)];
        Rm1[n] = ((Tkm - Tk[n]) * mkm) * m0;   // == +0 at k0 = 1 (km1 clamps to 1)
        R0[n] = ((Tk[n] - Tp1[n]) * m0) * m1;
        ftp[n] = 0.0;
    }
    double wkm1 = 0.0;

    for (int k = k0; k <= ke; k++) {
        const int kp2 = min(k + 2, g.nk);
        const double m2 = mk(a.mask, m3(g, i, j, kp2));
        const double wk = a.w[w3(g, i, j, k)];
        const double r = a.rho[d3(g, i, j, k)];
        const FaceCoef c = make_coef(dat * wk, fabs((wk * dtime) / r), m1 * m0);
        const bool live = (k >= ks);
#pragma unroll
        for (int n = 0; n < NT; n++) {
            const double Tp2 = a.T[n][d3(g, i, j, kp2)];
            const double Rp1 = ((Tp1[n] - Tp2) * m1) * m2;
            const double fbt = sweby_flux<VAR>(c, Rm1[n], R0[n], Rp1, Tp1[n], Tk[n], a.sl);
            if (live) {
                double t;
                if (VAR == VAR_ALL) {  // OTA:4191-4195
                    const double wz = (datr * (fbt - ftp[n])) + (Tk[n] * (wkm1 - wk));
                    t = Tk[n] + ((wz * dtime) / r);
                    if (DIAG && a.dadv[n]) a.dadv[n][d3(g, i, j, k)] = wz;
                } else {               // OTA:3892-3896
                    t = Tk[n] + ((dtime / r) * ((datr * (fbt - ftp[n])) + (Tk[n] * (wkm1 - wk))));
                }
                a.tm_in[n][t3(g, i, j, k)] = t;
                if (DIAG && a.flux[n]) a.flux[n][d3(g, i, j, k)] = fbt;
            }
            ftp[n] = fbt;
            Rm1[n] = R0[n];
            R0[n] = Rp1;
            Tk[n] = Tp1[n];
            Tp1[n] = Tp2;
        }
        wkm1 = wk;
        m0 = m1;
        m1 = m2;
    }
}

// =================================================================================================
// x sweep  (OTA:4251-4299 / 3916-3969)
// =================================================================================================
#define XBX 128
#define XBY 2

template <int NT, int VAR, bool DIAG>
__global__ void __launch_bounds__(XBX *XBY) k_sweby_x(const Geom g, const SwebyArgs<NT> a)
{
    __shared__ double s_flux[2][XBY][NT][XBX];
    __shared__ double s_mf[2][XBY][XBX];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * (XBX - 1) + tx;   // east-face index 0..ni; the same thread updates cell i if tx >= 1
    const int j = blockIdx.y * XBY + ty + 1;
    const bool face_ok = (i <= g.ni) && (j <= g.nj);
    const bool cell_ok = face_ok && (tx >= 1);
    const int ks = blockIdx.z * a.kc + 1;
    const int ke = min(ks + a.kc - 1, g.nk);
    const double dtime = a.dtime;
    double dyte = 0.0, dxte = 1.0, datr = 0.0;
    if (face_ok) {
        dyte = a.dyte[d2(g, i, j)];
        dxte = a.dxte[d2(g, i, j)];
        datr = a.datr[d2(g, i, j)];
    }
    int buf = 0;
    for (int k = ks; k <= ke; k++, buf ^= 1) {
        double f[NT], t0[NT];
        double mf = 0.0, rho_i = 1.0, m_i = 0.0;
        if (face_ok) {
            const size_t mq = m3(g, i, j, k);
            const double mm1 = mk(a.mask, mq - 1), mp1 = mk(a.mask, mq + 1), mp2 = mk(a.mask, mq + 2);
            m_i = mk(a.mask, mq);
            const size_t q = d3(g, i, j, k);
            const double uu = a.u[q];
            rho_i = a.rho[q];
            const double rho_e = a.rho[q + 1];
            mf = dyte * uu;
            const FaceCoef c = make_coef(mf, fabs(((uu * dtime) * 2.0) / ((rho_i + rho_e) * dxte)), m_i * mp1);
            const size_t tq = t3(g, i, j, k);
#pragma unroll
            for (int n = 0; n < NT; n++) {
                const double tm1 = a.tm_in[n][tq - 1], t1 = a.tm_in[n][tq + 1], t2 = a.tm_in[n][tq + 2];
                t0[n] = a.tm_in[n][tq];
                const double Rjp = ((t2 - t1) * mp2) * mp1;
                const double Rj = ((t1 - t0[n]) * mp1) * m_i;
                const double Rjm = ((t0[n] - tm1) * m_i) * mm1;
                f[n] = sweby_flux<VAR>(c, Rjp, Rj, Rjm, t0[n], t1, a.sl);
                s_flux[buf][ty][n][tx] = f[n];
                if (DIAG && a.flux[n]) a.flux[n][q] = f[n];
            }
            s_mf[buf][ty][tx] = mf;
        }
        __syncthreads();
        if (cell_ok) {
            const size_t q = d3(g, i, j, k);
            const double mfw = s_mf[buf][ty][tx - 1];
#pragma unroll
            for (int n = 0; n < NT; n++) {
                const double fw = s_flux[buf][ty][n][tx - 1];
                const double Tc = a.T[n][q];
                double t;
                if (VAR == VAR_ALL) {  // OTA:4288-4295
                    const double wx = (m_i * datr) * ((fw - f[n]) + (Tc * (mf - mfw)));
                    t = t0[n] + ((wx * dtime) / rho_i);
                    if (DIAG && a.dadv[n]) a.dadv[n][q] = wx;
                } else {               // OTA:3960-3965
                    t = t0[n] + ((((dtime * m_i) * datr) / rho_i) * ((fw - f[n]) + (Tc * (mf - mfw))));
                }
                a.tm_out[n][t3(g, i, j, k)] = t;
            }
        }
    }
}

// =================================================================================================
// y sweep + total tendency  (OTA:4362-4432 / 3980-4056)
// =================================================================================================
#define YBX 128

template <int NT, int VAR, bool DIAG>
__global__ void __launch_bounds__(YBX) k_sweby_y(const Geom g, const SwebyArgs<NT> a, const int nxt)
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

    // state at face jf = js-1
    int jf = js - 1;
    size_t mq = m3(g, i, jf, k);
    const double mm1 = mk(a.mask, mq - g.mpitch);
    double m0 = mk(a.mask, mq), m1 = mk(a.mask, mq + g.mpitch);
    size_t tq = t3(g, i, jf, k);
    double t0[NT], t1[NT], Rm1[NT], R0[NT], fprev[NT];
#pragma unroll
    for (int n = 0; n < NT; n++) {
        const double tmm = a.tm_in[n][tq - g.tpitch];
        t0[n] = a.tm_in[n][tq];
        t1[n] = a.tm_in[n][tq + g.tpitch];
        Rm1[n] = ((t0[n] - tmm) * m0) * mm1;
        R0[n] = ((t1[n] - t0[n]) * m1) * m0;
        fprev[n] = 0.0;
    }
    double rho0 = a.rho[d3(g, i, jf, k)];

    for (; jf <= je; jf++) {
        const size_t q = d3(g, i, jf, k);
        const double m2 = mk(a.mask, m3(g, i, jf + 2, k));
        const double vv = a.v[q];
        const double rho1 = a.rho[q + g.nxd];
        const double mf = a.dxtn[d2(g, i, jf)] * vv;
        const FaceCoef c = make_coef(mf, fabs(((vv * dtime) * 2.0) / ((rho0 + rho1) * a.dytn[d2(g, i, jf)])), m0 * m1);
        const bool live = (jf >= js);
        double wdiv = 0.0, datr = 0.0;
        if (live) {   // T*( (w(k)-wkm1) + datr*(dyte(i-1)*u(i-1) - dyte(i)*u(i)) )  (OTA:4402-4406)
            datr = a.datr[d2(g, i, jf)];
            const double wk = a.w[w3(g, i, jf, k)];
            const double wkm1 = (k == 1) ? 0.0 : a.w[w3(g, i, jf, k - 1)];
            wdiv = (wk - wkm1) + (datr * ((a.dyte[d2(g, i - 1, jf)] * a.u[q - 1]) - (a.dyte[d2(g, i, jf)] * a.u[q])));
        }
        const size_t tq2 = t3(g, i, jf + 2, k);
#pragma unroll
        for (int n = 0; n < NT; n++) {
            const double t2 = a.tm_in[n][tq2];
            const double Rp1 = ((t2 - t1[n]) * m2) * m1;
            const double f = sweby_flux<VAR>(c, Rp1, R0[n], Rm1[n], t0[n], t1[n], a.sl);
            if (DIAG && a.flux[n]) a.flux[n][q] = f;
            if (live) {
                const double Tc = a.T[n][q];
                double t;
                if (VAR == VAR_ALL) {  // OTA:4401-4413
                    const double wy = ((m0 * datr) * (fprev[n] - f)) + (Tc * wdiv);
                    t = t0[n] + ((wy * dtime) / rho0);
                    if (DIAG && a.dadv[n]) a.dadv[n][q] = wy;
                } else {               // OTA:4025-4040
                    t = t0[n] + ((((dtime * m0) * datr) / rho0) * (fprev[n] - f));
                    t = t + (((dtime * Tc) / rho0) * wdiv);
                }
                const double adv = ((rho0 * (t - Tc)) / dtime) * m0;
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
        m0 = m1;
        m1 = m2;
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
