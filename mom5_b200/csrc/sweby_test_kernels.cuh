// sweby_test_kernels.cuh -- advect_tracer_mdfl_sweby_test (OTA:3469-3746): the mass-weighted MDFL Sweby variant the
// reference authors recommend over the legacy one (OTA:3757-3767, 4077-4089); dispatcher arms ADVECT_MDFL_SWEBY_TEST
// (sweby_limiter = 1) and ADVECT_DST_LINEAR_TEST (sweby_limiter = 0), OTA:1970-1975.
//
// Differences from advect_tracer_mdfl_sweby: three running fields on the halo-2 scratch -- tracer_mdfl (tr),
// tracermass_mdfl (tms), mass_mdfl (ms) -- all three exchanged between the sweeps; the CFL number is
// |massflux|*dtime / mass of the UPWIND cell; theta's denominator is sign(1e-30,Rj)+Rj; the tracer is
// tracermass/mass wherever mass > 0.
//
// One tracer per call and not the benchmarked path: plain one-thread-per-point kernels (lanes along i, coalesced),
// every '/' is the compiler's IEEE division, --fmad=false; same expression trees as the oracle (oracle/mom5adv_oracle.c,
// orc_sweby_test_*), pinned by the reference-text goldens (tests/golden/*.npz, keys mdfl_sweby_test.* / dst_linear_test.*).
#pragma once

#include "mom5adv_internal.cuh"

struct STArgs {
    const double *T, *u, *v, *w, *rho;
    const uint8_t *mask;                // tmask_mdfl as u8, halo 2 (m3 layout)
    const double *dat, *datr, *dyte, *dxtn;
    double *tr, *tms, *ms;              // h2 scratch
    double *fx, *fy, *fz;               // data-domain flux arrays (fz may be null)
    double *th, *wrk1;
    double dtime, sl;
};

__device__ __forceinline__ double st_m(const Geom &g, const uint8_t *m, int i, int j, int k) { return m[m3(g, i, j, k)] ? 1.0 : 0.0; }

__device__ __forceinline__ double st_flux(double Rjp, double Rj, double Rjm, double mf, double cfl, double Tup, double Tdn,
                                          double mA, double mB, double sl)
{
    const double d0 = ((2.0 - cfl) * (1.0 - cfl)) * ONESIXTH;
    const double d1 = (1.0 - (cfl * cfl)) * ONESIXTH;
    const double den = copysign(1.0e-30, Rj) + Rj;       // sign(1.0e-30,Rj) + Rj
    const double thetaP = Rjm / den, thetaM = Rjp / den;
    const double rr = (1.0 - cfl) / (1.0e-30 + cfl);
    double psiP = d0 + (d1 * thetaP);
    psiP = (psiP * (1.0 - sl)) + (fmax_first(0.0, fmin_first(fmin_first(1.0, d0 + (d1 * thetaP)), rr * thetaP)) * sl);
    double psiM = d0 + (d1 * thetaM);
    psiM = (psiM * (1.0 - sl)) + (fmax_first(0.0, fmin_first(fmin_first(1.0, d0 + (d1 * thetaM)), rr * thetaM)) * sl);
    return ((0.5 * (((mf + fabs(mf)) * (Tup + (psiP * Rj))) + ((mf - fabs(mf)) * (Tdn - (psiM * Rj))))) * mA) * mB;
}

// upwind-mass CFL (OTA:3534-3540, 3602-3608, 3672-3678): mass_a is tested with > 0 first, then mass_b with < 0
__device__ __forceinline__ double st_cfl(double mf, double absnum, double dtime, double mass_a, double mass_b)
{
    if (mf * mass_a > 0.0) return (absnum * dtime) / mass_a;
    if (mf * mass_b < 0.0) return (absnum * dtime) / mass_b;
    return 0.0;
}

// z sweep (OTA:3505-3580): one thread per (i,j) column marching down k
__global__ void __launch_bounds__(128) k_st_z(const Geom g, const STArgs a)
{
    const int i = blockIdx.x * 128 + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.ni) return;
    const double dat = a.dat[d2(g, i, j)];
    double ftp = 0.0, wkm1 = 0.0;
    for (int k = 1; k <= g.nk; k++) {
        const int kp1 = min(k + 1, g.nk), kp2 = min(k + 2, g.nk), km1 = max(k - 1, 1);
        const double Tk = a.T[d3(g, i, j, k)], Tkp1 = a.T[d3(g, i, j, kp1)];
        const double mk = st_m(g, a.mask, i, j, k), mkp1 = st_m(g, a.mask, i, j, kp1);
        double ms = a.rho[d3(g, i, j, k)] * dat;
        double tms = ms * Tk;
        double tr = Tk;
        const double Rjp = ((a.T[d3(g, i, j, km1)] - Tk) * st_m(g, a.mask, i, j, km1)) * mk;
        const double Rj = ((Tk - Tkp1) * mk) * mkp1;
        const double Rjm = ((Tkp1 - a.T[d3(g, i, j, kp2)]) * mkp1) * st_m(g, a.mask, i, j, kp2);
        const double wk = a.w[w3(g, i, j, k)];
        const double mf = dat * wk;
        const double cfl = st_cfl(mf, fabs(wk), a.dtime, a.rho[d3(g, i, j, kp1)], a.rho[d3(g, i, j, k)]);
        const double fbt = st_flux(Rjp, Rj, Rjm, mf, cfl, Tkp1, Tk, mkp1, mk, a.sl);
        ms = ms + ((a.dtime * dat) * (wk - wkm1));
        tms = tms + (a.dtime * (fbt - ftp));
        if (ms > 0.) tr = tms / ms;
        const size_t h = t3(g, i, j, k);
        a.tr[h] = tr; a.tms[h] = tms; a.ms[h] = ms;
        if (a.fz) a.fz[d3(g, i, j, k)] = fbt;
        ftp = fbt;
        wkm1 = wk;
    }
}

// east-face fluxes (OTA:3588-3632): faces i = 0..ni
__global__ void __launch_bounds__(128) k_st_xflux(const Geom g, const STArgs a)
{
    const int i = blockIdx.x * 128 + threadIdx.x, j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > g.ni) return;
    const double t0 = a.tr[t3(g, i, j, k)], t1 = a.tr[t3(g, i + 1, j, k)];
    const double m0 = st_m(g, a.mask, i, j, k), m1 = st_m(g, a.mask, i + 1, j, k);
    const double Rjp = ((a.tr[t3(g, i + 2, j, k)] - t1) * st_m(g, a.mask, i + 2, j, k)) * m1;
    const double Rj = ((t1 - t0) * m1) * m0;
    const double Rjm = ((t0 - a.tr[t3(g, i - 1, j, k)]) * m0) * st_m(g, a.mask, i - 1, j, k);
    const double mf = a.dyte[d2(g, i, j)] * a.u[d3(g, i, j, k)];
    const double cfl = st_cfl(mf, fabs(mf), a.dtime, a.ms[t3(g, i, j, k)], a.ms[t3(g, i + 1, j, k)]);
    a.fx[d3(g, i, j, k)] = st_flux(Rjp, Rj, Rjm, mf, cfl, t0, t1, m0, m1, a.sl);
}

// x update (OTA:3636-3649)
__global__ void __launch_bounds__(128) k_st_xupd(const Geom g, const STArgs a)
{
    const int i = blockIdx.x * 128 + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > g.ni) return;
    const size_t h = t3(g, i, j, k), q = d3(g, i, j, k), c = d2(g, i, j);
    const double ms = a.ms[h] + (a.dtime * ((a.dyte[c - 1] * a.u[q - 1]) - (a.dyte[c] * a.u[q])));
    const double tms = a.tms[h] + (a.dtime * (a.fx[q - 1] - a.fx[q]));
    a.ms[h] = ms; a.tms[h] = tms;
    if (ms > 0.) a.tr[h] = tms / ms;
}

// north-face fluxes (OTA:3658-3702): faces j = 0..nj
__global__ void __launch_bounds__(128) k_st_yflux(const Geom g, const STArgs a)
{
    const int i = blockIdx.x * 128 + threadIdx.x + 1, j = blockIdx.y, k = blockIdx.z + 1;
    if (i > g.ni) return;
    const double t0 = a.tr[t3(g, i, j, k)], t1 = a.tr[t3(g, i, j + 1, k)];
    const double m0 = st_m(g, a.mask, i, j, k), m1 = st_m(g, a.mask, i, j + 1, k);
    const double Rjp = ((a.tr[t3(g, i, j + 2, k)] - t1) * st_m(g, a.mask, i, j + 2, k)) * m1;
    const double Rj = ((t1 - t0) * m1) * m0;
    const double Rjm = ((t0 - a.tr[t3(g, i, j - 1, k)]) * m0) * st_m(g, a.mask, i, j - 1, k);
    const double mf = a.dxtn[d2(g, i, j)] * a.v[d3(g, i, j, k)];
    const double cfl = st_cfl(mf, fabs(mf), a.dtime, a.ms[t3(g, i, j, k)], a.ms[t3(g, i, j + 1, k)]);
    a.fy[d3(g, i, j, k)] = st_flux(Rjp, Rj, Rjm, mf, cfl, t0, t1, m0, m1, a.sl);
}

// y update + overall tendency (OTA:3706-3736) + the dispatcher tail: Tracer%wrk1 = -f, th_tendency += wrk1 (OTA:1970-1996)
__global__ void __launch_bounds__(128) k_st_yupd(const Geom g, const STArgs a)
{
    const int i = blockIdx.x * 128 + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > g.ni) return;
    const size_t h = t3(g, i, j, k), q = d3(g, i, j, k), c = d2(g, i, j);
    const size_t nxd = (size_t)g.nxd;
    const double ms = a.ms[h] + (a.dtime * ((a.dxtn[c - nxd] * a.v[q - nxd]) - (a.dxtn[c] * a.v[q])));
    const double tms = a.tms[h] + (a.dtime * (a.fy[q - nxd] - a.fy[q]));
    a.ms[h] = ms; a.tms[h] = tms;
    if (ms > 0.) a.tr[h] = tms / ms;
    const double f = (st_m(g, a.mask, i, j, k) * ((a.rho[q] * a.T[q]) - (tms * a.datr[c]))) / a.dtime;
    const double wrk1 = -f;
    a.wrk1[q] = wrk1;
    a.th[q] = a.th[q] + wrk1;
}
