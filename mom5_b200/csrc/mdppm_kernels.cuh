// mdppm_kernels.cuh -- advect_tracer_mdppm (OTA:5990-6494) with the edge-value limiters ppm_limit_cw84 / _ifc / _sh
// (OTA:6510-6657): the multi-dimensional piecewise-parabolic scheme (dispatcher arm ADVECT_MDPPM, OTA:1966-1968).
//
// Same sweep structure as the MDFL schemes (z, E/W update, x, N/S update, y + overall tendency) but a 9-point
// stencil per direction: the running tracer and its mask live on a halo-4 scratch ("h4", q3 indexing).  Per direction:
//   slope  da(c)   from S(c-2..c+2) and the masks (4th-order estimate blended by the masks, then Lin's monotonic constraint)
//   edges  aL, aR  from S(c), the first differences and da(c-1), da(c), da(c+1); then ONE of three point-wise limiters
//   flux           from the parabola of the UPWIND cell of the face
// One tracer per call, not the benchmarked path: plain one-thread-per-point kernels, lanes along i; every '/' is the
// compiler's IEEE division, --fmad=false.  Expression trees as in the oracle (oracle/mom5adv_oracle.c: ppm_slope, ppm_diffs,
// ppm_edges, orc_mdppm_*), which is pinned bit for bit to the reference text (tests/golden: mdppm_cw84/_ifc/_sh.*).
#pragma once

#include "mom5adv_internal.cuh"

struct PPMArgs {
    const double *T, *u, *v, *w, *rho;
    const double *m4;                   // tmask_mdppm, h4
    const double *tmask;                // Grd%tmask, data domain (y update, OTA:6427)
    const double *dat, *datr, *dxte, *dyte, *dxtn, *dytn;
    double *tr;                         // tracer_mdppm, h4
    double *da;                         // slope scratch: data-domain layout in z (dak), h4 layout in x / y (da)
    double *fx, *fy, *fz;               // data-domain flux arrays
    double *th, *wrk1;
    double dtime;
    int limiter;                        // Tracer%ppm_hlimiter: 1 cw84, 2 ifc, 3 sh (the reference uses it in all three directions)
};

#define PPM_R12 (1. / 12.)
#define PPM_TWOTHIRDS (2. / 3.)
#define PPM_FOURTHIRDS (4. / 3.)

__device__ __forceinline__ double pmax(double a, double b) { return fmax_first(a, b); }
__device__ __forceinline__ double pmin(double a, double b) { return fmin_first(a, b); }
__device__ __forceinline__ double pmax3(double a, double b, double c) { return pmax(pmax(a, b), c); }
__device__ __forceinline__ double pmin3(double a, double b, double c) { return pmin(pmin(a, b), c); }
__device__ __forceinline__ double pmax4(double a, double b, double c, double d) { return pmax(pmax3(a, b, c), d); }
__device__ __forceinline__ double pmin4(double a, double b, double c, double d) { return pmin(pmin3(a, b, c), d); }

// S, m = values / masks at c-2 .. c+2
__device__ __forceinline__ double ppm_slope(const double *S, const double *m)
{
    const double da2 = 0.5 * (S[3] - S[1]);
    const double da3m = PPM_R12 * ((((2. * S[0]) - (12. * S[1])) + (6. * S[2])) + (4. * S[3]));
    const double da3p = PPM_R12 * ((((-(4. * S[1])) - (6. * S[2])) + (12. * S[3])) - (2. * S[4]));
    const double da = (((m[0] * (1. - (0.5 * m[4]))) * da3m) + ((m[4] * (1. - (0.5 * m[0]))) * da3p)) + (((1. - m[0]) * (1. - m[4])) * da2);
    const double dMx = pmax3(S[3], S[1], S[2]) - S[2];
    const double dMn = S[2] - pmin3(S[3], S[1], S[2]);
    return (((copysign(1., da) * pmin(fabs(da), 2. * pmin(dMx, dMn))) * m[1]) * m[3]) * m[2];
}

struct PPMPar { double aL, aR, a6; };

// parabola of one cell: edge values (Lin 1994 eq. B2), limiter, curvature
__device__ __forceinline__ PPMPar ppm_parabola(int limiter, const double *S, const double *m, double da_m, double da_0, double da_p)
{
    const double Si = S[2];
    const double d1m = (Si - S[1]) * m[1], d1p = (S[3] - Si) * m[3];
    const double Sim1 = Si - d1m, Sip1 = Si + d1p;
    double aL = (0.5 * (Sim1 + Si)) + (ONESIXTH * (da_m - da_0));
    double aR = (0.5 * (Si + Sip1)) + (ONESIXTH * (da_0 - da_p));
    if (limiter == 1) {          // ppm_limit_cw84 (OTA:6520-6536)
        if ((aR - Si) * (Si - aL) <= 0.) { aL = Si; aR = Si; }
        const double da2 = aR - aL, da4 = 0.5 * (aR + aL);
        const double da3m = (6. * da2) * (Si - da4), da3p = da2 * da2;
        if (da3m > da3p) aL = (3. * Si) - (2. * aR);
        if (da3m < -da3p) aR = (3. * Si) - (2. * aL);
    } else if (limiter == 2) {   // ppm_limit_ifc (OTA:6565-6575)
        const double ada = fabs(da_0), sda = copysign(1., da_0);
        aL = Si - (sda * pmin(ada, fabs(aL - Si)));
        aR = Si + (sda * pmin(ada, fabs(aR - Si)));
    } else {                     // ppm_limit_sh (OTA:6614-6655)
        const double d1mm = ((S[1] - S[0]) * m[1]) * m[0], d1pp = ((S[4] - S[3]) * m[3]) * m[4];
        double z = d1m - d1mm;
        const double w = d1p - d1m;
        double x = (4. * w) - z, y = (4. * z) - w;
        const double dM4m = pmax(0., pmin4(x, y, z, w)) + pmin(0., pmax4(x, y, z, w));
        z = d1pp - d1p;
        x = (4. * w) - z;
        y = (4. * z) - w;
        const double dM4p = pmax(0., pmin4(x, y, z, w)) + pmin(0., pmax4(x, y, z, w));
        double qAV = 0.5 * (Si + Sip1);
        x = d1p;
        y = 3. * d1m;
        double qMP = (Si + pmax(0., pmin(x, y))) + pmin(0., pmax(x, y));
        double qUL = Si + y;
        double qLC = (Si + (0.5 * d1m)) + (PPM_FOURTHIRDS * dM4m);
        double qMD = qAV - (0.5 * dM4p);
        double qMin = pmax(pmin3(qMD, Si, Sip1), pmin3(Si, qUL, qLC));
        double qMax = pmin(pmax3(qMD, Si, Sip1), pmax3(Si, qUL, qLC));
        if ((aR - Si) * (aR - qMP) > 1.e-10) aR = pmin(pmax(aR, qMin), qMax);
        qAV = 0.5 * (Si + Sim1);
        x = -d1m;
        y = -3. * d1p;
        qMP = (Si + pmax(0., pmin(x, y))) + pmin(0., pmax(x, y));
        qUL = Si + y;
        qLC = (Si - (0.5 * d1p)) + (PPM_FOURTHIRDS * dM4p);
        qMD = qAV - (0.5 * dM4m);
        qMin = pmax(pmin3(qMD, Si, Sim1), pmin3(Si, qUL, qLC));
        qMax = pmin(pmax3(qMD, Si, Sim1), pmax3(Si, qUL, qLC));
        if ((aL - Si) * (aL - qMP) > 1.e-10) aL = pmin(pmax(aL, qMin), qMax);
    }
    PPMPar p;
    p.aL = aL; p.aR = aR;
    p.a6 = (6. * Si) - (3. * (aR + aL));
    return p;
}

// ---- z (OTA:6036-6201) ----
// values and effective masks of column cell k: levels clamp at the surface / bottom and the clamped neighbours are masked
// out by the real(k-km1) factors (OTA:6083-6087)
__device__ __forceinline__ void ppm_zcell(const Geom &g, const PPMArgs &a, int i, int j, int k, double *S, double *m)
{
    const int km2 = max(k - 2, 1), km1 = max(k - 1, 1), kp1 = min(k + 1, g.nk), kp2 = min(k + 2, g.nk);
    S[0] = a.T[d3(g, i, j, km2)]; S[1] = a.T[d3(g, i, j, km1)]; S[2] = a.T[d3(g, i, j, k)];
    S[3] = a.T[d3(g, i, j, kp1)]; S[4] = a.T[d3(g, i, j, kp2)];
    m[0] = a.m4[q3(g, i, j, km2)] * (double)(km1 - km2); m[1] = a.m4[q3(g, i, j, km1)] * (double)(k - km1);
    m[2] = a.m4[q3(g, i, j, k)];
    m[3] = a.m4[q3(g, i, j, kp1)] * (double)(kp1 - k); m[4] = a.m4[q3(g, i, j, kp2)] * (double)(kp2 - kp1);
}

__global__ void __launch_bounds__(128) k_ppm_zslope(const Geom g, const PPMArgs a)   // dak on the compute domain -> a.da (data-domain layout)
{
    const int i = blockIdx.x * 128 + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > g.ni) return;
    double S[5], m[5];
    ppm_zcell(g, a, i, j, k, S, m);
    a.da[d3(g, i, j, k)] = ppm_slope(S, m);
}

// flux through the bottom face of cell (i,j,k): from cell k's parabola when the transport is <= 0, else from cell k+1's
__global__ void __launch_bounds__(128) k_ppm_zflux(const Geom g, const PPMArgs a)
{
    const int i = blockIdx.x * 128 + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > g.ni) return;
    const double dat = a.dat[d2(g, i, j)], wk = a.w[w3(g, i, j, k)];
    double S[5], m[5];
    ppm_zcell(g, a, i, j, k, S, m);
    const double mfA = ((dat * wk) * m[2]) * m[3];
    const double cfl = fabs((wk * a.dtime) / a.rho[d3(g, i, j, k)]);
    double flux;
    if (mfA <= 0.0) {
        const int km1 = max(k - 1, 1), kp1 = min(k + 1, g.nk);
        const PPMPar p = ppm_parabola(a.limiter, S, m, a.da[d3(g, i, j, km1)], a.da[d3(g, i, j, k)], a.da[d3(g, i, j, kp1)]);
        flux = mfA * (p.aR + ((0.5 * cfl) * ((p.aL - p.aR) + ((1. - (PPM_TWOTHIRDS * cfl)) * p.a6))));
    } else {   // mfA > 0 implies k < nk; cell k+1 sees this face as its km1 face (OTA:6166-6172)
        const int kc = k + 1, kp1 = min(kc + 1, g.nk);
        ppm_zcell(g, a, i, j, kc, S, m);
        const double mfB = ((dat * wk) * m[2]) * m[1];
        const PPMPar p = ppm_parabola(a.limiter, S, m, a.da[d3(g, i, j, k)], a.da[d3(g, i, j, kc)], a.da[d3(g, i, j, kp1)]);
        flux = mfB * (p.aL + ((0.5 * cfl) * ((p.aR - p.aL) + ((1. - (PPM_TWOTHIRDS * cfl)) * p.a6))));
    }
    a.fz[d3(g, i, j, k)] = flux;
}

__global__ void __launch_bounds__(128) k_ppm_zupd(const Geom g, const PPMArgs a)     // OTA:6180-6190
{
    const int i = blockIdx.x * 128 + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > g.ni) return;
    const int km1 = max(k - 1, 1);
    const double mskm1 = (double)(k - km1), Tk = a.T[d3(g, i, j, k)];
    a.tr[q3(g, i, j, k)] = Tk + ((a.dtime / a.rho[d3(g, i, j, k)]) *
                                 ((a.datr[d2(g, i, j)] * (a.fz[d3(g, i, j, k)] - (mskm1 * a.fz[d3(g, i, j, km1)]))) +
                                  (Tk * ((mskm1 * a.w[w3(g, i, j, km1)]) - a.w[w3(g, i, j, k)]))));
}

// ---- x / y (OTA:6205-6328, 6344-6455): DIR 0 = x, 1 = y ----
template <int DIR>
__device__ __forceinline__ void ppm_hcell(const Geom &g, const PPMArgs &a, int i, int j, int k, double *S, double *m)
{
#pragma unroll
    for (int q = 0; q < 5; q++) {
        const size_t o = q3(g, i + (DIR == 0 ? q - 2 : 0), j + (DIR == 1 ? q - 2 : 0), k);
        S[q] = a.tr[o];
        m[q] = a.m4[o];
    }
}

template <int DIR>
__global__ void __launch_bounds__(128) k_ppm_hslope(const Geom g, const PPMArgs a)   // da on cells -1..n+2 along DIR -> a.da (h4 layout)
{
    const int i = blockIdx.x * 128 + threadIdx.x + (DIR == 0 ? -1 : 1), j = (int)blockIdx.y + (DIR == 1 ? -1 : 1), k = blockIdx.z + 1;
    if (i > g.ni + (DIR == 0 ? 2 : 0)) return;
    double S[5], m[5];
    ppm_hcell<DIR>(g, a, i, j, k, S, m);
    a.da[q3(g, i, j, k)] = ppm_slope(S, m);
}

template <int DIR>
__global__ void __launch_bounds__(128) k_ppm_hflux(const Geom g, const PPMArgs a)    // faces 0..n along DIR
{
    const int i = blockIdx.x * 128 + threadIdx.x + (DIR == 0 ? 0 : 1), j = (int)blockIdx.y + (DIR == 1 ? 0 : 1), k = blockIdx.z + 1;
    if (i > g.ni) return;
    const int di = DIR == 0, dj = DIR == 1;
    const size_t q = d3(g, i, j, k), c = d2(g, i, j);
    const double vv = (DIR == 0 ? a.u : a.v)[q];
    const double massflux = (DIR == 0 ? a.dyte : a.dxtn)[c] * vv;
    const double cfl = ((vv * a.dtime) * 2.0) / ((a.rho[q] + a.rho[d3(g, i + di, j + dj, k)]) * (DIR == 0 ? a.dxte : a.dytn)[c]);
    const double mm = (massflux * a.m4[q3(g, i, j, k)]) * a.m4[q3(g, i + di, j + dj, k)];
    const int ci = massflux >= 0.0 ? i : i + di, cj = massflux >= 0.0 ? j : j + dj;   // upwind cell
    double S[5], m[5];
    ppm_hcell<DIR>(g, a, ci, cj, k, S, m);
    const PPMPar p = ppm_parabola(a.limiter, S, m, a.da[q3(g, ci - di, cj - dj, k)], a.da[q3(g, ci, cj, k)], a.da[q3(g, ci + di, cj + dj, k)]);
    double flux;
    if (massflux >= 0.0) flux = mm * (p.aR + ((0.5 * cfl) * ((p.aL - p.aR) + ((1. - (PPM_TWOTHIRDS * cfl)) * p.a6))));
    else flux = mm * (p.aL - ((0.5 * cfl) * ((p.aR - p.aL) + ((1. + (PPM_TWOTHIRDS * cfl)) * p.a6))));
    (DIR == 0 ? a.fx : a.fy)[q] = flux;
}

__global__ void __launch_bounds__(128) k_ppm_xupd(const Geom g, const PPMArgs a)     // OTA:6317-6327
{
    const int i = blockIdx.x * 128 + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > g.ni) return;
    const size_t q = d3(g, i, j, k), c = d2(g, i, j), h = q3(g, i, j, k);
    a.tr[h] = a.tr[h] + ((((a.dtime * a.m4[h]) * a.datr[c]) / a.rho[q]) *
                         ((a.fx[q - 1] - a.fx[q]) + (a.T[q] * ((a.dyte[c] * a.u[q]) - (a.dyte[c - 1] * a.u[q - 1])))));
}

// y update, overall tendency (OTA:6424-6447) and the dispatcher tail: Tracer%wrk1 = -f, th_tendency += wrk1 (OTA:1966-1996)
__global__ void __launch_bounds__(128) k_ppm_yupd(const Geom g, const PPMArgs a)
{
    const int i = blockIdx.x * 128 + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > g.ni) return;
    const size_t q = d3(g, i, j, k), c = d2(g, i, j), h = q3(g, i, j, k);
    const size_t nxd = (size_t)g.nxd;
    double t = a.tr[h] + ((((a.dtime * a.tmask[q]) * a.datr[c]) / a.rho[q]) * (a.fy[q - nxd] - a.fy[q]));
    const double wkm1 = (k > 1) ? a.w[w3(g, i, j, k - 1)] : 0.0;
    t = t + (((a.dtime * a.T[q]) / a.rho[q]) *
             ((a.w[w3(g, i, j, k)] - wkm1) + (a.datr[c] * ((a.dyte[c - 1] * a.u[q - 1]) - (a.dyte[c] * a.u[q])))));
    a.tr[h] = t;
    const double f = (((-a.rho[q]) * (t - a.T[q])) / a.dtime) * a.m4[h];
    const double wrk1 = -f;
    a.wrk1[q] = wrk1;
    a.th[q] = a.th[q] + wrk1;
}

// compute-domain copy data-domain -> h4 (mask staging), zero fill of an h4 array's halo is done with cudaMemset
__global__ void k_d1_to_h4(const Geom g, const double *__restrict__ src, double *__restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i <= g.ni) dst[q3(g, i, j, k)] = src[d3(g, i, j, k)];
}
