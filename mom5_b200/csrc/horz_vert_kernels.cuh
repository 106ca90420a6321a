// horz_vert_kernels.cuh -- quicker and first-order upwind flux paths (the schemes that share the dispatcher
// with MDFL Sweby) + quicker_init weights.
//
// Reference (OTA = src/mom5/ocean_tracers/ocean_tracer_advect.F90):
//   quicker_init                 OTA:1442-1586      horz_advect_tracer_quicker  OTA:2538-2653
//   vert_advect_tracer_quicker   OTA:2981-3031      horz_advect_tracer_upwind   OTA:2238-2294
//   vert_advect_tracer_upwind    OTA:2792-2824      dispatcher tail             OTA:1990-1996, 2162-2168
// These paths are pure streaming (no divisions, ~30 flops per face): plain coalesced one-thread-per-point
// kernels, lanes along i.
#pragma once

#include "mom5adv_internal.cuh"

struct QuickW {   // device arrays; 2-D ones in data-domain layout with a trailing component index
    double *quick_x, *quick_y;                       // (nxd, nyd, 2)
    double *curv_xp, *curv_xn, *curv_yp, *curv_yn;   // (nxd, nyd, 3)
    double *quick_z, *curv_zp, *curv_zn;             // (nk,2) (nk,3) (nk,3)
};

static int alloc_quickw(QuickW &q, const Geom &g)
{
    const size_t n2 = (size_t)g.slab;
    double **p2[] = {&q.quick_x, &q.quick_y};
    double **p3[] = {&q.curv_xp, &q.curv_xn, &q.curv_yp, &q.curv_yn};
    for (auto p : p2) { if (cudaMalloc(p, 2 * n2 * sizeof(double)) != cudaSuccess) return 1; cudaMemset(*p, 0, 2 * n2 * sizeof(double)); }
    for (auto p : p3) { if (cudaMalloc(p, 3 * n2 * sizeof(double)) != cudaSuccess) return 1; cudaMemset(*p, 0, 3 * n2 * sizeof(double)); }
    if (cudaMalloc(&q.quick_z, 2 * g.nk * sizeof(double)) != cudaSuccess) return 1;
    if (cudaMalloc(&q.curv_zp, 3 * g.nk * sizeof(double)) != cudaSuccess) return 1;
    if (cudaMalloc(&q.curv_zn, 3 * g.nk * sizeof(double)) != cudaSuccess) return 1;
    return 0;
}
static void free_quickw(QuickW &q)
{
    for (double *p : {q.quick_x, q.quick_y, q.curv_xp, q.curv_xn, q.curv_yp, q.curv_yn, q.quick_z, q.curv_zp, q.curv_zn})
        if (p) cudaFree(p);
}

// OTA:1516-1564; dx, dy are one-level h2 fields (dxt_quick, dyt_quick)
__global__ void k_quicker_weights(const Geom g, const double *__restrict__ dx, const double *__restrict__ dy, const QuickW q)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // 0..ni
    const int j = blockIdx.y;                              // 0..nj
    if (i > g.ni) return;
    const size_t c = d2(g, i, j), n2 = (size_t)g.slab;
    const double xm = dx[t3(g, i - 1, j, 1)], x0 = dx[t3(g, i, j, 1)], x1 = dx[t3(g, i + 1, j, 1)], x2 = dx[t3(g, i + 2, j, 1)];
    const double ym = dy[t3(g, i, j - 1, 1)], y0 = dy[t3(g, i, j, 1)], y1 = dy[t3(g, i, j + 1, 1)], y2 = dy[t3(g, i, j + 2, 1)];
    q.quick_x[c] = x1 / (x1 + x0);
    q.quick_x[c + n2] = x0 / (x1 + x0);
    q.quick_y[c] = y1 / (y1 + y0);
    q.quick_y[c + n2] = y0 / (y1 + y0);
    q.curv_xp[c] = (x0 * x1) / (((xm + (2.0 * x0)) + x1) * (x0 + x1));
    q.curv_xp[c + n2] = -((x0 * x1) / ((x0 + x1) * (xm + x0)));
    q.curv_xp[c + 2 * n2] = (x0 * x1) / (((xm + (2.0 * x0)) + x1) * (xm + x0));
    q.curv_xn[c] = (x0 * x1) / (((x0 + (2.0 * x1)) + x2) * (x1 + x2));
    q.curv_xn[c + n2] = -((x0 * x1) / ((x1 + x2) * (x0 + x1)));
    q.curv_xn[c + 2 * n2] = (x0 * x1) / (((x0 + (2.0 * x1)) + x2) * (x0 + x1));
    q.curv_yp[c] = (y0 * y1) / (((ym + (2.0 * y0)) + y1) * (y0 + y1));
    q.curv_yp[c + n2] = -((y0 * y1) / ((y0 + y1) * (ym + y0)));
    q.curv_yp[c + 2 * n2] = (y0 * y1) / (((ym + (2.0 * y0)) + y1) * (ym + y0));
    q.curv_yn[c] = (y0 * y1) / (((y0 + (2.0 * y1)) + y2) * (y1 + y2));
    q.curv_yn[c + n2] = -((y0 * y1) / ((y1 + y2) * (y0 + y1)));
    q.curv_yn[c + 2 * n2] = (y0 * y1) / (((y0 + (2.0 * y1)) + y2) * (y0 + y1));
}

// ---- horizontal fluxes: thread (i,j,k), i = 0..ni, j = 0..nj; flux_x for j >= 1, flux_y for i >= 1 ----
template <bool QUICKER>
__global__ void __launch_bounds__(128)
k_horz_flux(const Geom g, const QuickW q, const double *__restrict__ tmask, const uint8_t *__restrict__ mq,
            const double *__restrict__ dyte, const double *__restrict__ dxtn, const double *__restrict__ Tm1,
            const double *__restrict__ Tt, const double *__restrict__ tq, const double *__restrict__ tlimit, const int limit,
            const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ fx, double *__restrict__ fy)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, k = blockIdx.z + 1;
    if (i > g.ni) return;
    const size_t c = d3(g, i, j, k), c2 = d2(g, i, j), n2 = (size_t)g.slab;
    if (j >= 1) {   // east face of (i,j)
        double f;
        bool up = !QUICKER;
        if (QUICKER && limit) up = (tlimit[c] == 1.0);
        if (!up) {   // OTA:2592-2604
            const double vel = dyte[c2] * u[c];
            const double upos = 0.5 * (vel + fabs(vel)), uneg = 0.5 * (vel - fabs(vel));
            const size_t mi = m3(g, i, j, k), ti = t3(g, i, j, k);
            const double m_m1 = mq[mi - 1] ? 1.0 : 0.0, m_0 = mq[mi] ? 1.0 : 0.0, m_1 = mq[mi + 1] ? 1.0 : 0.0, m_2 = mq[mi + 2] ? 1.0 : 0.0;
            const double eastmsk = m_0 * (1.0 - m_m1), westmsk = m_1 * (1.0 - m_2);
            const double t0 = tq[ti], t1 = tq[ti + 1];
            f = ((vel * ((q.quick_x[c2] * Tt[c]) + (q.quick_x[c2 + n2] * Tt[c + 1]))) -
                 (upos * (((q.curv_xp[c2] * t1) + (q.curv_xp[c2 + n2] * t0)) +
                          (q.curv_xp[c2 + 2 * n2] * ((tq[ti - 1] * (1.0 - eastmsk)) + (t0 * eastmsk)))))) -
                (uneg * (((q.curv_xn[c2] * ((tq[ti + 2] * (1.0 - westmsk)) + (t1 * westmsk))) + (q.curv_xn[c2 + n2] * t1)) +
                         (q.curv_xn[c2 + 2 * n2] * t0)));
        } else if (QUICKER) {   // OTA:2613-2620: vel = u (not dyte*u), upos = 0.5*(vel+|vel|)
            const double vel = u[c];
            const double upos = 0.5 * (vel + fabs(vel)), uneg = 0.5 * (vel - fabs(vel));
            f = ((dyte[c2] * ((upos * Tm1[c]) + (uneg * Tm1[c + 1]))) * tmask[c]) * tmask[c + 1];
        } else {                // OTA:2261-2265: velocity = 0.5*u, upos = velocity+|velocity|
            const double velocity = 0.5 * u[c];
            const double upos = velocity + fabs(velocity), uneg = velocity - fabs(velocity);
            f = ((dyte[c2] * ((upos * Tm1[c]) + (uneg * Tm1[c + 1]))) * tmask[c]) * tmask[c + 1];
        }
        fx[c] = f;
    }
    if (i >= 1) {   // north face of (i,j)
        double f;
        bool up = !QUICKER;
        if (QUICKER && limit) up = (tlimit[c] == 1.0);
        if (!up) {   // OTA:2574-2586
            const double vel = dxtn[c2] * v[c];
            const double upos = 0.5 * (vel + fabs(vel)), uneg = 0.5 * (vel - fabs(vel));
            const size_t mi = m3(g, i, j, k), ti = t3(g, i, j, k);
            const size_t mp = g.mpitch, tp = g.tpitch;
            const double m_m1 = mq[mi - mp] ? 1.0 : 0.0, m_0 = mq[mi] ? 1.0 : 0.0, m_1 = mq[mi + mp] ? 1.0 : 0.0, m_2 = mq[mi + 2 * mp] ? 1.0 : 0.0;
            const double rnormsk = m_0 * (1.0 - m_m1), soutmsk = m_1 * (1.0 - m_2);
            const double t0 = tq[ti], t1 = tq[ti + tp];
            f = ((vel * ((q.quick_y[c2] * Tt[c]) + (q.quick_y[c2 + n2] * Tt[c + g.nxd]))) -
                 (upos * (((q.curv_yp[c2] * t1) + (q.curv_yp[c2 + n2] * t0)) +
                          (q.curv_yp[c2 + 2 * n2] * ((tq[ti - tp] * (1.0 - rnormsk)) + (t0 * rnormsk)))))) -
                (uneg * (((q.curv_yn[c2] * ((tq[ti + 2 * tp] * (1.0 - soutmsk)) + (t1 * soutmsk))) + (q.curv_yn[c2 + n2] * t1)) +
                         (q.curv_yn[c2 + 2 * n2] * t0)));
        } else if (QUICKER) {   // OTA:2625-2631
            const double vel = v[c];
            const double upos = 0.5 * (vel + fabs(vel)), uneg = 0.5 * (vel - fabs(vel));
            f = ((dxtn[c2] * ((upos * Tm1[c]) + (uneg * Tm1[c + g.nxd]))) * tmask[c]) * tmask[c + g.nxd];
        } else {                // OTA:2273-2277
            const double velocity = 0.5 * v[c];
            const double upos = velocity + fabs(velocity), uneg = velocity - fabs(velocity);
            f = ((dxtn[c2] * ((upos * Tm1[c]) + (uneg * Tm1[c + g.nxd]))) * tmask[c]) * tmask[c + g.nxd];
        }
        fy[c] = f;
    }
}

// ---- the same fluxes as device functions of ONE face, for the fused pass below --------------------------------------
struct HorzIn {
    const double *tmask, *dyte, *dxtn, *Tm1, *Tt, *tq, *tlimit, *u, *v;
    const uint8_t *mq;
    int limit;
};
template <bool QUICKER>
__device__ __forceinline__ double east_flux(const Geom &g, const QuickW &q, const HorzIn &a, int i, int j, int k)
{
    const size_t c = d3(g, i, j, k), c2 = d2(g, i, j), n2 = (size_t)g.slab;
    bool up = !QUICKER;
    if (QUICKER && a.limit) up = (a.tlimit[c] == 1.0);
    if (!up) {   // OTA:2592-2604
        const double vel = a.dyte[c2] * a.u[c];
        const double upos = 0.5 * (vel + fabs(vel)), uneg = 0.5 * (vel - fabs(vel));
        const size_t mi = m3(g, i, j, k), ti = t3(g, i, j, k);
        const double m_m1 = a.mq[mi - 1] ? 1.0 : 0.0, m_0 = a.mq[mi] ? 1.0 : 0.0, m_1 = a.mq[mi + 1] ? 1.0 : 0.0, m_2 = a.mq[mi + 2] ? 1.0 : 0.0;
        const double eastmsk = m_0 * (1.0 - m_m1), westmsk = m_1 * (1.0 - m_2);
        const double t0 = a.tq[ti], t1 = a.tq[ti + 1];
        return ((vel * ((q.quick_x[c2] * a.Tt[c]) + (q.quick_x[c2 + n2] * a.Tt[c + 1]))) -
                (upos * (((q.curv_xp[c2] * t1) + (q.curv_xp[c2 + n2] * t0)) +
                         (q.curv_xp[c2 + 2 * n2] * ((a.tq[ti - 1] * (1.0 - eastmsk)) + (t0 * eastmsk)))))) -
               (uneg * (((q.curv_xn[c2] * ((a.tq[ti + 2] * (1.0 - westmsk)) + (t1 * westmsk))) + (q.curv_xn[c2 + n2] * t1)) +
                        (q.curv_xn[c2 + 2 * n2] * t0)));
    }
    if (QUICKER) {   // OTA:2613-2620: vel = u (not dyte*u), upos = 0.5*(vel+|vel|)
        const double vel = a.u[c];
        const double upos = 0.5 * (vel + fabs(vel)), uneg = 0.5 * (vel - fabs(vel));
        return ((a.dyte[c2] * ((upos * a.Tm1[c]) + (uneg * a.Tm1[c + 1]))) * a.tmask[c]) * a.tmask[c + 1];
    }
    const double velocity = 0.5 * a.u[c];   // OTA:2261-2265: velocity = 0.5*u, upos = velocity+|velocity|
    const double upos = velocity + fabs(velocity), uneg = velocity - fabs(velocity);
    return ((a.dyte[c2] * ((upos * a.Tm1[c]) + (uneg * a.Tm1[c + 1]))) * a.tmask[c]) * a.tmask[c + 1];
}
template <bool QUICKER>
__device__ __forceinline__ double north_flux(const Geom &g, const QuickW &q, const HorzIn &a, int i, int j, int k)
{
    const size_t c = d3(g, i, j, k), c2 = d2(g, i, j), n2 = (size_t)g.slab;
    bool up = !QUICKER;
    if (QUICKER && a.limit) up = (a.tlimit[c] == 1.0);
    if (!up) {   // OTA:2574-2586
        const double vel = a.dxtn[c2] * a.v[c];
        const double upos = 0.5 * (vel + fabs(vel)), uneg = 0.5 * (vel - fabs(vel));
        const size_t mi = m3(g, i, j, k), ti = t3(g, i, j, k);
        const size_t mp = g.mpitch, tp = g.tpitch;
        const double m_m1 = a.mq[mi - mp] ? 1.0 : 0.0, m_0 = a.mq[mi] ? 1.0 : 0.0, m_1 = a.mq[mi + mp] ? 1.0 : 0.0, m_2 = a.mq[mi + 2 * mp] ? 1.0 : 0.0;
        const double rnormsk = m_0 * (1.0 - m_m1), soutmsk = m_1 * (1.0 - m_2);
        const double t0 = a.tq[ti], t1 = a.tq[ti + tp];
        return ((vel * ((q.quick_y[c2] * a.Tt[c]) + (q.quick_y[c2 + n2] * a.Tt[c + g.nxd]))) -
                (upos * (((q.curv_yp[c2] * t1) + (q.curv_yp[c2 + n2] * t0)) +
                         (q.curv_yp[c2 + 2 * n2] * ((a.tq[ti - tp] * (1.0 - rnormsk)) + (t0 * rnormsk)))))) -
               (uneg * (((q.curv_yn[c2] * ((a.tq[ti + 2 * tp] * (1.0 - soutmsk)) + (t1 * soutmsk))) + (q.curv_yn[c2 + n2] * t1)) +
                        (q.curv_yn[c2 + 2 * n2] * t0)));
    }
    if (QUICKER) {   // OTA:2625-2631
        const double vel = a.v[c];
        const double upos = 0.5 * (vel + fabs(vel)), uneg = 0.5 * (vel - fabs(vel));
        return ((a.dxtn[c2] * ((upos * a.Tm1[c]) + (uneg * a.Tm1[c + g.nxd]))) * a.tmask[c]) * a.tmask[c + g.nxd];
    }
    const double velocity = 0.5 * a.v[c];   // OTA:2273-2277
    const double upos = velocity + fabs(velocity), uneg = velocity - fabs(velocity);
    return ((a.dxtn[c2] * ((upos * a.Tm1[c]) + (uneg * a.Tm1[c + g.nxd]))) * a.tmask[c]) * a.tmask[c + g.nxd];
}

// ---- flux, fold line and divergence in ONE pass (horz_advect_tracer_quicker / _upwind as the dispatcher uses them) -----------
// A warp owns 32 consecutive east faces (i_w .. i_w+31) of one (j,k) row = 31 cells, as the Sweby x sweep does: the west flux of a
// cell comes from the neighbouring lane by shuffle; the south flux is recomputed (no flux array is materialised, no memset,
// no second pass).  On the folded north edge of a tripolar grid the reference makes the NORTH flux of the top row antisymmetric
// after the fact (OTA:2640, mpp_update_domains(flux_x, flux_y, Dom_flux, gridtype=CGRID_NE)): flux_y(i, nj) := -flux_y(ni+1-i, nj)
// for the eastern half.  With the whole fold line on this rank (layout_x = 1) that value is simply computed at the mirror
// column; layouts that split the fold line keep the three-kernel path with its strip exchange (fold_line_fix).
// flux_x / flux_y are written only when the caller asked for them (diagnostics): the same points the reference defines,
// flux_x for i = 0..ni, flux_y for j = 0..nj.
#define HFW 4
template <bool QUICKER, bool FLUX>
__global__ void __launch_bounds__(32 * HFW)
k_horz_fused(const Geom g, const QuickW q, const HorzIn a, const double *__restrict__ datr, const int fold_mid,
             double *__restrict__ th, double *__restrict__ wrk1, double *__restrict__ fx, double *__restrict__ fy)
{
    const int lane = threadIdx.x, wy = threadIdx.y;
    const int i = (int)blockIdx.x * 31 + lane;             // east-face index 0..ni; lanes >= 1 own cell i
    const int j = (int)blockIdx.y * HFW + wy + 1, k = blockIdx.z + 1;
    if (j > g.nj) return;                                  // whole warp
    const bool face_ok = (i <= g.ni), cell_ok = face_ok && lane >= 1;
    const int ic = min(i, g.ni);
    const double fe = east_flux<QUICKER>(g, q, a, ic, j, k);
    const double fw = __shfl_up_sync(0xffffffffu, fe, 1);
    if (FLUX && fx && face_ok) fx[d3(g, ic, j, k)] = fe;
    if (!cell_ok) return;
    // fold_mid > 0: this rank holds the whole folded top row; columns >= fold_mid take minus the mirror column's north flux
    auto nflux = [&](int jj) -> double {
        if (fold_mid > 0 && jj == g.nj && ic >= fold_mid) return -north_flux<QUICKER>(g, q, a, g.ni + 1 - ic, jj, k);
        return north_flux<QUICKER>(g, q, a, ic, jj, k);
    };
    const double fn = nflux(j), fs = nflux(j - 1);
    const size_t c = d3(g, ic, j, k);
    if (FLUX && fy) {
        fy[c] = fn;
        if (j == 1) fy[c - g.nxd] = fs;
    }
    const double r = (a.tmask[c] * (((fe - fw) + fn) - fs)) * datr[d2(g, ic, j)];   // OTA:2645-2646 / 2282-2287
    const double w = -r;                                                             // the dispatcher negates (OTA:1936-1949)
    wrk1[c] = w;
    th[c] = th[c] + w;                                                               // OTA:1990-1996
}

// OTA:2640  mpp_update_domains(flux_x, flux_y, Dom_flux, gridtype=CGRID_NE) on the folded north edge: the NORTH-position
// component on the fold line is made antisymmetric, flux_y(i, nj) := -flux_y(ni_g+1-i, nj) for global i >= ni_g/2+1
// (MPPI/mpp_domains_define.inc:1617,2535-2549; see oracle/mom5adv_oracle.c:orc_fold_fix_flux).
// Segment descriptors are in LOCAL indices; element p of a segment is enumerated west -> east on the SOURCE side and
// lands at (dst_i1 - p) on the destination side.
struct FoldSeg {
    int src_i0, dst_i1, w;
    long long off;      // offset in the packed buffer
};
#define FOLD_MAXSEG 16
struct FoldArgs {
    int nseg;
    long long start[FOLD_MAXSEG + 1];
    FoldSeg s[FOLD_MAXSEG];
    double *fy;
    double *buf;
};
// mode 0: local (src and dst on this rank); 1: pack source row segments; 2: unpack (negated, reversed)
template <int MODE>
__global__ void k_fold_line(const Geom g, const FoldArgs a)
{
    const long long tot = a.start[a.nseg];
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
        int m = 0;
        while (e >= a.start[m + 1]) m++;
        const FoldSeg &sg = a.s[m];
        const long long r = e - a.start[m];
        const int p = (int)(r % sg.w), k = (int)(r / sg.w) + 1;
        if (MODE == 0) a.fy[d3(g, sg.dst_i1 - p, g.nj, k)] = -a.fy[d3(g, sg.src_i0 + p, g.nj, k)];
        if (MODE == 1) a.buf[sg.off + r] = a.fy[d3(g, sg.src_i0 + p, g.nj, k)];
        if (MODE == 2) a.fy[d3(g, sg.dst_i1 - p, g.nj, k)] = -a.buf[sg.off + r];
    }
}

// OTA:2642-2649 == 2282-2287, negated by the dispatcher (OTA:1936-1949), + th_tendency += wrk1 (OTA:1990-1996)
// and the zeroing of wrk1 over the data domain (OTA:1925-1931): thread (i,j,k) over the whole data domain.
__global__ void __launch_bounds__(128)
k_horz_div(const Geom g, const double *__restrict__ tmask, const double *__restrict__ datr, const double *__restrict__ fx,
           const double *__restrict__ fy, double *__restrict__ th, double *__restrict__ wrk1)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // 0..ni+1
    const int j = blockIdx.y, k = blockIdx.z + 1;          // 0..nj+1
    if (i > g.ni + 1) return;
    const size_t c = d3(g, i, j, k);
    if (i < 1 || i > g.ni || j < 1 || j > g.nj) { wrk1[c] = 0.0; return; }
    const double r = (tmask[c] * (((fx[c] - fx[c - 1]) + fy[c]) - fy[c - g.nxd])) * datr[d2(g, i, j)];
    const double w = -r;
    wrk1[c] = w;
    th[c] = th[c] + w;
}

static int horz_upwind_dev(const Geom &g, const double *tmask, const double *dyte, const double *dxtn, const double *datr,
                           const double *T, const double *u, const double *v, double *th, double *wrk1, double *fx, double *fy,
                           cudaStream_t st, int64_t *launches)
{
    if (!fx || !fy) { set_error("upwind: flux_x and flux_y work arrays are required"); return -1; }
    QuickW none{};
    dim3 gf((g.ni + 1 + 127) / 128, g.nj + 1, g.nk);
    k_horz_flux<false><<<gf, 128, 0, st>>>(g, none, tmask, nullptr, dyte, dxtn, T, nullptr, nullptr, nullptr, 0, u, v, fx, fy);
    dim3 gd((g.ni + 2 + 127) / 128, g.nj + 2, g.nk);
    k_horz_div<<<gd, 128, 0, st>>>(g, tmask, datr, fx, fy, th, wrk1);
    *launches += 2;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

static int horz_quicker_flux_dev(const Geom &g, const QuickW &q, const double *tmask, const uint8_t *mq, const double *dyte,
                                 const double *dxtn, const double *Tm1, const double *Tt, const double *tq, const double *tlimit,
                                 int limit, const double *u, const double *v, double *fx, double *fy, cudaStream_t st,
                                 int64_t *launches)
{
    // flux_x = flux_y = 0 (OTA:2568-2569)
    const size_t bytes = (size_t)g.slab * g.nk * sizeof(double);
    if (cudaMemsetAsync(fx, 0, bytes, st) != cudaSuccess || cudaMemsetAsync(fy, 0, bytes, st) != cudaSuccess) return -2;
    dim3 gf((g.ni + 1 + 127) / 128, g.nj + 1, g.nk);
    k_horz_flux<true><<<gf, 128, 0, st>>>(g, q, tmask, mq, dyte, dxtn, Tm1, Tt, tq, tlimit, limit, u, v, fx, fy);
    *launches += 1;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

static int horz_div_dev(const Geom &g, const double *tmask, const double *datr, const double *fx, const double *fy, double *th,
                        double *wrk1, cudaStream_t st, int64_t *launches)
{
    dim3 gd((g.ni + 2 + 127) / 128, g.nj + 2, g.nk);
    k_horz_div<<<gd, 128, 0, st>>>(g, tmask, datr, fx, fy, th, wrk1);
    *launches += 1;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- vertical: one thread per column marching down k (ft1 carried in a register) ----
template <bool QUICKER>
__global__ void __launch_bounds__(128)
k_vert(const Geom g, const QuickW q, const double *__restrict__ tmask, const double *__restrict__ dat,
       const double *__restrict__ Tm1, const double *__restrict__ Tt, const double *__restrict__ tlimit,
       const double *__restrict__ w, double *__restrict__ th, double *__restrict__ wrk1, double *__restrict__ fz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // 0..ni+1 (halo ring columns only zero wrk1)
    const int j = blockIdx.y;
    if (i > g.ni + 1) return;
    if (i < 1 || i > g.ni || j < 1 || j > g.nj) {          // OTA:2116-2122
        for (int k = 1; k <= g.nk; k++) wrk1[d3(g, i, j, k)] = 0.0;
        return;
    }
    const double da = dat[d2(g, i, j)];
    double ft1 = 0.0;
    for (int k = 1; k <= g.nk; k++) {
        const int km1 = max(k - 1, 1), kp1 = min(k + 1, g.nk), p2 = min(k + 2, g.nk);
        const size_t c = d3(g, i, j, k), cp1 = d3(g, i, j, kp1);
        double ft2;
        if (!QUICKER) {          // OTA:2810-2814
            const double velocity = 0.5 * w[w3(g, i, j, k)];
            const double wpos = velocity + fabs(velocity), wneg = velocity - fabs(velocity);
            ft2 = (((wneg * Tm1[c]) + (wpos * Tm1[cp1])) * tmask[c]) * tmask[cp1];
        } else {
            const double vel = w[w3(g, i, j, k)];
            const double upos = 0.5 * (vel + fabs(vel)), uneg = 0.5 * (vel - fabs(vel));
            if (tlimit[c] == 1.0) {   // OTA:3006-3008
                ft2 = (((uneg * Tm1[c]) + (upos * Tm1[cp1])) * tmask[c]) * tmask[cp1];
            } else {                  // OTA:3010-3019
                const double mp2 = tmask[d3(g, i, j, p2)];
                const int kp2 = (int)llround((mp2 * (double)p2) + ((1.0 - mp2) * (double)kp1));
                const double upmsk = tmask[c] * (1.0 - tmask[d3(g, i, j, km1)]);
                const int nk = g.nk;
                ft2 = ((vel * ((q.quick_z[k - 1] * Tt[c]) + (q.quick_z[k - 1 + nk] * Tt[cp1]))) -
                       (uneg * (((q.curv_zp[k - 1] * Tm1[cp1]) + (q.curv_zp[k - 1 + nk] * Tm1[c])) +
                                (q.curv_zp[k - 1 + 2 * nk] * Tm1[d3(g, i, j, km1)])))) -
                      (upos * (((q.curv_zn[k - 1] * Tm1[d3(g, i, j, kp2)]) + (q.curv_zn[k - 1 + nk] * Tm1[cp1])) +
                               (q.curv_zn[k - 1 + 2 * nk] * ((Tm1[c] * (1.0 - upmsk)) + (Tm1[cp1] * upmsk)))));
            }
        }
        if (fz) fz[c] = da * ft2;
        const double wv = -(tmask[c] * (ft1 - ft2));
        wrk1[c] = wv;
        th[c] = th[c] + wv;          // OTA:2162-2168
        ft1 = ft2;
    }
}

static int vert_dev(const Geom &g, const QuickW &q, const double *tmask, const double *dat, const double *Tm1, const double *Tt,
                    const double *tlimit, const double *w, double *th, double *wrk1, double *fz, int quicker, cudaStream_t st,
                    int64_t *launches)
{
    dim3 grid((g.ni + 2 + 127) / 128, g.nj + 2);
    if (quicker) k_vert<true><<<grid, 128, 0, st>>>(g, q, tmask, dat, Tm1, Tt, tlimit, w, th, wrk1, fz);
    else k_vert<false><<<grid, 128, 0, st>>>(g, q, tmask, dat, Tm1, Tt, tlimit, w, th, wrk1, fz);
    *launches += 1;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- continuity on the T grid ----
//   diverge_t = tmask*(BDX_ET(uhrho_et) + BDY_NT(vhrho_nt))                          ocean_advection_velocity.F90:660
//   BDX_ET(i,j) = (dyte(i,j)*a(i,j) - dyte(i-1,j)*a(i-1,j))*datr(i,j), 0 at i = isd   ocean_operators.F90:945-958
//   BDY_NT(i,j) = (dxtn(i,j)*a(i,j) - dxtn(i,j-1)*a(i,j-1))*datr(i,j), 0 at j = jsd   ocean_operators.F90:1230-1243
//   wrho_bt(k) = ((rho_dzt_tendency - mass_source) + diverge_t(k) + wrho_bt(k-1))*tmask                         :666-669
// one thread per data-domain column marching down k; wrho_bt(:,:,0) is the caller's.
__global__ void __launch_bounds__(128)
k_continuity(const Geom g, const double *__restrict__ tmask, const double *__restrict__ dyte, const double *__restrict__ dxtn,
             const double *__restrict__ datr, const double *__restrict__ u, const double *__restrict__ v,
             const double *__restrict__ tend, const double *__restrict__ src, double *__restrict__ w, double *__restrict__ div_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // 0..ni+1
    const int j = blockIdx.y;                              // 0..nj+1
    if (i > g.ni + 1) return;
    const size_t c2 = d2(g, i, j);
    const double dr = datr[c2], dye = dyte[c2], dxn = dxtn[c2];
    const double dyw = (i >= 1) ? dyte[c2 - 1] : 0.0, dxs = (j >= 1) ? dxtn[c2 - g.nxd] : 0.0;
    double wk = w[w3(g, i, j, 0)];
    for (int k = 1; k <= g.nk; k++) {
        const size_t c = d3(g, i, j, k);
        double bdx = 0.0, bdy = 0.0;
        if (i >= 1) bdx = ((dye * u[c]) - (dyw * u[c - 1])) * dr;
        if (j >= 1) bdy = ((dxn * v[c]) - (dxs * v[c - g.nxd])) * dr;
        const double m = tmask[c];
        const double dv = m * (bdx + bdy);
        const double tmp = (tend ? tend[c] : 0.0) - (src ? src[c] : 0.0);
        if (div_out) div_out[c] = dv;
        wk = ((tmp + dv) + wk) * m;
        w[w3(g, i, j, k)] = wk;
    }
}
