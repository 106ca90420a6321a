// halo.cuh -- halo update of halo-2 ("h2") scratch fields: the GPU replacement, for this path only, of
// FMS mpp_update_domains / mpp_start|complete_update_domains (OTA:4213-4240, 4302-4343, 3912, 3970, 2566).
//
// Semantics restated from src/shared/mpp/include/mpp_do_update.h:57-78,186-293 and
// mpp_domains_define.inc:4865-4885 (scalar at CENTER position):
//   XUPDATE = E/W strips for j in the compute domain; YUPDATE = N/S strips for i in the compute domain;
//   both = all eight directions.  Interior neighbour and cyclic wrap are plain copies; the folded north edge
//   maps (i, nj+m) <- (ni+1-i, nj+1-m); a solid wall has no image and its halo is left untouched.
// Every halo point's source is a COMPUTE-domain point of exactly one rank, so one round of messages
// suffices and there is no ordering requirement between directions.
//
// Strips whose source is this rank (single-rank cyclic wrap, fold onto the same rank) are served by one
// local copy kernel; the rest are packed per peer, exchanged with ncclSend/ncclRecv inside one group, and
// unpacked.  All tracers of a call travel in one message per peer (the reference's `complete=` aggregation).
#pragma once

#include "mom5adv_internal.cuh"

#define HALO_MAXMSG 24
#define HALO_MAXF 16

struct CopyDesc {
    int si0, sj0, si1, sj1;   // source rectangle (local h2 indices) -- used by pack and local copy
    int di0, dj0;             // destination rectangle origin        -- used by unpack and local copy
    int w, h, flip;
    long long off;            // element offset of this strip inside the packed buffer (per field-level block layout below)
};

struct CopyArgs {
    int nmsg, nf, nk;
    long long total;              // total elements = sum_m w*h*nk*nf
    long long start[HALO_MAXMSG + 1];  // prefix of per-message element counts
    CopyDesc d[HALO_MAXMSG];
    double *f[HALO_MAXF];
    double *buf;
};

// mode 0: local copy (src rect -> dst rect, same rank); 1: pack (src rect -> buf); 2: unpack (buf -> dst rect)
// LAYOUT 0: h2 scratch fields (halo 2, t3 indexing); LAYOUT 1: caller's data-domain arrays (halo 1, d3 indexing);
// LAYOUT 2: h4 scratch fields (halo 4, q3 indexing; MDPPM)
template <int LAYOUT>
__device__ __forceinline__ size_t halo_idx(const Geom &g, int i, int j, int k)
{
    return LAYOUT == 0 ? t3(g, i, j, k) : LAYOUT == 1 ? d3(g, i, j, k) : q3(g, i, j, k);
}

template <int MODE, int LAYOUT>
__global__ void k_halo(const Geom g, const CopyArgs a)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < a.total; e += (long long)gridDim.x * blockDim.x) {
        int m = 0;
        while (e >= a.start[m + 1]) m++;
        const CopyDesc &d = a.d[m];
        long long r = e - a.start[m];
        const int p = (int)(r % d.w); r /= d.w;
        const int q = (int)(r % d.h); r /= d.h;
        const int k = (int)(r % a.nk) + 1;
        const int n = (int)(r / a.nk);
        // element (p,q) is enumerated in the SENDER's orientation
        const int si = d.si0 + p, sj = d.sj0 + q;
        const int di = d.flip ? d.di0 + (d.w - 1 - p) : d.di0 + p;
        const int dj = d.flip ? d.dj0 + (d.h - 1 - q) : d.dj0 + q;
        if (MODE == 0) a.f[n][halo_idx<LAYOUT>(g, di, dj, k)] = a.f[n][halo_idx<LAYOUT>(g, si, sj, k)];
        if (MODE == 1) a.buf[d.off + (e - a.start[m])] = a.f[n][halo_idx<LAYOUT>(g, si, sj, k)];
        if (MODE == 2) a.f[n][halo_idx<LAYOUT>(g, di, dj, k)] = a.buf[d.off + (e - a.start[m])];
    }
}

// field(taup1) = (rho_dzt(taum1)*field(taum1) + dtime*th_tendency)*rho_dztr(taup1)   (ocean_tracer.F90:2341-2350)
template <int NT>
struct UpdArgs {
    const double *T[NT], *th[NT];
    double *Tnew[NT];
    const double *rho_m1, *rho_r;
    double dtime;
};
template <int NT>
__global__ void __launch_bounds__(128) k_tracer_update(const Geom g, const UpdArgs<NT> a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > g.ni) return;
    const size_t q = d3(g, i, j, k);
    const double r0 = a.rho_m1[q], rr = a.rho_r[q];
#pragma unroll
    for (int n = 0; n < NT; n++)
        if (a.Tnew[n]) a.Tnew[n][q] = ((r0 * a.T[n][q]) + (a.dtime * a.th[n][q])) * rr;
}

// compute_adv_diss (OTA:7547-7712), element-wise parts: wrk1 = field(tau)**2 over the data domain (OTA:7574-7580) ...
__global__ void k_square(const size_t n, const double *__restrict__ T, double *__restrict__ out)
{
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) out[q] = T[q] * T[q];
}
// ... and wrk1 = wrk2 + wrk3, wrk4 = -(conversion**2)*dtimer*(term1 + term2) on the compute domain, 0 elsewhere (OTA:7680-7701)
struct DissArgs {
    const double *rho_tau, *rho_taup1, *T_tau, *adv, *wrk2, *wrk3;
    double *t2, *diss;
    double dtime, dtimer, conversion;
};
__global__ void __launch_bounds__(128) k_adv_diss(const Geom g, const DissArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;        // 0..ni+1
    const int j = blockIdx.y, k = blockIdx.z + 1;                // 0..nj+1
    if (i > g.ni + 1) return;
    const size_t q = d3(g, i, j, k);
    double w1 = 0.0, d = 0.0;
    if (i >= 1 && i <= g.ni && j >= 1 && j <= g.nj) {
        w1 = a.wrk2[q] + a.wrk3[q];
        const double at = a.adv[q];
        const double term1 = at * (((2.0 * a.rho_tau[q]) * a.T_tau[q]) + (a.dtime * at));
        const double term2 = -(a.rho_taup1[q] * w1);
        d = (-(a.conversion * a.conversion) * a.dtimer) * (term1 + term2);
    }
    if (a.t2) a.t2[q] = w1;
    a.diss[q] = d;
}

// z-integrated flux diagnostics (OTA:4317-4326, 4449-4458): out(i,j) = sum_k flux(i,j,k) in k order, compute domain, 0 elsewhere
__global__ void __launch_bounds__(128) k_flux_int_z(const Geom g, const double *__restrict__ flux, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i > g.ni + 1) return;
    double acc = 0.0;
    if (i >= 1 && i <= g.ni && j >= 1 && j <= g.nj)
        for (int k = 1; k <= g.nk; k++) acc = acc + flux[d3(g, i, j, k)];
    out[d2(g, i, j)] = acc;
}

// compute-domain copy between a data-domain array and an h2 field (tracer_quick / tmask staging)
__global__ void k_d1_to_h2(const Geom g, const double *__restrict__ src, double *__restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i <= g.ni) dst[t3(g, i, j, k)] = src[d3(g, i, j, k)];
}

__global__ void k_h2_to_mask(const Geom g, const double *__restrict__ src, uint8_t *__restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x - 1;   // -1..ni+2
    const int j = (int)blockIdx.y - 1, k = blockIdx.z + 1;
    if (i <= g.ni + 2) dst[m3(g, i, j, k)] = (src[t3(g, i, j, k)] != 0.0) ? 1 : 0;
}
