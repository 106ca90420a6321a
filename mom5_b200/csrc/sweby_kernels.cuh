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
//   x sweep: one thread per (i,j) east face, looping over k with the 2-D metrics in registers.  A warp owns
//            32 consecutive faces = 31 cells; the west-face flux and mass flux come from the neighbouring lane by
//            warp shuffle and tm(i-1..i+2) from the warp's staged row, so warps are independent (no block barrier).
//   y sweep: one thread per (i,k) marching north over a chunk of j with a rolling register window
//            (tm(j-1..j+2), R(j-1..j+1), flux(j-1)); blocks are ordered k-fastest so that the 2-D metrics and
//            w(k-1) of concurrently resident blocks hit in L2.
// All three stage the next iteration's operands global -> shared with per-thread cp.async (LDGSTS) one iteration
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
    double *flux2[NT], *dadv2[NT];   // fused x+y pass: the y sweep's diagnostics (flux/dadv are then the x sweep's)
    double *Tnew[NT];           // fused pass with the tracer time update in its epilogue (UPD): field(taup1)
    const double *rho_m1, *rho_r;    // UPD: rho_dzt(taum1), rho_dztr(taup1)
    const double *u, *v, *w, *rho;
    const uint8_t *nib;         // mask nibbles of this sweep's direction, data-domain layout
    const uint8_t *nib2;        // fused x+y pass: the y nibbles (nib holds the x nibbles)
    const unsigned *zbits;      // z sweep: per-column mask bit strings (ZBits), nzw words per column
    int nzw;
    const double *dat, *datr, *dxte, *dyte, *dxtn, *dytn;
    double dtime, sl;
    int kc;                     // z/x: levels per k-chunk;  y: rows per j-chunk
    int accumulate;             // y: th += adv
    int tile_first, tile_step;  // z, x: i-tile = tile_first + blockIdx.x*tile_step; y: j-chunk likewise (interior / edge launches)
    int row_first, row_last;    // z, x: rows row_first..row_last (whole sweep: 1..nj; the fused pass needs x on the edge rows only,
                                // the banded host-pointer pipeline runs z band by band)
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- per-thread asynchronous staging (LDGSTS): global -> shared one iteration ahead, no register cost ----
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Each sweep's arithmetic for ONE stencil level lives in a `*_level<..., EXACT>` function working on a small
// struct of per-thread operands.  The kernels run the EXACT=false flavour (shared reciprocals, branch-free); if any of
// its division guards fails (`bad`), the level is redone by the out-of-line EXACT=true flavour (plain `/`).

// =================================================================================================
// z sweep  (OTA:4150-4211 / 3843-3911)
// =================================================================================================
#ifndef ZBX
#define ZBX 128
#endif
#ifndef ZMINB
#define ZMINB 4
#endif
#ifndef XMINB
#define XMINB 4
#endif
#ifndef YMINB
#define YMINB 4
#endif

// Land/sea mask of one (i,j) column as a bit string: bit b = tmask_mdfl(i,j,clamp(b,1,nk)) for b = 0..nk+2 (k_build_zbits), so the
// z sweep's neighbourhood nibble of level k -- m(km1), m(k), m(kp1), m(kp2) with the reference's clamped indices
// (OTA:4157-4159) -- is bits k-1..k+2.  Word w of column (i,j) sits at zbits[w * slab + d2(i,j)] (coalesced across a warp);
// 75 levels are three words, loaded once, so the k loop carries no mask load.
struct ZBits {
    unsigned lo, hi, nn;    // words wi, wi+1, wi+2 of the column
    int wi, sh;
    static __device__ __forceinline__ unsigned word(const unsigned *col, unsigned stride, int nw, int w)
    {
        return (w < nw) ? col[(size_t)w * stride] : 0u;
    }
    __device__ __forceinline__ void init(const unsigned *col, unsigned stride, int nw, int k0)
    {
        wi = (k0 - 1) >> 5; sh = (k0 - 1) & 31;
        lo = word(col, stride, nw, wi); hi = word(col, stride, nw, wi + 1); nn = word(col, stride, nw, wi + 2);
    }
    // nibble of the current level: bit0 = m(km1), bit1 = m(k), bit2 = m(kp1), bit3 = m(kp2)
    __device__ __forceinline__ unsigned nib() const { return __funnelshift_r(lo, hi, sh) & 15u; }
    __device__ __forceinline__ void next(const unsigned *col, unsigned stride, int nw)   // once per level; reloads every 32 levels
    {
        if (++sh == 32) { sh = 0; wi++; lo = hi; hi = nn; nn = word(col, stride, nw, wi + 2); }
    }
};

template <int NT>
struct ZLevel {
    double dat, datr, wk, wkm1, r, dtime, sl;
    unsigned nb;
    double Tk[NT], Tp1[NT], Tp2[NT], Rm1[NT], R0[NT], ftp[NT];   // in
    double fbt[NT], Rp1[NT], t[NT], wz[NT];                        // out
};

template <int NT, int VAR, bool EXACT>
__device__ __forceinline__ unsigned z_level(ZLevel<NT> &L)
{
    unsigned bad = 0;
    const Div<EXACT> rr = Div<EXACT>::make(L.r, bad);
    const FaceCoef c = make_coef<EXACT>(L.dat * L.wk, fabs(rr.template operator()<false>(L.wk * L.dtime, bad)), nib_and(L.nb, 6u), bad);
    const double mm12 = nib_and(L.nb, 12u);           // m(kp1)*m(kp2)
    const double dtr = (VAR == VAR_ONE) ? rr(L.dtime, bad) : 0.0;
#pragma unroll
    for (int n = 0; n < NT; n++) {
        L.Rp1[n] = (L.Tp1[n] - L.Tp2[n]) * mm12;
        const double fbt = sweby_flux<VAR, EXACT>(c, L.Rm1[n], L.R0[n], L.Rp1[n], L.Tp1[n], L.Tk[n], L.sl, bad);
        L.fbt[n] = fbt;
        if (VAR == VAR_ALL) {  // OTA:4191-4195
            L.wz[n] = (L.datr * (fbt - L.ftp[n])) + (L.Tk[n] * (L.wkm1 - L.wk));
            L.t[n] = L.Tk[n] + rr(L.wz[n] * L.dtime, bad);
        } else {               // OTA:3892-3896
            L.wz[n] = 0.0;
            L.t[n] = L.Tk[n] + (dtr * ((L.datr * (fbt - L.ftp[n])) + (L.Tk[n] * (L.wkm1 - L.wk))));
        }
    }
    return bad;
}
template <int NT, int VAR>
__device__ __noinline__ void z_level_exact(ZLevel<NT> *L) { z_level<NT, VAR, true>(*L); }

#ifndef ZSTAGES
#define ZSTAGES 2    // staging depth: operands of level k + ZSTAGES - 1 are in flight while level k is computed
#endif

template <int NT, int VAR, bool DIAG>
__global__ void __launch_bounds__(ZBX, ZMINB) k_sweby_z(const Geom g, const SwebyArgs<NT> a)
{
    constexpr int NF = NT + 2;                           // T(kp2)[NT], w, rho of a level
    constexpr int D = ZSTAGES;
    __shared__ double sm[D][NF][ZBX];
    const int tx = threadIdx.x;
    const int i = (a.tile_first + (int)blockIdx.x * a.tile_step) * ZBX + tx;   // tile t = data-domain columns t*ZBX .. (capi.cu: z_tiles)
    const int j = a.row_first + (int)blockIdx.y;
    if (i < 1 || i > g.ni) return;                       // staging is per thread: no collective operation follows
    const int ks = blockIdx.z * a.kc + 1;
    const int ke = min(ks + a.kc - 1, g.nk);
    const int k0 = ks > 1 ? ks - 1 : 1;  // first face evaluated (warm-up face when the chunk starts below the surface)
    // element offsets are kept in 32 bits (every array of this path has < 2^32 elements, checked in mom5adv_init):
    // one IMAD.WIDE per address instead of a 64-bit add chain
    typedef unsigned ofs_t;
    const ofs_t c2 = (ofs_t)d2(g, i, j);
    const ofs_t slab = (ofs_t)g.slab;

    ofs_t qd = (ofs_t)d3(g, i, j, k0);                            // data-domain offset of level k
    ofs_t qt = (ofs_t)t3(g, i, j, k0);                            // h2 offset of level k
    const ofs_t qkm = (k0 > 1) ? qd - slab : qd, qkp = (k0 < g.nk) ? qd + slab : qd;

    // staging front: level kf, its data-domain offset qf and the offset q2f of level min(kf+2, nk)
    int kf = k0;
    ofs_t qf = qd, q2f = qd + slab * (ofs_t)(min(k0 + 2, g.nk) - k0);
    auto stage_next = [&](int slot) {                             // stages T(min(kf+2,nk)), w(kf), rho(kf)
        if (kf <= ke) {
#pragma unroll
            for (int n = 0; n < NT; n++) cp_async8(&sm[slot][n][tx], a.T[n] + q2f);
            cp_async8(&sm[slot][NT][tx], a.w + qf + slab);        // w3(k) = d3(k) + slab
            cp_async8(&sm[slot][NT + 1][tx], a.rho + qf);
        }
        qf += slab;
        if (kf + 3 <= g.nk) q2f += slab;
        kf++;
    };
#pragma unroll
    for (int d = 0; d < D - 1; d++) {
        stage_next(d);
        cp_async_commit();
    }

    // the column's mask bits (no per-level mask load: a plain LDG in the loop exposed one DRAM latency per level, see ZBits)
    ZBits zb;
    zb.init(a.zbits + c2, slab, a.nzw, k0);
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

    int st = 0, sf = D - 1;                                       // slot consumed / slot filled this iteration
#pragma unroll 3
    for (int k = k0; k <= ke; k++) {
        // ---- stage the operands of level k+D-1 (each thread reads back only what it staged itself) ----
        stage_next(sf);
        cp_async_commit();
        cp_async_wait<D - 1>();
#pragma unroll
        for (int n = 0; n < NT; n++) L.Tp2[n] = sm[st][n][tx];
        L.wk = sm[st][NT][tx];
        L.r = sm[st][NT + 1][tx];
        L.nb = zb.nib();
        // ---- level k ----
        if (z_level<NT, VAR, false>(L)) {
            ZLevel<NT> X = L;
            z_level_exact<NT, VAR>(&X);
#pragma unroll
            for (int n = 0; n < NT; n++) { L.fbt[n] = X.fbt[n]; L.Rp1[n] = X.Rp1[n]; L.t[n] = X.t[n]; L.wz[n] = X.wz[n]; }
        }
        const bool live = (k >= ks);
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
        zb.next(a.zbits + c2, slab, a.nzw);
        qd += slab;
        qt += (ofs_t)g.tslab;
        st = (st + 1 == D) ? 0 : st + 1;
        sf = (sf + 1 == D) ? 0 : sf + 1;
    }
}

// =================================================================================================
// x sweep  (OTA:4251-4299 / 3916-3969)
// =================================================================================================
#ifndef XWARPS
#define XWARPS 4    // warps per block, stacked along j
#endif
#define XROW 36     // staged row: tm(i_w-1 .. i_w+33) = 35 values (+1 pad)

template <int NT>
struct XFace {   // east face of (i,j,k)
    double dyte, dxte, uu, rho_i, rho_e, dtime, sl;
    unsigned nb;
    double tm1[NT], t0[NT], t1[NT], t2[NT];   // in
    double mf, f[NT];                          // out
};
template <int NT, int VAR, bool EXACT>
__device__ __forceinline__ unsigned x_face(XFace<NT> &L)
{
    unsigned bad = 0;
    L.mf = L.dyte * L.uu;
    const FaceCoef c = make_coef<EXACT>(L.mf, fabs(Div<EXACT>::make((L.rho_i + L.rho_e) * L.dxte, bad).template operator()<false>((L.uu * L.dtime) * 2.0, bad)),
                                        nib_and(L.nb, 6u), bad);
    const double mm01 = nib_and(L.nb, 3u), mm23 = nib_and(L.nb, 12u);
#pragma unroll
    for (int n = 0; n < NT; n++) {
        const double Rjp = (L.t2[n] - L.t1[n]) * mm23;      // ((tm(i+2)-tm(i+1))*m(i+2))*m(i+1)
        const double Rj = (L.t1[n] - L.t0[n]) * c.mm;       // ((tm(i+1)-tm(i))*m(i+1))*m(i)
        const double Rjm = (L.t0[n] - L.tm1[n]) * mm01;     // ((tm(i)-tm(i-1))*m(i))*m(i-1)
        L.f[n] = sweby_flux<VAR, EXACT>(c, Rjp, Rj, Rjm, L.t0[n], L.t1[n], L.sl, bad);
    }
    return bad;
}
template <int NT, int VAR>
__device__ __noinline__ void x_face_exact(XFace<NT> *L) { x_face<NT, VAR, true>(*L); }

template <int NT>
struct XCell {   // cell (i,j,k)
    double m_i, datr, rho_i, dtime, mf, mfw;
    double f[NT], fw[NT], Tc[NT], t0[NT];     // in
    double t[NT], wx[NT];                      // out
};
template <int NT, int VAR, bool EXACT>
__device__ __forceinline__ unsigned x_cell(XCell<NT> &L)
{
    unsigned bad = 0;
    const Div<EXACT> rr = Div<EXACT>::make(L.rho_i, bad);
    const double coef = (VAR == VAR_ONE) ? rr((L.dtime * L.m_i) * L.datr, bad) : 0.0;
#pragma unroll
    for (int n = 0; n < NT; n++) {
        if (VAR == VAR_ALL) {  // OTA:4288-4295
            L.wx[n] = (L.m_i * L.datr) * ((L.fw[n] - L.f[n]) + (L.Tc[n] * (L.mf - L.mfw)));
            L.t[n] = L.t0[n] + rr(L.wx[n] * L.dtime, bad);
        } else {               // OTA:3960-3965
            L.wx[n] = 0.0;
            L.t[n] = L.t0[n] + (coef * ((L.fw[n] - L.f[n]) + (L.Tc[n] * (L.mf - L.mfw))));
        }
    }
    return bad;
}
template <int NT, int VAR>
__device__ __noinline__ void x_cell_exact(XCell<NT> *L) { x_cell<NT, VAR, true>(*L); }

template <int NT, int VAR, bool DIAG>
__global__ void __launch_bounds__(32 * XWARPS, XMINB) k_sweby_x(const Geom g, const SwebyArgs<NT> a)
{
    constexpr int NF = 2 * NT + 2;                       // tm[NT], T[NT], u, rho
    __shared__ double sm[2][XWARPS][NF][XROW];
    const int lane = threadIdx.x, wy = threadIdx.y;
    const int iw = (a.tile_first + (int)blockIdx.x * a.tile_step) * 31;   // first east-face index of this warp
    const int i = iw + lane;                             // east-face index 0..ni; lanes >= 1 also update cell i
    const int j = a.row_first + blockIdx.y * XWARPS + wy;
    if (j > a.row_last) return;                          // whole warp leaves together
    const bool face_ok = (i <= g.ni);
    const bool cell_ok = face_ok && (lane >= 1);
    const int ic = min(i, g.ni);                         // clamped index for the loads of idle lanes
    const int ks = blockIdx.z * a.kc + 1;
    const int ke = min(ks + a.kc - 1, g.nk);
    typedef unsigned ofs_t;   // 32-bit element offsets (see k_sweby_z)
    const ofs_t c2 = (ofs_t)d2(g, ic, j);
    XFace<NT> F;
    XCell<NT> C;
    F.dyte = a.dyte[c2]; F.dxte = a.dxte[c2]; F.dtime = a.dtime; F.sl = a.sl;
    C.datr = a.datr[c2]; C.dtime = a.dtime;
    const ofs_t slab = (ofs_t)g.slab, tslab = (ofs_t)g.tslab;
    ofs_t q = (ofs_t)d3(g, ic, j, ks);
    // staged elements: tm element e = iw-1+lane (slot lane) and, for lanes 0..2, iw+31+lane (slot 32+lane);
    // rho element iw+lane (slot lane) and, for lane 0, iw+32 (slot 32); T, u at ic (slot lane)
    ofs_t tqa = (ofs_t)t3(g, min(iw - 1 + lane, g.ni + 2), j, ks);
    ofs_t tqb = (ofs_t)t3(g, min(iw + 31 + lane, g.ni + 2), j, ks);
    ofs_t qr = (ofs_t)d3(g, min(iw + lane, g.ni + 1), j, ks);
    ofs_t qrb = (ofs_t)d3(g, min(iw + 32, g.ni + 1), j, ks);

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
        const ofs_t q_cur = q, tq_cur = tqa + 1;        // tm(i) lives one element right of the lane's staged element
        if (k < ke) {
            q += slab; qr += slab; qrb += slab; tqa += tslab; tqb += tslab;
            stage(st ^ 1);
            nb_n = a.nib[q];
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        double(*S)[XROW] = sm[st][wy];
        F.nb = nb;
        F.uu = S[2 * NT][lane];
        F.rho_i = S[2 * NT + 1][lane];
        F.rho_e = S[2 * NT + 1][lane + 1];
#pragma unroll
        for (int n = 0; n < NT; n++) { F.tm1[n] = S[n][lane]; F.t0[n] = S[n][lane + 1]; F.t1[n] = S[n][lane + 2]; F.t2[n] = S[n][lane + 3]; }
        if (x_face<NT, VAR, false>(F)) {
            XFace<NT> X = F;
            x_face_exact<NT, VAR>(&X);
            F.mf = X.mf;
#pragma unroll
            for (int n = 0; n < NT; n++) F.f[n] = X.f[n];
        }
        C.m_i = nib_and(nb, 2u); C.rho_i = F.rho_i; C.mf = F.mf;
        C.mfw = __shfl_up_sync(0xffffffffu, F.mf, 1);
#pragma unroll
        for (int n = 0; n < NT; n++) {
            C.f[n] = F.f[n];
            C.fw[n] = __shfl_up_sync(0xffffffffu, F.f[n], 1);
            C.Tc[n] = S[NT + n][lane];
            C.t0[n] = F.t0[n];
            if (DIAG && face_ok && a.flux[n]) a.flux[n][q_cur] = F.f[n];
        }
        if (cell_ok) {
            if (x_cell<NT, VAR, false>(C)) {
                XCell<NT> X = C;
                x_cell_exact<NT, VAR>(&X);
#pragma unroll
                for (int n = 0; n < NT; n++) { C.t[n] = X.t[n]; C.wx[n] = X.wx[n]; }
            }
#pragma unroll
            for (int n = 0; n < NT; n++) {
                a.tm_out[n][tq_cur] = C.t[n];
                if (DIAG && VAR == VAR_ALL && a.dadv[n]) a.dadv[n][q_cur] = C.wx[n];
            }
        }
        nb = nb_n;
    }
}

// =================================================================================================
// y sweep + total tendency  (OTA:4362-4432 / 3980-4056)
// =================================================================================================
#ifndef YWARPS
#define YWARPS 4
#endif
#define YROW 33     // staged row: element 0 = (i_w - 1), elements 1..32 = the lanes' own i

template <int NT>
struct YLevel {   // north face of (i,jf,k) and, when live, cell (i,jf,k)
    double vv, rho0, rho1, dxtn, dytn, datr, wk, wkm1, dyte_w, u_w, dyte_c, u_c, dtime, sl;
    unsigned nb;
    int live;
    double t0[NT], t1[NT], t2[NT], Rm1[NT], R0[NT], fprev[NT], Tc[NT];   // in
    double f[NT], Rp1[NT], adv[NT], wy[NT];                                // out
};
template <int NT, int VAR, bool EXACT>
__device__ __forceinline__ unsigned y_level(YLevel<NT> &L)
{
    unsigned bad = 0;
    const double mf = L.dxtn * L.vv;
    const FaceCoef c = make_coef<EXACT>(mf, fabs(Div<EXACT>::make((L.rho0 + L.rho1) * L.dytn, bad).template operator()<false>((L.vv * L.dtime) * 2.0, bad)),
                                        nib_and(L.nb, 6u), bad);
    const double mm23 = nib_and(L.nb, 12u), m0 = nib_and(L.nb, 2u);
#pragma unroll
    for (int n = 0; n < NT; n++) {
        L.Rp1[n] = (L.t2[n] - L.t1[n]) * mm23;   // ((tm(j+2)-tm(j+1))*m(j+2))*m(j+1)
        L.f[n] = sweby_flux<VAR, EXACT>(c, L.Rp1[n], L.R0[n], L.Rm1[n], L.t0[n], L.t1[n], L.sl, bad);
    }
    if (L.live) {
        // T*( (w(k)-wkm1) + datr*(dyte(i-1)*u(i-1) - dyte(i)*u(i)) )  (OTA:4402-4406)
        const double wdiv = (L.wk - L.wkm1) + (L.datr * ((L.dyte_w * L.u_w) - (L.dyte_c * L.u_c)));
        const Div<EXACT> rr = Div<EXACT>::make(L.rho0, bad), rdt = Div<EXACT>::make(L.dtime, bad);
        const double c1 = (VAR == VAR_ONE) ? rr((L.dtime * m0) * L.datr, bad) : 0.0;
#pragma unroll
        for (int n = 0; n < NT; n++) {
            double t;
            if (VAR == VAR_ALL) {  // OTA:4401-4413
                L.wy[n] = ((m0 * L.datr) * (L.fprev[n] - L.f[n])) + (L.Tc[n] * wdiv);
                t = L.t0[n] + rr(L.wy[n] * L.dtime, bad);
            } else {               // OTA:4025-4040
                L.wy[n] = 0.0;
                t = L.t0[n] + (c1 * (L.fprev[n] - L.f[n]));
                t = t + (rr(L.dtime * L.Tc[n], bad) * wdiv);
            }
            L.adv[n] = rdt(L.rho0 * (t - L.Tc[n]), bad) * m0;
        }
    }
    return bad;
}
template <int NT, int VAR>
__device__ __noinline__ void y_level_exact(YLevel<NT> *L) { y_level<NT, VAR, true>(*L); }

template <int NT, int VAR, bool DIAG>
__global__ void __launch_bounds__(32 * YWARPS, YMINB) k_sweby_y(const Geom g, const SwebyArgs<NT> a, const int nxt)
{
    constexpr int NF = 3 * NT + 9;   // tm, T, th [NT each]; v, rho, u, wk, wkm1, dxtn, dytn, datr, dyte
    constexpr int F_T = NT, F_TH = 2 * NT, F_V = 3 * NT, F_RHO = F_V + 1, F_U = F_V + 2, F_WK = F_V + 3, F_WM = F_V + 4,
                  F_DXTN = F_V + 5, F_DYTN = F_V + 6, F_DATR = F_V + 7, F_DYTE = F_V + 8;
    __shared__ double sm[2][YWARPS][NF][YROW];
    // linear block id, k fastest
    const int lin = blockIdx.x;
    const int k = lin % g.nk + 1;
    const int rest = lin / g.nk;
    const int xt = rest % nxt, jc = a.tile_first + (rest / nxt) * a.tile_step;
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const int i_raw = xt * (32 * YWARPS) + threadIdx.x + 1;
    if (i_raw - lane > g.ni) return;                     // the whole warp is outside: leave together
    const bool ok = (i_raw <= g.ni);
    const int i = min(i_raw, g.ni);
    const int js = jc * a.kc + 1;
    const int je = min(js + a.kc - 1, g.nj);
    typedef unsigned ofs_t;   // 32-bit element offsets (see k_sweby_z)
    const ofs_t nxd = (ofs_t)g.nxd, tp = (ofs_t)g.tpitch;
    const ofs_t wofs = (ofs_t)g.slab;    // w3(k) = d3(k) + slab ; w3(k-1) = d3(k)
    const bool has_km1 = (k > 1);

    ofs_t q = (ofs_t)d3(g, i, js - 1, k);        // data-domain offset of (i, jf, k)
    ofs_t c2 = (ofs_t)d2(g, i, js - 1);
    ofs_t tq = (ofs_t)t3(g, i, js - 1, k);

    auto stage = [&](int st, ofs_t qq, ofs_t cc, ofs_t tt) {   // operands of the iteration at face row jf (offsets of that row)
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
    YLevel<NT> L;
    L.dtime = a.dtime; L.sl = a.sl;
    L.nb = a.nib[q];
#pragma unroll
    for (int n = 0; n < NT; n++) {
        const double tmm = a.tm_in[n][tq - tp];
        L.t0[n] = a.tm_in[n][tq];
        L.t1[n] = a.tm_in[n][tq + tp];
        L.Rm1[n] = (L.t0[n] - tmm) * nib_and(L.nb, 3u);     // ((tm(j)-tm(j-1))*m(j))*m(j-1)
        L.R0[n] = (L.t1[n] - L.t0[n]) * nib_and(L.nb, 6u);  // ((tm(j+1)-tm(j))*m(j+1))*m(j)
        L.fprev[n] = 0.0;
    }
    L.rho0 = a.rho[q];

    int st = 0;
#pragma unroll 3
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
        L.live = (jf >= js);
        L.vv = S[F_V][lane + 1];
        L.rho1 = S[F_RHO][lane + 1];
        L.dxtn = S[F_DXTN][lane + 1];
        L.dytn = S[F_DYTN][lane + 1];
        L.datr = S[F_DATR][lane + 1];
        L.wk = S[F_WK][lane + 1];
        L.wkm1 = has_km1 ? S[F_WM][lane + 1] : 0.0;
        L.dyte_w = S[F_DYTE][lane]; L.u_w = S[F_U][lane];
        L.dyte_c = S[F_DYTE][lane + 1]; L.u_c = S[F_U][lane + 1];
#pragma unroll
        for (int n = 0; n < NT; n++) { L.t2[n] = S[n][lane + 1]; L.Tc[n] = S[F_T + n][lane + 1]; }
        if (y_level<NT, VAR, false>(L)) {
            YLevel<NT> X = L;
            y_level_exact<NT, VAR>(&X);
#pragma unroll
            for (int n = 0; n < NT; n++) { L.f[n] = X.f[n]; L.Rp1[n] = X.Rp1[n]; L.adv[n] = X.adv[n]; L.wy[n] = X.wy[n]; }
        }
#pragma unroll
        for (int n = 0; n < NT; n++) {
            if (DIAG && ok && a.flux[n]) a.flux[n][q] = L.f[n];
            if (L.live && ok) {
                a.adv[n][q] = L.adv[n];
                if (a.accumulate) a.th[n][q] = S[F_TH + n][lane + 1] + L.adv[n];
                if (DIAG && VAR == VAR_ALL && a.dadv[n]) a.dadv[n][q] = L.wy[n];
            }
            L.fprev[n] = L.f[n];
            L.Rm1[n] = L.R0[n];
            L.R0[n] = L.Rp1[n];
            L.t0[n] = L.t1[n];
            L.t1[n] = L.t2[n];
        }
        L.rho0 = L.rho1;
        L.nb = nb_n;
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

// per-column mask bit strings for the z sweep (ZBits): bit b = m(i,j,clamp(b,1,nk)), b = 0..nk+2; data-domain 2-D layout per word
__global__ void k_build_zbits(const Geom g, const uint8_t *__restrict__ m, unsigned *__restrict__ zbits, const int nzw)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // 0..ni+1
    const int j = blockIdx.y;                              // 0..nj+1
    if (i > g.ni + 1) return;
    for (int w = 0; w < nzw; w++) {
        unsigned v = 0;
        for (int q = 0; q < 32; q++) {
            const int b = 32 * w + q;
            if (b > g.nk + 2) break;
            const int k = min(max(b, 1), g.nk);
            v |= (m[m3(g, i, j, k)] ? 1u : 0u) << q;
        }
        zbits[(size_t)w * g.slab + d2(g, i, j)] = v;
    }
}

// mask nibbles from the halo-2 u8 mask for the x and y sweeps: m(-1), m(0), m(+1), m(+2) along the sweep direction in bits 0..3.
// Two layouts: data-domain (nx, ny: what the per-thread kernels index with their own cell offset) and, for the TMA-staged fused
// pass, rows padded to a pitch of `np` bytes (a multiple of 16: tensor-map strides must be) -- nxp, nyp, (np, nyd, nk).
// (The z sweep reads the column bit strings of k_build_zbits instead.)
__global__ void k_build_nibbles(const Geom g, const uint8_t *__restrict__ m, uint8_t *__restrict__ nx, uint8_t *__restrict__ ny,
                                uint8_t *__restrict__ nxp, uint8_t *__restrict__ nyp, const int np)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // 0..ni+1
    const int j = blockIdx.y, k = blockIdx.z + 1;          // 0..nj+1
    if (i > g.ni + 1) return;
    const size_t q = d3(g, i, j, k), qp = (size_t)i + (size_t)np * ((size_t)j + (size_t)g.nyd * (size_t)(k - 1));
    auto M = [&](int ii, int jj, int kk) -> unsigned { return m[m3(g, ii, jj, kk)] ? 1u : 0u; };
    const uint8_t vx = (i <= g.ni) ? (uint8_t)(M(i - 1, j, k) | (M(i, j, k) << 1) | (M(i + 1, j, k) << 2) | (M(i + 2, j, k) << 3)) : 0;
    const uint8_t vy = (j <= g.nj) ? (uint8_t)(M(i, j - 1, k) | (M(i, j, k) << 1) | (M(i, j + 1, k) << 2) | (M(i, j + 2, k) << 3)) : 0;
    nx[q] = vx; ny[q] = vy;
    nxp[qp] = vx; nyp[qp] = vy;
}
