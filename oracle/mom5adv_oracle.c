/*
 * mom5adv_oracle.c -- CPU ORACLE (test infrastructure; see mom5adv_oracle.h header comment).
 *
 * Every loop nest below cites the reference lines it restates.
 *   OTA  = /root/reference/src/mom5/ocean_tracers/ocean_tracer_advect.F90
 *   MPPI = /root/reference/src/shared/mpp/include
 * Expression trees are written with explicit parentheses that reproduce Fortran's
 * left-to-right evaluation of equal-precedence operators.  Compile WITHOUT FMA contraction.
 */
#include "../include/mpp_layout.h"
#include "mom5adv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ONESIXTH (1.0 / 6.0) /* ocean_parameters.F90:92 */

/* max/min: first argument wins ties (and is kept when the second is NaN); inputs are NaN-free. */
static inline double orc_max(double a, double b) { return (b > a) ? b : a; }
static inline double orc_min(double a, double b) { return (b < a) ? b : a; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* index helpers (Fortran indices: i in isd..ied etc.; k 1-based) */
#define NX1(b) ((b)->ni + 2)
#define NY1(b) ((b)->nj + 2)
#define NX2(b) ((b)->ni + 4)
#define NY2(b) ((b)->nj + 4)
#define D2(b, i, j) ((size_t)(i) + (size_t)NX1(b) * (size_t)(j))
#define D3(b, i, j, k) ((size_t)(i) + (size_t)NX1(b) * ((size_t)(j) + (size_t)NY1(b) * (size_t)((k)-1)))
#define W3(b, i, j, k) ((size_t)(i) + (size_t)NX1(b) * ((size_t)(j) + (size_t)NY1(b) * (size_t)(k))) /* k = 0..nk */
#define H2(b, i, j) ((size_t)((i) + 1) + (size_t)NX2(b) * (size_t)((j) + 1))
#define H3(b, i, j, k) ((size_t)((i) + 1) + (size_t)NX2(b) * ((size_t)((j) + 1) + (size_t)NY2(b) * (size_t)((k)-1)))

/* ------------------------------------------------------------------------------------------------
 * Layout: mpp_define_layout2D (MPPI/mpp_domains_define.inc:28-55)
 * ---------------------------------------------------------------------------------------------- */
void orc_define_layout(int ni_g, int nj_g, int ndivs, int *layout2) { mpp_define_layout2d_c(ni_g, nj_g, ndivs, layout2); }

/* mpp_compute_extent (MPPI/mpp_domains_define.inc:187-273), no user extent. Returns 0 on success.  The rule itself lives in
 * include/mpp_layout.h, shared with the product library: it is layout bookkeeping, not arithmetic under test. */
int orc_compute_extent(int isg, int ieg, int ndivs, int *ibegin, int *iend) { return mpp_compute_extent_c(isg, ieg, ndivs, ibegin, iend); }

/* ------------------------------------------------------------------------------------------------
 * Halo update among in-process blocks.
 * Semantics (scalar field at CENTER position):
 *   - interior neighbour / cyclic wrap: plain copy                         (MPPI/mpp_do_update.h:186-293)
 *   - folded north edge: (i, nj_g+m) <- (ni_g+1-i, nj_g+1-m), no sign      (MPPI/mpp_domains_define.inc:4865-4885,
 *                                                                            test_mpp_domains.F90:3749-3766)
 *   - solid wall: halo left untouched                                      (no overlap is generated)
 *   - XUPDATE: E/W strips, j in compute domain only; YUPDATE: N/S strips, i in compute only;
 *     both: all eight directions                                           (MPPI/mpp_do_update.h:57-78)
 * ---------------------------------------------------------------------------------------------- */
static int find_div(const int *beg, const int *end, int n, int g)
{
    for (int d = 0; d < n; d++)
        if (g >= beg[d] && g <= end[d]) return d;
    return -1;
}

/* map a global (possibly out-of-range) point to its source global compute point; 0 if none (wall) */
static int map_source(const orc_layout *L, int ig, int jg, int *igs, int *jgs)
{
    if (jg > L->nj_g) {
        if (L->fold_north) {
            jg = 2 * L->nj_g + 1 - jg;
            ig = L->ni_g + 1 - ig;
        } else if (L->cyclic_y) {
            jg -= L->nj_g;
        } else
            return 0;
    } else if (jg < 1) {
        if (L->cyclic_y)
            jg += L->nj_g;
        else
            return 0;
    }
    if (ig < 1) {
        if (!L->cyclic_x) return 0;
        ig += L->ni_g;
    } else if (ig > L->ni_g) {
        if (!L->cyclic_x) return 0;
        ig -= L->ni_g;
    }
    if (ig < 1 || ig > L->ni_g || jg < 1 || jg > L->nj_g) return 0;
    *igs = ig;
    *jgs = jg;
    return 1;
}

void orc_update_halo(const orc_layout *L, double *const *fields, int nk, int halo, int flags)
{
    int nb = L->px * L->py;
    /* Two-phase (gather into per-block staging, then write) so that a block's halo never feeds another
     * block's halo within one update -- sources are compute-domain points only, so single phase is safe:
     * halo cells are written, compute cells are read, the two sets are disjoint. */
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int b = 0; b < nb; b++) {
        int bx = b % L->px, by = b / L->px;
        int ni = L->iend[bx] - L->ibeg[bx] + 1, nj = L->jend[by] - L->jbeg[by] + 1;
        int nxh = ni + 2 * halo, nyh = nj + 2 * halo;
        double *dst = fields[b];
        for (int jl = 1 - halo; jl <= nj + halo; jl++) {
            int j_in = (jl >= 1 && jl <= nj);
            for (int il = 1 - halo; il <= ni + halo; il++) {
                int i_in = (il >= 1 && il <= ni);
                if (i_in && j_in) { il = ni; continue; } /* skip the compute domain */
                int want;
                if (!i_in && j_in)
                    want = flags & ORC_XUPDATE;
                else if (i_in && !j_in)
                    want = flags & ORC_YUPDATE;
                else
                    want = (flags & ORC_XUPDATE) && (flags & ORC_YUPDATE);
                if (!want) continue;
                int igs, jgs;
                if (!map_source(L, L->ibeg[bx] + il - 1, L->jbeg[by] + jl - 1, &igs, &jgs)) continue;
                int sx = find_div(L->ibeg, L->iend, L->px, igs), sy = find_div(L->jbeg, L->jend, L->py, jgs);
                int sni = L->iend[sx] - L->ibeg[sx] + 1, snj = L->jend[sy] - L->jbeg[sy] + 1;
                int snxh = sni + 2 * halo, snyh = snj + 2 * halo;
                const double *src = fields[sx + L->px * sy];
                int isl = igs - L->ibeg[sx] + 1, jsl = jgs - L->jbeg[sy] + 1;
                size_t so = (size_t)(isl - 1 + halo) + (size_t)snxh * (size_t)(jsl - 1 + halo);
                size_t dofs = (size_t)(il - 1 + halo) + (size_t)nxh * (size_t)(jl - 1 + halo);
                for (int k = 0; k < nk; k++)
                    dst[dofs + (size_t)nxh * nyh * k] = src[so + (size_t)snxh * snyh * k];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * mdfl_init: tmask_mdfl = 0; compute domain := Grd%tmask (OTA:1660-1674); halo update by caller (OTA:1675)
 * ---------------------------------------------------------------------------------------------- */
void orc_mdfl_init_mask(const orc_block *b)
{
    memset(b->tmask_h2, 0, sizeof(double) * (size_t)NX2(b) * NY2(b) * b->nk);
    for (int k = 1; k <= b->nk; k++)
        for (int j = 1; j <= b->nj; j++)
            for (int i = 1; i <= b->ni; i++) b->tmask_h2[H3(b, i, j, k)] = b->tmask[D3(b, i, j, k)];
}

/* shared Sweby limiter block (OTA:4174-4189 == 4266-4281 == 4377-4392), see SURVEY.md Appendix A.1 */
static inline double sweby_flux(double Rjp, double Rj, double Rjm, double massflux, double cfl, double Tup,
                                double Tdn, double mA, double mB)
{
    double d0 = ((2.0 - cfl) * (1.0 - cfl)) * ONESIXTH;
    double d1 = (1.0 - (cfl * cfl)) * ONESIXTH;
    double thetaP = Rjm / (1.0e-30 + Rj);
    double thetaM = Rjp / (1.0e-30 + Rj);
    double psiP = orc_max(0.0, orc_min(orc_min(1.0, d0 + (d1 * thetaP)), ((1.0 - cfl) / (1.0e-30 + cfl)) * thetaP));
    double psiM = orc_max(0.0, orc_min(orc_min(1.0, d0 + (d1 * thetaM)), ((1.0 - cfl) / (1.0e-30 + cfl)) * thetaM));
    return ((0.5 * (((massflux + fabs(massflux)) * (Tup + (psiP * Rj))) +
                    ((massflux - fabs(massflux)) * (Tdn - (psiM * Rj))))) *
            mA) *
           mB;
}

/* a2 variant: psi = psi_lin*(1-sl) + psi_lim*sl (OTA:3874-3884) */
static inline double sweby_flux_sl(double Rjp, double Rj, double Rjm, double massflux, double cfl, double Tup,
                                   double Tdn, double mA, double mB, double sl)
{
    double d0 = ((2.0 - cfl) * (1.0 - cfl)) * ONESIXTH;
    double d1 = (1.0 - (cfl * cfl)) * ONESIXTH;
    double thetaP = Rjm / (1.0e-30 + Rj);
    double thetaM = Rjp / (1.0e-30 + Rj);
    double psiP = d0 + (d1 * thetaP);
    psiP = (psiP * (1.0 - sl)) +
           (orc_max(0.0, orc_min(orc_min(1.0, d0 + (d1 * thetaP)), ((1.0 - cfl) / (1.0e-30 + cfl)) * thetaP)) * sl);
    double psiM = d0 + (d1 * thetaM);
    psiM = (psiM * (1.0 - sl)) +
           (orc_max(0.0, orc_min(orc_min(1.0, d0 + (d1 * thetaM)), ((1.0 - cfl) / (1.0e-30 + cfl)) * thetaM)) * sl);
    return ((0.5 * (((massflux + fabs(massflux)) * (Tup + (psiP * Rj))) +
                    ((massflux - fabs(massflux)) * (Tdn - (psiM * Rj))))) *
            mA) *
           mB;
}

/* ------------------------------------------------------------------------------------------------
 * advect_tracer_sweby_all, z sweep (OTA:4136-4211)
 * ---------------------------------------------------------------------------------------------- */
void orc_sweby_all_z(const orc_block *b, int ntr, double dtime, const double *const *T, const double *w,
                     const double *rho, double *const *tm, double *const *flux_z, double *const *adv_z)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *m = b->tmask_h2;
    double *ftp = (double *)malloc(sizeof(double) * (size_t)ni * nj);
    double *wkm1 = (double *)malloc(sizeof(double) * (size_t)ni * nj);
    for (int n = 0; n < ntr; n++) {
        const double *Tn = T[n];
        double *tmn = tm[n];
        for (size_t q = 0; q < (size_t)ni * nj; q++) { ftp[q] = 0.0; wkm1[q] = 0.0; } /* OTA:4138-4140 */
        for (int k = 1; k <= nk; k++) {
            int kp1 = imin(k + 1, nk), kp2 = imin(k + 2, nk), km1 = imax(k - 1, 1);
            for (int j = 1; j <= nj; j++)
                for (int i = 1; i <= ni; i++) {
                    size_t c = (size_t)(i - 1) + (size_t)ni * (j - 1);
                    double Tk = Tn[D3(b, i, j, k)], Tkp1 = Tn[D3(b, i, j, kp1)];
                    double Rjp = ((Tn[D3(b, i, j, km1)] - Tk) * m[H3(b, i, j, km1)]) * m[H3(b, i, j, k)];
                    double Rj = ((Tk - Tkp1) * m[H3(b, i, j, k)]) * m[H3(b, i, j, kp1)];
                    double Rjm = ((Tkp1 - Tn[D3(b, i, j, kp2)]) * m[H3(b, i, j, kp1)]) * m[H3(b, i, j, kp2)];
                    double wk = w[W3(b, i, j, k)], r = rho[D3(b, i, j, k)];
                    double massflux = b->dat[D2(b, i, j)] * wk;
                    double cfl = fabs((wk * dtime) / r);
                    double fbt = sweby_flux(Rjp, Rj, Rjm, massflux, cfl, Tkp1, Tk, m[H3(b, i, j, kp1)], m[H3(b, i, j, k)]);
                    double wz = (b->datr[D2(b, i, j)] * (fbt - ftp[c])) + (Tk * (wkm1[c] - wk));  /* OTA:4191-4192 */
                    tmn[H3(b, i, j, k)] = Tk + ((wz * dtime) / r);                               /* OTA:4154,4194-4195 */
                    if (flux_z && flux_z[n]) flux_z[n][D3(b, i, j, k)] = fbt;
                    if (adv_z && adv_z[n]) adv_z[n][D3(b, i, j, k)] = wz;
                    ftp[c] = fbt;
                }
            for (int j = 1; j <= nj; j++) /* OTA:4205-4209 */
                for (int i = 1; i <= ni; i++) wkm1[(size_t)(i - 1) + (size_t)ni * (j - 1)] = w[W3(b, i, j, k)];
        }
    }
    free(ftp);
    free(wkm1);
}

/* x sweep (OTA:4243-4300) */
void orc_sweby_all_x(const orc_block *b, int ntr, double dtime, const double *const *T, const double *u,
                     const double *rho, double *const *tm, double *const *flux_x, double *const *adv_x)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *m = b->tmask_h2;
    double *fx = (double *)malloc(sizeof(double) * (size_t)(ni + 1)); /* faces i = 0..ni of one row */
    for (int n = 0; n < ntr; n++) {
        const double *Tn = T[n];
        double *tmn = tm[n];
        for (int k = 1; k <= nk; k++)
            for (int j = 1; j <= nj; j++) {
                for (int i = 0; i <= ni; i++) { /* OTA:4254-4282 */
                    double t0 = tmn[H3(b, i, j, k)], t1 = tmn[H3(b, i + 1, j, k)];
                    double Rjp = ((tmn[H3(b, i + 2, j, k)] - t1) * m[H3(b, i + 2, j, k)]) * m[H3(b, i + 1, j, k)];
                    double Rj = ((t1 - t0) * m[H3(b, i + 1, j, k)]) * m[H3(b, i, j, k)];
                    double Rjm = ((t0 - tmn[H3(b, i - 1, j, k)]) * m[H3(b, i, j, k)]) * m[H3(b, i - 1, j, k)];
                    double uu = u[D3(b, i, j, k)];
                    double massflux = b->dyte[D2(b, i, j)] * uu;
                    double cfl = fabs(((uu * dtime) * 2.0) /
                                      ((rho[D3(b, i, j, k)] + rho[D3(b, i + 1, j, k)]) * b->dxte[D2(b, i, j)]));
                    fx[i] = sweby_flux(Rjp, Rj, Rjm, massflux, cfl, t0, t1, m[H3(b, i, j, k)], m[H3(b, i + 1, j, k)]);
                    if (flux_x && flux_x[n]) flux_x[n][D3(b, i, j, k)] = fx[i];
                }
                for (int i = 1; i <= ni; i++) { /* OTA:4286-4296 */
                    double wx = (m[H3(b, i, j, k)] * b->datr[D2(b, i, j)]) *
                                ((fx[i - 1] - fx[i]) +
                                 (Tn[D3(b, i, j, k)] * ((b->dyte[D2(b, i, j)] * u[D3(b, i, j, k)]) -
                                                       (b->dyte[D2(b, i - 1, j)] * u[D3(b, i - 1, j, k)]))));
                    tmn[H3(b, i, j, k)] = tmn[H3(b, i, j, k)] + ((wx * dtime) / rho[D3(b, i, j, k)]);
                    if (adv_x && adv_x[n]) adv_x[n][D3(b, i, j, k)] = wx;
                }
            }
    }
    free(fx);
}

/* y sweep + total tendency (OTA:4347-4432) */
void orc_sweby_all_y(const orc_block *b, int ntr, double dtime, const double *const *T, const double *u,
                     const double *v, const double *w, const double *rho, double *const *tm,
                     double *const *th, double *const *adv, double *const *flux_y, double *const *adv_y)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *m = b->tmask_h2;
    double *fy = (double *)malloc(sizeof(double) * (size_t)ni * (nj + 1)); /* faces j = 0..nj */
    for (int n = 0; n < ntr; n++) {
        const double *Tn = T[n];
        double *tmn = tm[n];
        /* T_prog(n)%wrk1 zeroed over the whole data domain (OTA:4142-4148) */
        memset(adv[n], 0, sizeof(double) * (size_t)NX1(b) * NY1(b) * nk);
        for (int k = 1; k <= nk; k++) {
            for (int j = 0; j <= nj; j++) /* OTA:4364-4395 */
                for (int i = 1; i <= ni; i++) {
                    double t0 = tmn[H3(b, i, j, k)], t1 = tmn[H3(b, i, j + 1, k)];
                    double Rjp = ((tmn[H3(b, i, j + 2, k)] - t1) * m[H3(b, i, j + 2, k)]) * m[H3(b, i, j + 1, k)];
                    double Rj = ((t1 - t0) * m[H3(b, i, j + 1, k)]) * m[H3(b, i, j, k)];
                    double Rjm = ((t0 - tmn[H3(b, i, j - 1, k)]) * m[H3(b, i, j, k)]) * m[H3(b, i, j - 1, k)];
                    double vv = v[D3(b, i, j, k)];
                    double massflux = b->dxtn[D2(b, i, j)] * vv;
                    double cfl = fabs(((vv * dtime) * 2.0) /
                                      ((rho[D3(b, i, j, k)] + rho[D3(b, i, j + 1, k)]) * b->dytn[D2(b, i, j)]));
                    double f = sweby_flux(Rjp, Rj, Rjm, massflux, cfl, t0, t1, m[H3(b, i, j, k)], m[H3(b, i, j + 1, k)]);
                    fy[(size_t)(i - 1) + (size_t)ni * j] = f;
                    if (flux_y && flux_y[n]) flux_y[n][D3(b, i, j, k)] = f;
                }
            for (int j = 1; j <= nj; j++) /* OTA:4398-4424 */
                for (int i = 1; i <= ni; i++) {
                    double Tk = Tn[D3(b, i, j, k)], r = rho[D3(b, i, j, k)];
                    double wkm1 = (k == 1) ? 0.0 : w[W3(b, i, j, k - 1)]; /* OTA:4356-4360, 4427-4431 */
                    double wy = ((m[H3(b, i, j, k)] * b->datr[D2(b, i, j)]) *
                                 (fy[(size_t)(i - 1) + (size_t)ni * (j - 1)] - fy[(size_t)(i - 1) + (size_t)ni * j])) +
                                (Tk * ((w[W3(b, i, j, k)] - wkm1) +
                                       (b->datr[D2(b, i, j)] * ((b->dyte[D2(b, i - 1, j)] * u[D3(b, i - 1, j, k)]) -
                                                                (b->dyte[D2(b, i, j)] * u[D3(b, i, j, k)])))));
                    double t = tmn[H3(b, i, j, k)] + ((wy * dtime) / r);
                    tmn[H3(b, i, j, k)] = t;
                    double a = ((r * (t - Tk)) / dtime) * m[H3(b, i, j, k)];
                    adv[n][D3(b, i, j, k)] = a;
                    th[n][D3(b, i, j, k)] = th[n][D3(b, i, j, k)] + a;
                    if (adv_y && adv_y[n]) adv_y[n][D3(b, i, j, k)] = wy;
                }
        }
    }
    free(fy);
}

/* ------------------------------------------------------------------------------------------------
 * advect_tracer_mdppm (OTA:5990-6494) with ppm_limit_cw84 / _ifc / _sh (OTA:6510-6657): multi-dimensional
 * piecewise-parabolic scheme; tracer_mdppm and tmask_mdppm live on a halo-4 scratch (h4).
 * ---------------------------------------------------------------------------------------------- */
#define NX4(b) ((b)->ni + 8)
#define NY4(b) ((b)->nj + 8)
#define Q2(b, i, j) ((size_t)((i) + 3) + (size_t)NX4(b) * (size_t)((j) + 3))                 /* 2-D h4 work arrays */
#define Q3(b, i, j, k) (Q2(b, i, j) + (size_t)NX4(b) * NY4(b) * (size_t)((k)-1))             /* i, j = -3 .. n+4 */
static const double R12 = 1. / 12., TWOTHIRDS = 2. / 3., FOURTHIRDS = 4. / 3.;

static inline double max3(double a, double b, double c) { return orc_max(orc_max(a, b), c); }
static inline double min3(double a, double b, double c) { return orc_min(orc_min(a, b), c); }
static inline double max4(double a, double b, double c, double d) { return orc_max(max3(a, b, c), d); }
static inline double min4(double a, double b, double c, double d) { return orc_min(min3(a, b, c), d); }

/* slope estimate + monotonic constraint (OTA:6060-6098, 6214-6245, 6349-6380); S = Sim2,Sim1,Si,Sip1,Sip2, m likewise */
static inline double ppm_slope(const double *S, const double *m)
{
    double da2 = 0.5 * (S[3] - S[1]);
    double da3m = R12 * ((((2. * S[0]) - (12. * S[1])) + (6. * S[2])) + (4. * S[3]));
    double da3p = R12 * ((((-(4. * S[1])) - (6. * S[2])) + (12. * S[3])) - (2. * S[4]));
    double da = (((m[0] * (1. - (0.5 * m[4]))) * da3m) + ((m[4] * (1. - (0.5 * m[0]))) * da3p)) + (((1. - m[0]) * (1. - m[4])) * da2);
    double dMx = max3(S[3], S[1], S[2]) - S[2];
    double dMn = S[2] - min3(S[3], S[1], S[2]);
    return (((copysign(1., da) * orc_min(fabs(da), 2. * orc_min(dMx, dMn))) * m[1]) * m[3]) * m[2];
}

typedef struct { double d1m, d1p, d1mm, d1pp; } ppm_d1;
static inline ppm_d1 ppm_diffs(const double *S, const double *m)
{
    ppm_d1 d;
    d.d1m = (S[2] - S[1]) * m[1];
    d.d1p = (S[3] - S[2]) * m[3];
    d.d1mm = ((S[1] - S[0]) * m[1]) * m[0];
    d.d1pp = ((S[4] - S[3]) * m[3]) * m[4];
    return d;
}

/* edge values (Lin 1994 eq. B2) then one of the three limiters, all point-wise */
static inline void ppm_edges(int limiter, double Si, ppm_d1 d, double da_m, double da_0, double da_p, double *aLo, double *aRo)
{
    double Sim1 = Si - d.d1m, Sip1 = Si + d.d1p;
    double aL = (0.5 * (Sim1 + Si)) + (ONESIXTH * (da_m - da_0));
    double aR = (0.5 * (Si + Sip1)) + (ONESIXTH * (da_0 - da_p));
    if (limiter == 1) { /* ppm_limit_cw84, OTA:6520-6536 */
        if ((aR - Si) * (Si - aL) <= 0.) { aL = Si; aR = Si; }
        double da2 = aR - aL, da4 = 0.5 * (aR + aL);
        double da3m = (6. * da2) * (Si - da4), da3p = da2 * da2;
        if (da3m > da3p) aL = (3. * Si) - (2. * aR);
        if (da3m < -da3p) aR = (3. * Si) - (2. * aL);
    } else if (limiter == 2) { /* ppm_limit_ifc, OTA:6565-6575 */
        double ada = fabs(da_0), sda = copysign(1., da_0);
        aL = Si - (sda * orc_min(ada, fabs(aL - Si)));
        aR = Si + (sda * orc_min(ada, fabs(aR - Si)));
    } else { /* ppm_limit_sh, OTA:6614-6655 */
        double z = d.d1m - d.d1mm, w = d.d1p - d.d1m;
        double x = (4. * w) - z, y = (4. * z) - w;
        double dM4m = orc_max(0., min4(x, y, z, w)) + orc_min(0., max4(x, y, z, w));
        z = d.d1pp - d.d1p;
        x = (4. * w) - z;
        y = (4. * z) - w;
        double dM4p = orc_max(0., min4(x, y, z, w)) + orc_min(0., max4(x, y, z, w));
        double qAV = 0.5 * (Si + Sip1);
        x = d.d1p;
        y = 3. * d.d1m;
        double qMP = (Si + orc_max(0., orc_min(x, y))) + orc_min(0., orc_max(x, y));
        double qUL = Si + y;
        double qLC = (Si + (0.5 * d.d1m)) + (FOURTHIRDS * dM4m);
        double qMD = qAV - (0.5 * dM4p);
        double qMin = orc_max(min3(qMD, Si, Sip1), min3(Si, qUL, qLC));
        double qMax = orc_min(max3(qMD, Si, Sip1), max3(Si, qUL, qLC));
        if ((aR - Si) * (aR - qMP) > 1.e-10) aR = orc_min(orc_max(aR, qMin), qMax);
        qAV = 0.5 * (Si + Sim1);
        x = -d.d1m;
        y = -3. * d.d1p;
        qMP = (Si + orc_max(0., orc_min(x, y))) + orc_min(0., orc_max(x, y));
        qUL = Si + y;
        qLC = (Si - (0.5 * d.d1p)) + (FOURTHIRDS * dM4p);
        qMD = qAV - (0.5 * dM4m);
        qMin = orc_max(min3(qMD, Si, Sim1), min3(Si, qUL, qLC));
        qMax = orc_min(max3(qMD, Si, Sim1), max3(Si, qUL, qLC));
        if ((aL - Si) * (aL - qMP) > 1.e-10) aL = orc_min(orc_max(aL, qMin), qMax);
    }
    *aLo = aL;
    *aRo = aR;
}

/* OTA:6036-6201: vertical fluxes and update.  tr (h4) := 0, compute domain := updated tracer; flux_z on the compute domain. */
void orc_mdppm_z(const orc_block *b, double dtime, int limiter, const double *T, const double *w, const double *rho,
                 const double *m4, double *tr, double *flux_z)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    memset(tr, 0, sizeof(double) * (size_t)NX4(b) * NY4(b) * nk);
    double *dak = (double *)calloc((size_t)nk + 2, sizeof(double)), *aL = (double *)calloc((size_t)nk + 2, sizeof(double));
    double *aR = (double *)calloc((size_t)nk + 2, sizeof(double));
    for (int j = 1; j <= nj; j++)
        for (int i = 1; i <= ni; i++) {
            double S[5], m[5];
#define ZCELL(k)                                                                                              \
    do {                                                                                                      \
        int km2 = imax((k)-2, 1), km1 = imax((k)-1, 1), kp1 = imin((k) + 1, nk), kp2 = imin((k) + 2, nk);     \
        S[0] = T[D3(b, i, j, km2)]; S[1] = T[D3(b, i, j, km1)]; S[2] = T[D3(b, i, j, (k))];                   \
        S[3] = T[D3(b, i, j, kp1)]; S[4] = T[D3(b, i, j, kp2)];                                               \
        m[0] = m4[Q3(b, i, j, km2)] * (double)(km1 - km2); m[1] = m4[Q3(b, i, j, km1)] * (double)((k)-km1);   \
        m[2] = m4[Q3(b, i, j, (k))];                                                                          \
        m[3] = m4[Q3(b, i, j, kp1)] * (double)(kp1 - (k)); m[4] = m4[Q3(b, i, j, kp2)] * (double)(kp2 - kp1); \
    } while (0)
            for (int k = 1; k <= nk; k++) { ZCELL(k); dak[k] = ppm_slope(S, m); }
            for (int k = 1; k <= nk; k++) {
                int km1 = imax(k - 1, 1), kp1 = imin(k + 1, nk);
                ZCELL(k);
                ppm_edges(limiter, S[2], ppm_diffs(S, m), dak[km1], dak[k], dak[kp1], &aL[k], &aR[k]);
                double a6 = (6. * S[2]) - (3. * (aR[k] + aL[k]));
                double dat = b->dat[D2(b, i, j)];
                double massflux = ((dat * w[W3(b, i, j, k)]) * m[2]) * m[3];
                if (massflux <= 0.0) {
                    double cfl = fabs((w[W3(b, i, j, k)] * dtime) / rho[D3(b, i, j, k)]);
                    flux_z[D3(b, i, j, k)] = massflux * (aR[k] + ((0.5 * cfl) * ((aL[k] - aR[k]) + ((1. - (TWOTHIRDS * cfl)) * a6))));
                }
                massflux = ((dat * w[W3(b, i, j, km1)]) * m[2]) * m[1];
                if (massflux > 0.0) {
                    double cfl = fabs((w[W3(b, i, j, km1)] * dtime) / rho[D3(b, i, j, km1)]);
                    flux_z[D3(b, i, j, km1)] = massflux * (aL[k] + ((0.5 * cfl) * ((aR[k] - aL[k]) + ((1. - (TWOTHIRDS * cfl)) * a6))));
                }
            }
#undef ZCELL
            for (int k = 1; k <= nk; k++) { /* OTA:6180-6190 */
                int km1 = imax(k - 1, 1);
                double mskm1 = (double)(k - km1), Tk = T[D3(b, i, j, k)];
                tr[Q3(b, i, j, k)] = Tk + ((dtime / rho[D3(b, i, j, k)]) *
                                           ((b->datr[D2(b, i, j)] * (flux_z[D3(b, i, j, k)] - (mskm1 * flux_z[D3(b, i, j, km1)]))) +
                                            (Tk * ((mskm1 * w[W3(b, i, j, km1)]) - w[W3(b, i, j, k)]))));
            }
        }
    free(dak); free(aL); free(aR);
}

/* one horizontal direction: dir 0 = x (OTA:6205-6328), 1 = y (OTA:6344-6455); fluxes only */
static void mdppm_hflux(const orc_block *b, double dtime, int limiter, int dir, const double *vel, const double *rho,
                        const double *m4, const double *tr, double *flux)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const int di = dir == 0, dj = dir == 1;
    const size_t n2 = (size_t)NX4(b) * NY4(b);
    double *da = (double *)calloc(n2, sizeof(double)), *aL = (double *)calloc(n2, sizeof(double));
    double *aR = (double *)calloc(n2, sizeof(double)), *a6 = (double *)calloc(n2, sizeof(double));
    ppm_d1 *dd = (ppm_d1 *)calloc(n2, sizeof(ppm_d1));
    const double *met_f = dir == 0 ? b->dyte : b->dxtn, *met_c = dir == 0 ? b->dxte : b->dytn;
    for (int k = 1; k <= nk; k++) {
        for (int j = 1 - 2 * dj; j <= nj + 2 * dj; j++)
            for (int i = 1 - 2 * di; i <= ni + 2 * di; i++) {
                double S[5], m[5];
                for (int q = 0; q < 5; q++) {
                    S[q] = tr[Q3(b, i + (q - 2) * di, j + (q - 2) * dj, k)];
                    m[q] = m4[Q3(b, i + (q - 2) * di, j + (q - 2) * dj, k)];
                }
                da[Q2(b, i, j)] = ppm_slope(S, m);
                dd[Q2(b, i, j)] = ppm_diffs(S, m);
            }
        for (int j = 1 - dj; j <= nj + dj; j++)
            for (int i = 1 - di; i <= ni + di; i++) {
                double Si = tr[Q3(b, i, j, k)];
                ppm_edges(limiter, Si, dd[Q2(b, i, j)], da[Q2(b, i - di, j - dj)], da[Q2(b, i, j)], da[Q2(b, i + di, j + dj)],
                          &aL[Q2(b, i, j)], &aR[Q2(b, i, j)]);
                a6[Q2(b, i, j)] = (6. * Si) - (3. * (aR[Q2(b, i, j)] + aL[Q2(b, i, j)]));
            }
        for (int j = 1 - dj; j <= nj; j++)
            for (int i = 1 - di; i <= ni; i++) {
                double vv = vel[D3(b, i, j, k)];
                double massflux = met_f[D2(b, i, j)] * vv;
                double cfl = ((vv * dtime) * 2.0) / ((rho[D3(b, i, j, k)] + rho[D3(b, i + di, j + dj, k)]) * met_c[D2(b, i, j)]);
                double mm = (massflux * m4[Q3(b, i, j, k)]) * m4[Q3(b, i + di, j + dj, k)];
                size_t c = Q2(b, i, j), e = Q2(b, i + di, j + dj);
                if (massflux >= 0.0)
                    flux[D3(b, i, j, k)] = mm * (aR[c] + ((0.5 * cfl) * ((aL[c] - aR[c]) + ((1. - (TWOTHIRDS * cfl)) * a6[c]))));
                else
                    flux[D3(b, i, j, k)] = mm * (aL[e] - ((0.5 * cfl) * ((aR[e] - aL[e]) + ((1. + (TWOTHIRDS * cfl)) * a6[e]))));
            }
    }
    free(da); free(aL); free(aR); free(a6); free(dd);
}

void orc_mdppm_x(const orc_block *b, double dtime, int limiter, const double *T, const double *u, const double *rho,
                 const double *m4, double *tr, double *flux_x)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    mdppm_hflux(b, dtime, limiter, 0, u, rho, m4, tr, flux_x);
    for (int k = 1; k <= nk; k++) /* OTA:6317-6327 */
        for (int j = 1; j <= nj; j++)
            for (int i = 1; i <= ni; i++)
                tr[Q3(b, i, j, k)] = tr[Q3(b, i, j, k)] +
                                     ((((dtime * m4[Q3(b, i, j, k)]) * b->datr[D2(b, i, j)]) / rho[D3(b, i, j, k)]) *
                                      ((flux_x[D3(b, i - 1, j, k)] - flux_x[D3(b, i, j, k)]) +
                                       (T[D3(b, i, j, k)] * ((b->dyte[D2(b, i, j)] * u[D3(b, i, j, k)]) - (b->dyte[D2(b, i - 1, j)] * u[D3(b, i - 1, j, k)])))));
}

/* y sweep + overall tendency; wrk1_out = Tracer%wrk1 (the dispatcher's negation, OTA:1966-1968, applied) */
void orc_mdppm_y(const orc_block *b, double dtime, int limiter, const double *T, const double *u, const double *v, const double *w,
                 const double *rho, const double *m4, double *tr, double *flux_y, double *wrk1_out)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    mdppm_hflux(b, dtime, limiter, 1, v, rho, m4, tr, flux_y);
    for (int k = 1; k <= nk; k++)
        for (int j = 1; j <= nj; j++)
            for (int i = 1; i <= ni; i++) {
                size_t q = D3(b, i, j, k), c = D2(b, i, j);
                double t = tr[Q3(b, i, j, k)] + ((((dtime * b->tmask[q]) * b->datr[c]) / rho[q]) * (flux_y[D3(b, i, j - 1, k)] - flux_y[q]));
                double wkm1 = (k > 1) ? w[W3(b, i, j, k - 1)] : 0.0;
                t = t + (((dtime * T[q]) / rho[q]) *
                         ((w[W3(b, i, j, k)] - wkm1) + (b->datr[c] * ((b->dyte[D2(b, i - 1, j)] * u[D3(b, i - 1, j, k)]) - (b->dyte[c] * u[q])))));
                tr[Q3(b, i, j, k)] = t;
                double f = (((-rho[q]) * (t - T[q])) / dtime) * m4[Q3(b, i, j, k)];
                wrk1_out[q] = -f;
            }
}

/* ------------------------------------------------------------------------------------------------
 * compute_adv_diss (OTA:7547-7712): dissipation from advection truncation errors.  The scheme's operators are applied
 * to the squared tracer by the caller (oracle.py: Oracle.adv_diss, same functions as the dispatcher arms); here the two
 * element-wise parts.
 * ---------------------------------------------------------------------------------------------- */
void orc_square(const orc_block *b, const double *T, double *out) /* OTA:7574-7580, data domain */
{
    const size_t n = (size_t)NX1(b) * NY1(b) * b->nk;
    for (size_t q = 0; q < n; q++) out[q] = T[q] * T[q];
}

/* OTA:7680-7701.  wrk2 / wrk3 = horizontal / vertical operator on T**2 (data-domain arrays, compute domain filled);
 * t2_tendency = wrk1 (zero outside the compute domain), diss = wrk4 (zero outside). */
void orc_adv_diss_final(const orc_block *b, double dtime, double conversion, const double *rho_tau, const double *rho_taup1,
                        const double *T_tau, const double *advect_tendency, const double *wrk2, const double *wrk3,
                        double *t2_tendency, double *diss)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const size_t n = (size_t)NX1(b) * NY1(b) * nk;
    const double dtimer = 1.0 / dtime;
    memset(t2_tendency, 0, sizeof(double) * n);
    memset(diss, 0, sizeof(double) * n);
    for (int k = 1; k <= nk; k++)
        for (int j = 1; j <= nj; j++)
            for (int i = 1; i <= ni; i++) {
                size_t q = D3(b, i, j, k);
                double w1 = wrk2[q] + wrk3[q];
                double at = advect_tendency[q];
                double term1 = at * (((2.0 * rho_tau[q]) * T_tau[q]) + (dtime * at));
                double term2 = -(rho_taup1[q] * w1);
                t2_tendency[q] = w1;
                diss[q] = (-(conversion * conversion) * dtimer) * (term1 + term2);
            }
}

/* ------------------------------------------------------------------------------------------------
 * advect_tracer_mdfl_sweby_test (OTA:3469-3746): the mass-weighted variant.  Three running fields on the halo-2
 * scratch: tracer_mdfl (tr), tracermass_mdfl (tms), mass_mdfl (ms); the CFL number is |massflux|*dtime/mass of
 * the upwind cell; theta's denominator is sign(1e-30,Rj)+Rj (sign BIT of Rj, as IEEE processors do).
 * ---------------------------------------------------------------------------------------------- */
static inline double test_flux(double Rjp, double Rj, double Rjm, double massflux, double cfl, double Tup, double Tdn,
                               double mA, double mB, double sl)
{
    double d0 = ((2.0 - cfl) * (1.0 - cfl)) * ONESIXTH;
    double d1 = (1.0 - (cfl * cfl)) * ONESIXTH;
    double den = copysign(1.0e-30, Rj) + Rj;
    double thetaP = Rjm / den;
    double thetaM = Rjp / den;
    double psiP = d0 + (d1 * thetaP);
    psiP = (psiP * (1.0 - sl)) +
           (orc_max(0.0, orc_min(orc_min(1.0, d0 + (d1 * thetaP)), ((1.0 - cfl) / (1.0e-30 + cfl)) * thetaP)) * sl);
    double psiM = d0 + (d1 * thetaM);
    psiM = (psiM * (1.0 - sl)) +
           (orc_max(0.0, orc_min(orc_min(1.0, d0 + (d1 * thetaM)), ((1.0 - cfl) / (1.0e-30 + cfl)) * thetaM)) * sl);
    return ((0.5 * (((massflux + fabs(massflux)) * (Tup + (psiP * Rj))) + ((massflux - fabs(massflux)) * (Tdn - (psiM * Rj))))) * mA) * mB;
}

/* OTA:3505-3580.  tr/tms/ms: h2 scratch, zeroed here (OTA:3500-3502); flux_z optional. */
void orc_sweby_test_z(const orc_block *b, double dtime, double sl, const double *T, const double *w, const double *rho,
                      double *tr, double *tms, double *ms, double *flux_z)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *m = b->tmask_h2;
    const size_t nh = (size_t)NX2(b) * NY2(b) * nk;
    double *ftp = (double *)calloc((size_t)ni * nj, sizeof(double));
    double *wkm1 = (double *)calloc((size_t)ni * nj, sizeof(double));
    memset(tr, 0, sizeof(double) * nh);
    memset(tms, 0, sizeof(double) * nh);
    memset(ms, 0, sizeof(double) * nh);
    for (int k = 1; k <= nk; k++) {
        int kp1 = imin(k + 1, nk), kp2 = imin(k + 2, nk), km1 = imax(k - 1, 1);
        for (int j = 1; j <= nj; j++)
            for (int i = 1; i <= ni; i++) {
                size_t c = (size_t)(i - 1) + (size_t)ni * (j - 1), h = H3(b, i, j, k);
                double Tk = T[D3(b, i, j, k)], Tkp1 = T[D3(b, i, j, kp1)];
                double dat = b->dat[D2(b, i, j)];
                tr[h] = Tk;
                ms[h] = rho[D3(b, i, j, k)] * dat;
                tms[h] = ms[h] * Tk;
                double Rjp = ((T[D3(b, i, j, km1)] - Tk) * m[H3(b, i, j, km1)]) * m[H3(b, i, j, k)];
                double Rj = ((Tk - Tkp1) * m[H3(b, i, j, k)]) * m[H3(b, i, j, kp1)];
                double Rjm = ((Tkp1 - T[D3(b, i, j, kp2)]) * m[H3(b, i, j, kp1)]) * m[H3(b, i, j, kp2)];
                double wk = w[W3(b, i, j, k)];
                double massflux = dat * wk, cfl;
                if (massflux * rho[D3(b, i, j, kp1)] > 0.0) cfl = (fabs(wk) * dtime) / rho[D3(b, i, j, kp1)];
                else if (massflux * rho[D3(b, i, j, k)] < 0.0) cfl = (fabs(wk) * dtime) / rho[D3(b, i, j, k)];
                else cfl = 0.0;
                double fbt = test_flux(Rjp, Rj, Rjm, massflux, cfl, Tkp1, Tk, m[H3(b, i, j, kp1)], m[H3(b, i, j, k)], sl);
                ms[h] = ms[h] + ((dtime * dat) * (wk - wkm1[c]));
                tms[h] = tms[h] + (dtime * (fbt - ftp[c]));
                if (ms[h] > 0.) tr[h] = tms[h] / ms[h];
                if (flux_z) flux_z[D3(b, i, j, k)] = fbt;
                ftp[c] = fbt;
                wkm1[c] = wk;
            }
    }
    free(ftp);
    free(wkm1);
}

/* OTA:3585-3649.  flux_x: data-domain array, zeroed by the caller (OTA:3503), faces i = 0..ni written. */
void orc_sweby_test_x(const orc_block *b, double dtime, double sl, const double *u, double *tr, double *tms, double *ms,
                      double *flux_x)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *m = b->tmask_h2;
    for (int k = 1; k <= nk; k++) {
        for (int j = 1; j <= nj; j++)
            for (int i = 0; i <= ni; i++) {
                double t0 = tr[H3(b, i, j, k)], t1 = tr[H3(b, i + 1, j, k)];
                double Rjp = ((tr[H3(b, i + 2, j, k)] - t1) * m[H3(b, i + 2, j, k)]) * m[H3(b, i + 1, j, k)];
                double Rj = ((t1 - t0) * m[H3(b, i + 1, j, k)]) * m[H3(b, i, j, k)];
                double Rjm = ((t0 - tr[H3(b, i - 1, j, k)]) * m[H3(b, i, j, k)]) * m[H3(b, i - 1, j, k)];
                double massflux = b->dyte[D2(b, i, j)] * u[D3(b, i, j, k)], cfl;
                if (massflux * ms[H3(b, i, j, k)] > 0.0) cfl = (fabs(massflux) * dtime) / ms[H3(b, i, j, k)];
                else if (massflux * ms[H3(b, i + 1, j, k)] < 0.0) cfl = (fabs(massflux) * dtime) / ms[H3(b, i + 1, j, k)];
                else cfl = 0.0;
                flux_x[D3(b, i, j, k)] = test_flux(Rjp, Rj, Rjm, massflux, cfl, t0, t1, m[H3(b, i, j, k)], m[H3(b, i + 1, j, k)], sl);
            }
        for (int j = 1; j <= nj; j++)
            for (int i = 1; i <= ni; i++) {
                size_t h = H3(b, i, j, k);
                ms[h] = ms[h] + (dtime * ((b->dyte[D2(b, i - 1, j)] * u[D3(b, i - 1, j, k)]) - (b->dyte[D2(b, i, j)] * u[D3(b, i, j, k)])));
                tms[h] = tms[h] + (dtime * (flux_x[D3(b, i - 1, j, k)] - flux_x[D3(b, i, j, k)]));
                if (ms[h] > 0.) tr[h] = tms[h] / ms[h];
            }
    }
}

/* OTA:3655-3736.  wrk1_out = Tracer%wrk1 on the compute domain: the caller's negation (OTA:1970-1975) applied. */
void orc_sweby_test_y(const orc_block *b, double dtime, double sl, const double *T, const double *v, const double *rho,
                      double *tr, double *tms, double *ms, double *flux_y, double *wrk1_out)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *m = b->tmask_h2;
    for (int k = 1; k <= nk; k++) {
        for (int j = 0; j <= nj; j++)
            for (int i = 1; i <= ni; i++) {
                double t0 = tr[H3(b, i, j, k)], t1 = tr[H3(b, i, j + 1, k)];
                double Rjp = ((tr[H3(b, i, j + 2, k)] - t1) * m[H3(b, i, j + 2, k)]) * m[H3(b, i, j + 1, k)];
                double Rj = ((t1 - t0) * m[H3(b, i, j + 1, k)]) * m[H3(b, i, j, k)];
                double Rjm = ((t0 - tr[H3(b, i, j - 1, k)]) * m[H3(b, i, j, k)]) * m[H3(b, i, j - 1, k)];
                double massflux = b->dxtn[D2(b, i, j)] * v[D3(b, i, j, k)], cfl;
                if (massflux * ms[H3(b, i, j, k)] > 0.0) cfl = (fabs(massflux) * dtime) / ms[H3(b, i, j, k)];
                else if (massflux * ms[H3(b, i, j + 1, k)] < 0.0) cfl = (fabs(massflux) * dtime) / ms[H3(b, i, j + 1, k)];
                else cfl = 0.0;
                flux_y[D3(b, i, j, k)] = test_flux(Rjp, Rj, Rjm, massflux, cfl, t0, t1, m[H3(b, i, j, k)], m[H3(b, i, j + 1, k)], sl);
            }
        for (int j = 1; j <= nj; j++)
            for (int i = 1; i <= ni; i++) {
                size_t h = H3(b, i, j, k);
                ms[h] = ms[h] + (dtime * ((b->dxtn[D2(b, i, j - 1)] * v[D3(b, i, j - 1, k)]) - (b->dxtn[D2(b, i, j)] * v[D3(b, i, j, k)])));
                tms[h] = tms[h] + (dtime * (flux_y[D3(b, i, j - 1, k)] - flux_y[D3(b, i, j, k)]));
                if (ms[h] > 0.) tr[h] = tms[h] / ms[h];
            }
        for (int j = 1; j <= nj; j++)
            for (int i = 1; i <= ni; i++) {
                double f = (m[H3(b, i, j, k)] * ((rho[D3(b, i, j, k)] * T[D3(b, i, j, k)]) - (tms[H3(b, i, j, k)] * b->datr[D2(b, i, j)]))) / dtime;
                wrk1_out[D3(b, i, j, k)] = -f;
            }
    }
}

/* ------------------------------------------------------------------------------------------------
 * advect_tracer_mdfl_sweby (OTA:3806-4066)
 * ---------------------------------------------------------------------------------------------- */
void orc_mdfl_sweby_z(const orc_block *b, double dtime, double sl, const double *T, const double *w,
                      const double *rho, double *tm, double *flux_z)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *m = b->tmask_h2;
    double *ftp = (double *)calloc((size_t)ni * nj, sizeof(double));
    double *wkm1 = (double *)calloc((size_t)ni * nj, sizeof(double));
    memset(tm, 0, sizeof(double) * (size_t)NX2(b) * NY2(b) * nk); /* tracer_mdfl = 0.0 (OTA:3836) */
    for (int k = 1; k <= nk; k++) {
        int kp1 = imin(k + 1, nk), kp2 = imin(k + 2, nk), km1 = imax(k - 1, 1);
        for (int j = 1; j <= nj; j++)
            for (int i = 1; i <= ni; i++) {
                size_t c = (size_t)(i - 1) + (size_t)ni * (j - 1);
                double Tk = T[D3(b, i, j, k)], Tkp1 = T[D3(b, i, j, kp1)];
                double Rjp = ((T[D3(b, i, j, km1)] - Tk) * m[H3(b, i, j, km1)]) * m[H3(b, i, j, k)];
                double Rj = ((Tk - Tkp1) * m[H3(b, i, j, k)]) * m[H3(b, i, j, kp1)];
                double Rjm = ((Tkp1 - T[D3(b, i, j, kp2)]) * m[H3(b, i, j, kp1)]) * m[H3(b, i, j, kp2)];
                double wk = w[W3(b, i, j, k)], r = rho[D3(b, i, j, k)];
                double massflux = b->dat[D2(b, i, j)] * wk;
                double cfl = fabs((wk * dtime) / r);
                double fbt = sweby_flux_sl(Rjp, Rj, Rjm, massflux, cfl, Tkp1, Tk, m[H3(b, i, j, kp1)], m[H3(b, i, j, k)], sl);
                /* OTA:3892-3896 */
                tm[H3(b, i, j, k)] = Tk + ((dtime / r) * ((b->datr[D2(b, i, j)] * (fbt - ftp[c])) + (Tk * (wkm1[c] - wk))));
                if (flux_z) flux_z[D3(b, i, j, k)] = fbt;
                ftp[c] = fbt;
            }
        for (int j = 1; j <= nj; j++)
            for (int i = 1; i <= ni; i++) wkm1[(size_t)(i - 1) + (size_t)ni * (j - 1)] = w[W3(b, i, j, k)];
    }
    free(ftp);
    free(wkm1);
}

void orc_mdfl_sweby_x(const orc_block *b, double dtime, double sl, const double *T, const double *u,
                      const double *rho, double *tm, double *flux_x)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *m = b->tmask_h2;
    double *fx = (double *)malloc(sizeof(double) * (size_t)(ni + 1));
    if (flux_x) memset(flux_x, 0, sizeof(double) * (size_t)NX1(b) * NY1(b) * nk); /* OTA:3837 */
    for (int k = 1; k <= nk; k++)
        for (int j = 1; j <= nj; j++) {
            for (int i = 0; i <= ni; i++) { /* OTA:3918-3955 */
                double t0 = tm[H3(b, i, j, k)], t1 = tm[H3(b, i + 1, j, k)];
                double Rjp = ((tm[H3(b, i + 2, j, k)] - t1) * m[H3(b, i + 2, j, k)]) * m[H3(b, i + 1, j, k)];
                double Rj = ((t1 - t0) * m[H3(b, i + 1, j, k)]) * m[H3(b, i, j, k)];
                double Rjm = ((t0 - tm[H3(b, i - 1, j, k)]) * m[H3(b, i, j, k)]) * m[H3(b, i - 1, j, k)];
                double uu = u[D3(b, i, j, k)];
                double massflux = b->dyte[D2(b, i, j)] * uu;
                double cfl = fabs(((uu * dtime) * 2.0) /
                                  ((rho[D3(b, i, j, k)] + rho[D3(b, i + 1, j, k)]) * b->dxte[D2(b, i, j)]));
                fx[i] = sweby_flux_sl(Rjp, Rj, Rjm, massflux, cfl, t0, t1, m[H3(b, i, j, k)], m[H3(b, i + 1, j, k)], sl);
                if (flux_x) flux_x[D3(b, i, j, k)] = fx[i];
            }
            for (int i = 1; i <= ni; i++) { /* OTA:3958-3967 */
                double coef = ((dtime * m[H3(b, i, j, k)]) * b->datr[D2(b, i, j)]) / rho[D3(b, i, j, k)];
                tm[H3(b, i, j, k)] =
                    tm[H3(b, i, j, k)] +
                    (coef * ((fx[i - 1] - fx[i]) +
                             (T[D3(b, i, j, k)] * ((b->dyte[D2(b, i, j)] * u[D3(b, i, j, k)]) -
                                                   (b->dyte[D2(b, i - 1, j)] * u[D3(b, i - 1, j, k)])))));
            }
        }
    free(fx);
}

void orc_mdfl_sweby_y(const orc_block *b, double dtime, double sl, const double *T, const double *u,
                      const double *v, const double *w, const double *rho, double *tm, double *flux_y,
                      double *wrk1_out)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *m = b->tmask_h2;
    double *fy = (double *)malloc(sizeof(double) * (size_t)ni * (nj + 1));
    if (flux_y) memset(flux_y, 0, sizeof(double) * (size_t)NX1(b) * NY1(b) * nk); /* OTA:3838 */
    memset(wrk1_out, 0, sizeof(double) * (size_t)NX1(b) * NY1(b) * nk);           /* OTA:1925-1931 */
    for (int k = 1; k <= nk; k++) {
        for (int j = 0; j <= nj; j++) /* OTA:3982-4020 */
            for (int i = 1; i <= ni; i++) {
                double t0 = tm[H3(b, i, j, k)], t1 = tm[H3(b, i, j + 1, k)];
                double Rjp = ((tm[H3(b, i, j + 2, k)] - t1) * m[H3(b, i, j + 2, k)]) * m[H3(b, i, j + 1, k)];
                double Rj = ((t1 - t0) * m[H3(b, i, j + 1, k)]) * m[H3(b, i, j, k)];
                double Rjm = ((t0 - tm[H3(b, i, j - 1, k)]) * m[H3(b, i, j, k)]) * m[H3(b, i, j - 1, k)];
                double vv = v[D3(b, i, j, k)];
                double massflux = b->dxtn[D2(b, i, j)] * vv;
                double cfl = fabs(((vv * dtime) * 2.0) /
                                  ((rho[D3(b, i, j, k)] + rho[D3(b, i, j + 1, k)]) * b->dytn[D2(b, i, j)]));
                double f = sweby_flux_sl(Rjp, Rj, Rjm, massflux, cfl, t0, t1, m[H3(b, i, j, k)], m[H3(b, i, j + 1, k)], sl);
                fy[(size_t)(i - 1) + (size_t)ni * j] = f;
                if (flux_y) flux_y[D3(b, i, j, k)] = f;
            }
        for (int j = 1; j <= nj; j++)
            for (int i = 1; i <= ni; i++) {
                double Tk = T[D3(b, i, j, k)], r = rho[D3(b, i, j, k)];
                double wkm1 = (k == 1) ? 0.0 : w[W3(b, i, j, k - 1)];
                /* OTA:4025-4027 */
                double t = tm[H3(b, i, j, k)] +
                           ((((dtime * m[H3(b, i, j, k)]) * b->datr[D2(b, i, j)]) / r) *
                            (fy[(size_t)(i - 1) + (size_t)ni * (j - 1)] - fy[(size_t)(i - 1) + (size_t)ni * j]));
                /* OTA:4035-4040 */
                t = t + (((dtime * Tk) / r) *
                         ((w[W3(b, i, j, k)] - wkm1) +
                          (b->datr[D2(b, i, j)] * ((b->dyte[D2(b, i - 1, j)] * u[D3(b, i - 1, j, k)]) -
                                                   (b->dyte[D2(b, i, j)] * u[D3(b, i, j, k)])))));
                tm[H3(b, i, j, k)] = t;
                /* function value = -(rho*(tm-T)/dtime*m) (OTA:4043-4045); caller stores wrk1 = -value (OTA:1958-1959) */
                double fval = -(((r * (t - Tk)) / dtime) * m[H3(b, i, j, k)]);
                wrk1_out[D3(b, i, j, k)] = -fval;
            }
    }
    free(fy);
}

/* ------------------------------------------------------------------------------------------------
 * quicker_init (OTA:1442-1586)
 * ---------------------------------------------------------------------------------------------- */
void orc_quicker_init_pre(const orc_block *b)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    memset(b->tmask_h2, 0, sizeof(double) * (size_t)NX2(b) * NY2(b) * nk);
    memset(b->dxt_h2, 0, sizeof(double) * (size_t)NX2(b) * NY2(b));
    memset(b->dyt_h2, 0, sizeof(double) * (size_t)NX2(b) * NY2(b));
    for (int i = 1; i <= ni; i++) /* OTA:1478-1486 */
        for (int j = 1; j <= nj; j++) {
            b->dxt_h2[H2(b, i, j)] = b->dxt[D2(b, i, j)];
            b->dyt_h2[H2(b, i, j)] = b->dyt[D2(b, i, j)];
            for (int k = 1; k <= nk; k++) b->tmask_h2[H3(b, i, j, k)] = b->tmask[D3(b, i, j, k)];
        }
}

void orc_quicker_init_edges(const orc_block *b)
{
    const int ni = b->ni, nj = b->nj;
    double *dx = b->dxt_h2, *dy = b->dyt_h2;
    for (int i = -1; i <= 0; i++) { /* OTA:1491-1494 */
        for (int j = 1; j <= nj; j++) dx[H2(b, i, j)] = dx[H2(b, 1, j)];
        for (int j = -1; j <= nj + 2; j++) dy[H2(b, i, j)] = dy[H2(b, 1, j)];
    }
    for (int i = ni + 1; i <= ni + 2; i++) { /* OTA:1496-1499 */
        for (int j = 1; j <= nj; j++) dx[H2(b, i, j)] = dx[H2(b, ni, j)];
        for (int j = -1; j <= nj + 2; j++) dy[H2(b, i, j)] = dy[H2(b, ni, j)];
    }
    for (int j = -1; j <= 0; j++) /* OTA:1501-1504 */
        for (int i = -1; i <= ni + 2; i++) {
            dx[H2(b, i, j)] = dx[H2(b, i, 1)];
            dy[H2(b, i, j)] = dy[H2(b, i, 1)];
        }
    for (int j = nj + 1; j <= nj + 2; j++) /* OTA:1506-1509 */
        for (int i = -1; i <= ni + 2; i++) {
            dx[H2(b, i, j)] = dx[H2(b, i, nj)];
            dy[H2(b, i, j)] = dy[H2(b, i, nj)];
        }
}

#define Q2(b, i, j, c) ((size_t)(i) + (size_t)NX1(b) * ((size_t)(j) + (size_t)NY1(b) * (size_t)((c)-1)))

void orc_quicker_init_weights(const orc_block *b)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *dx = b->dxt_h2, *dy = b->dyt_h2;
    size_t n2 = (size_t)NX1(b) * NY1(b);
    memset(b->quick_x, 0, sizeof(double) * n2 * 2);
    memset(b->quick_y, 0, sizeof(double) * n2 * 2);
    memset(b->curv_xp, 0, sizeof(double) * n2 * 3);
    memset(b->curv_xn, 0, sizeof(double) * n2 * 3);
    memset(b->curv_yp, 0, sizeof(double) * n2 * 3);
    memset(b->curv_yn, 0, sizeof(double) * n2 * 3);
    for (int i = 0; i <= ni; i++) /* OTA:1516-1564 */
        for (int j = 0; j <= nj; j++) {
            double xm = dx[H2(b, i - 1, j)], x0 = dx[H2(b, i, j)], x1 = dx[H2(b, i + 1, j)], x2 = dx[H2(b, i + 2, j)];
            double ym = dy[H2(b, i, j - 1)], y0 = dy[H2(b, i, j)], y1 = dy[H2(b, i, j + 1)], y2 = dy[H2(b, i, j + 2)];
            b->quick_x[Q2(b, i, j, 1)] = x1 / (x1 + x0);
            b->quick_x[Q2(b, i, j, 2)] = x0 / (x1 + x0);
            b->quick_y[Q2(b, i, j, 1)] = y1 / (y1 + y0);
            b->quick_y[Q2(b, i, j, 2)] = y0 / (y1 + y0);

            b->curv_xp[Q2(b, i, j, 1)] = (x0 * x1) / (((xm + (2.0 * x0)) + x1) * (x0 + x1));
            b->curv_xp[Q2(b, i, j, 2)] = -((x0 * x1) / ((x0 + x1) * (xm + x0)));
            b->curv_xp[Q2(b, i, j, 3)] = (x0 * x1) / (((xm + (2.0 * x0)) + x1) * (xm + x0));

            b->curv_xn[Q2(b, i, j, 1)] = (x0 * x1) / (((x0 + (2.0 * x1)) + x2) * (x1 + x2));
            b->curv_xn[Q2(b, i, j, 2)] = -((x0 * x1) / ((x1 + x2) * (x0 + x1)));
            b->curv_xn[Q2(b, i, j, 3)] = (x0 * x1) / (((x0 + (2.0 * x1)) + x2) * (x0 + x1));

            b->curv_yp[Q2(b, i, j, 1)] = (y0 * y1) / (((ym + (2.0 * y0)) + y1) * (y0 + y1));
            b->curv_yp[Q2(b, i, j, 2)] = -((y0 * y1) / ((y0 + y1) * (ym + y0)));
            b->curv_yp[Q2(b, i, j, 3)] = (y0 * y1) / (((ym + (2.0 * y0)) + y1) * (ym + y0));

            b->curv_yn[Q2(b, i, j, 1)] = (y0 * y1) / (((y0 + (2.0 * y1)) + y2) * (y1 + y2));
            b->curv_yn[Q2(b, i, j, 2)] = -((y0 * y1) / ((y1 + y2) * (y0 + y1)));
            b->curv_yn[Q2(b, i, j, 3)] = (y0 * y1) / (((y0 + (2.0 * y1)) + y2) * (y0 + y1));
        }
    for (int k = 1; k <= nk; k++) { /* OTA:1566-1578 ; quick_z(k,c) at (k-1) + nk*(c-1) */
        int kp2 = imin(k + 2, nk), kp1 = imin(k + 1, nk), km1 = imax(k - 1, 1);
        double zm = b->dzt[km1 - 1], z0 = b->dzt[k - 1], z1 = b->dzt[kp1 - 1], z2 = b->dzt[kp2 - 1];
        b->quick_z[(k - 1) + nk * 0] = z1 / (z1 + z0);
        b->quick_z[(k - 1) + nk * 1] = z0 / (z1 + z0);
        b->curv_zp[(k - 1) + nk * 0] = (z0 * z1) / (((zm + (2.0 * z0)) + z1) * (z0 + z1));
        b->curv_zp[(k - 1) + nk * 1] = -((z0 * z1) / ((z0 + z1) * (zm + z0)));
        b->curv_zp[(k - 1) + nk * 2] = (z0 * z1) / (((zm + (2.0 * z0)) + z1) * (zm + z0));
        b->curv_zn[(k - 1) + nk * 0] = (z0 * z1) / (((z0 + (2.0 * z1)) + z2) * (z1 + z2));
        b->curv_zn[(k - 1) + nk * 1] = -((z0 * z1) / ((z1 + z2) * (z0 + z1)));
        b->curv_zn[(k - 1) + nk * 2] = (z0 * z1) / (((z0 + (2.0 * z1)) + z2) * (z0 + z1));
    }
}

/* tracer_quick = 0; compute domain := T(taum1) (OTA:2558-2565); full halo update by caller (OTA:2566) */
void orc_quicker_prep(const orc_block *b, const double *T_taum1, double *tq)
{
    memset(tq, 0, sizeof(double) * (size_t)NX2(b) * NY2(b) * b->nk);
    for (int k = 1; k <= b->nk; k++)
        for (int j = 1; j <= b->nj; j++)
            for (int i = 1; i <= b->ni; i++) tq[H3(b, i, j, k)] = T_taum1[D3(b, i, j, k)];
}

/* OTA:2568-2637 */
void orc_horz_quicker_flux(const orc_block *b, const double *Tm1, const double *Tt, const double *tq,
                           const double *u, const double *v, const double *tmask_limit, int limit_with_upwind,
                           double *flux_x, double *flux_y)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *mq = b->tmask_h2;
    memset(flux_x, 0, sizeof(double) * (size_t)NX1(b) * NY1(b) * nk);
    memset(flux_y, 0, sizeof(double) * (size_t)NX1(b) * NY1(b) * nk);
    for (int k = 1; k <= nk; k++) {
        for (int j = 0; j <= nj; j++) /* OTA:2572-2588 */
            for (int i = 1; i <= ni; i++) {
                double vel = b->dxtn[D2(b, i, j)] * v[D3(b, i, j, k)];
                double upos = 0.5 * (vel + fabs(vel));
                double uneg = 0.5 * (vel - fabs(vel));
                double rnormsk = mq[H3(b, i, j, k)] * (1.0 - mq[H3(b, i, j - 1, k)]);
                double soutmsk = mq[H3(b, i, j + 1, k)] * (1.0 - mq[H3(b, i, j + 2, k)]);
                double t0 = tq[H3(b, i, j, k)], t1 = tq[H3(b, i, j + 1, k)];
                flux_y[D3(b, i, j, k)] =
                    ((vel * ((b->quick_y[Q2(b, i, j, 1)] * Tt[D3(b, i, j, k)]) +
                             (b->quick_y[Q2(b, i, j, 2)] * Tt[D3(b, i, j + 1, k)]))) -
                     (upos * (((b->curv_yp[Q2(b, i, j, 1)] * t1) + (b->curv_yp[Q2(b, i, j, 2)] * t0)) +
                              (b->curv_yp[Q2(b, i, j, 3)] * ((tq[H3(b, i, j - 1, k)] * (1.0 - rnormsk)) + (t0 * rnormsk)))))) -
                    (uneg * (((b->curv_yn[Q2(b, i, j, 1)] * ((tq[H3(b, i, j + 2, k)] * (1.0 - soutmsk)) + (t1 * soutmsk))) +
                              (b->curv_yn[Q2(b, i, j, 2)] * t1)) +
                             (b->curv_yn[Q2(b, i, j, 3)] * t0)));
            }
        for (int j = 1; j <= nj; j++) /* OTA:2590-2606 */
            for (int i = 0; i <= ni; i++) {
                double vel = b->dyte[D2(b, i, j)] * u[D3(b, i, j, k)];
                double upos = 0.5 * (vel + fabs(vel));
                double uneg = 0.5 * (vel - fabs(vel));
                double eastmsk = mq[H3(b, i, j, k)] * (1.0 - mq[H3(b, i - 1, j, k)]);
                double westmsk = mq[H3(b, i + 1, j, k)] * (1.0 - mq[H3(b, i + 2, j, k)]);
                double t0 = tq[H3(b, i, j, k)], t1 = tq[H3(b, i + 1, j, k)];
                flux_x[D3(b, i, j, k)] =
                    ((vel * ((b->quick_x[Q2(b, i, j, 1)] * Tt[D3(b, i, j, k)]) +
                             (b->quick_x[Q2(b, i, j, 2)] * Tt[D3(b, i + 1, j, k)]))) -
                     (upos * (((b->curv_xp[Q2(b, i, j, 1)] * t1) + (b->curv_xp[Q2(b, i, j, 2)] * t0)) +
                              (b->curv_xp[Q2(b, i, j, 3)] * ((tq[H3(b, i - 1, j, k)] * (1.0 - eastmsk)) + (t0 * eastmsk)))))) -
                    (uneg * (((b->curv_xn[Q2(b, i, j, 1)] * ((tq[H3(b, i + 2, j, k)] * (1.0 - westmsk)) + (t1 * westmsk))) +
                              (b->curv_xn[Q2(b, i, j, 2)] * t1)) +
                             (b->curv_xn[Q2(b, i, j, 3)] * t0)));
            }
        if (limit_with_upwind) { /* OTA:2610-2635 */
            for (int j = 1; j <= nj; j++)
                for (int i = 0; i <= ni; i++)
                    if (tmask_limit[D3(b, i, j, k)] == 1.0) {
                        double vel = u[D3(b, i, j, k)];
                        double upos = 0.5 * (vel + fabs(vel));
                        double uneg = 0.5 * (vel - fabs(vel));
                        flux_x[D3(b, i, j, k)] =
                            ((b->dyte[D2(b, i, j)] * ((upos * Tm1[D3(b, i, j, k)]) + (uneg * Tm1[D3(b, i + 1, j, k)]))) *
                             b->tmask[D3(b, i, j, k)]) *
                            b->tmask[D3(b, i + 1, j, k)];
                    }
            for (int j = 0; j <= nj; j++)
                for (int i = 1; i <= ni; i++)
                    if (tmask_limit[D3(b, i, j, k)] == 1.0) {
                        double vel = v[D3(b, i, j, k)];
                        double upos = 0.5 * (vel + fabs(vel));
                        double uneg = 0.5 * (vel - fabs(vel));
                        flux_y[D3(b, i, j, k)] =
                            ((b->dxtn[D2(b, i, j)] * ((upos * Tm1[D3(b, i, j, k)]) + (uneg * Tm1[D3(b, i, j + 1, k)]))) *
                             b->tmask[D3(b, i, j, k)]) *
                            b->tmask[D3(b, i, j + 1, k)];
                    }
        }
    }
}

/* OTA:2640  mpp_update_domains(flux_x, flux_y, Dom_flux, gridtype=CGRID_NE) on a folded-north domain.
 * Result-changing effect on the points the divergence reads: on the fold line j = nj_g the NORTH-position
 * component of the eastern half is replaced by minus its mirror image,
 *     flux_y(i, nj_g) := -flux_y(ni_g+1-i, nj_g)   for i >= middle = (isg+ieg)/2+1
 * (MPPI/mpp_domains_define.inc:1617, 2535-2549: recv overlap isd=max(isd,middle), j=jeg, folded=.true. ->
 *  180-degree rotation flips the sign of non-SCALAR_PAIR vectors, MPPI/mpp_do_updateV.h).  Halo-1 points
 * of flux_x/flux_y also get refreshed from their owners, but the owners compute bit-identical values there
 * (same inputs through consistent halos), so only the fold line changes results.                        */
void orc_fold_fix_flux(const orc_layout *L, double *const *flux_x, double *const *flux_y, int nk)
{
    (void)flux_x;
    if (!L->fold_north) return;
    int middle = (1 + L->ni_g) / 2 + 1;
    int by = L->py - 1;
    int njl = L->jend[by] - L->jbeg[by] + 1;
    for (int ig = middle; ig <= L->ni_g; ig++) {
        int is = L->ni_g + 1 - ig;
        int dx = find_div(L->ibeg, L->iend, L->px, ig), sx = find_div(L->ibeg, L->iend, L->px, is);
        int dni = L->iend[dx] - L->ibeg[dx] + 1, sni = L->iend[sx] - L->ibeg[sx] + 1;
        double *d = flux_y[dx + L->px * by];
        const double *s = flux_y[sx + L->px * by];
        int il = ig - L->ibeg[dx] + 1, isl = is - L->ibeg[sx] + 1;
        for (int k = 0; k < nk; k++)
            d[(size_t)il + (size_t)(dni + 2) * ((size_t)njl + (size_t)(njl + 2) * k)] =
                -s[(size_t)isl + (size_t)(sni + 2) * ((size_t)njl + (size_t)(njl + 2) * k)];
    }
}

/* OTA:2642-2649 (quicker) == OTA:2282-2287 (upwind): tmask*(fx(i)-fx(i-1)+fy(j)-fy(j-1))*datr, negated by caller */
void orc_horz_div(const orc_block *b, const double *fx, const double *fy, double *wrk1_out)
{
    memset(wrk1_out, 0, sizeof(double) * (size_t)NX1(b) * NY1(b) * b->nk); /* OTA:1925-1931 */
    for (int k = 1; k <= b->nk; k++)
        for (int j = 1; j <= b->nj; j++)
            for (int i = 1; i <= b->ni; i++) {
                double r = (b->tmask[D3(b, i, j, k)] *
                            (((fx[D3(b, i, j, k)] - fx[D3(b, i - 1, j, k)]) + fy[D3(b, i, j, k)]) - fy[D3(b, i, j - 1, k)])) *
                           b->datr[D2(b, i, j)];
                wrk1_out[D3(b, i, j, k)] = -r;
            }
}

/* OTA:2981-3031 */
void orc_vert_quicker(const orc_block *b, const double *Tm1, const double *Tt, const double *w,
                      const double *tmask_limit, double *flux_z, double *wrk1_out)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *tmask = b->tmask;
    double *ft1 = (double *)calloc((size_t)ni * nj, sizeof(double));
    memset(wrk1_out, 0, sizeof(double) * (size_t)NX1(b) * NY1(b) * nk); /* OTA:2116-2122 */
    for (int k = 1; k <= nk; k++) {
        int km1 = imax(k - 1, 1), kp1 = imin(k + 1, nk), p2 = imin(k + 2, nk);
        for (int j = 1; j <= nj; j++)
            for (int i = 1; i <= ni; i++) {
                size_t c = (size_t)(i - 1) + (size_t)ni * (j - 1);
                double vel = w[W3(b, i, j, k)];
                double upos = 0.5 * (vel + fabs(vel));
                double uneg = 0.5 * (vel - fabs(vel));
                double ft2;
                if (tmask_limit[D3(b, i, j, k)] == 1.0) {
                    ft2 = (((uneg * Tm1[D3(b, i, j, k)]) + (upos * Tm1[D3(b, i, j, kp1)])) * tmask[D3(b, i, j, k)]) *
                          tmask[D3(b, i, j, kp1)];
                } else {
                    double mp2 = tmask[D3(b, i, j, p2)];
                    int kp2 = (int)lround((mp2 * (double)p2) + ((1.0 - mp2) * (double)kp1)); /* nint() */
                    double upmsk = tmask[D3(b, i, j, k)] * (1.0 - tmask[D3(b, i, j, km1)]);
                    const double *qz = b->quick_z, *zp = b->curv_zp, *zn = b->curv_zn;
                    ft2 = ((vel * ((qz[(k - 1)] * Tt[D3(b, i, j, k)]) + (qz[(k - 1) + nk] * Tt[D3(b, i, j, kp1)]))) -
                           (uneg * (((zp[(k - 1)] * Tm1[D3(b, i, j, kp1)]) + (zp[(k - 1) + nk] * Tm1[D3(b, i, j, k)])) +
                                    (zp[(k - 1) + 2 * nk] * Tm1[D3(b, i, j, km1)])))) -
                          (upos * (((zn[(k - 1)] * Tm1[D3(b, i, j, kp2)]) + (zn[(k - 1) + nk] * Tm1[D3(b, i, j, kp1)])) +
                                   (zn[(k - 1) + 2 * nk] *
                                    ((Tm1[D3(b, i, j, k)] * (1.0 - upmsk)) + (Tm1[D3(b, i, j, kp1)] * upmsk)))));
                }
                if (flux_z) flux_z[D3(b, i, j, k)] = b->dat[D2(b, i, j)] * ft2;
                wrk1_out[D3(b, i, j, k)] = -(tmask[D3(b, i, j, k)] * (ft1[c] - ft2));
                ft1[c] = ft2;
            }
    }
    free(ft1);
}

/* OTA:2238-2294 */
void orc_horz_upwind(const orc_block *b, const double *T, const double *u, const double *v, double *flux_x,
                     double *flux_y, double *wrk1_out)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *tmask = b->tmask;
    for (int k = 1; k <= nk; k++) {
        for (int j = 1; j <= nj; j++)
            for (int i = 0; i <= ni; i++) {
                double velocity = 0.5 * u[D3(b, i, j, k)];
                double upos = velocity + fabs(velocity);
                double uneg = velocity - fabs(velocity);
                flux_x[D3(b, i, j, k)] =
                    ((b->dyte[D2(b, i, j)] * ((upos * T[D3(b, i, j, k)]) + (uneg * T[D3(b, i + 1, j, k)]))) *
                     tmask[D3(b, i, j, k)]) *
                    tmask[D3(b, i + 1, j, k)];
            }
        for (int j = 0; j <= nj; j++)
            for (int i = 1; i <= ni; i++) {
                double velocity = 0.5 * v[D3(b, i, j, k)];
                double upos = velocity + fabs(velocity);
                double uneg = velocity - fabs(velocity);
                flux_y[D3(b, i, j, k)] =
                    ((b->dxtn[D2(b, i, j)] * ((upos * T[D3(b, i, j, k)]) + (uneg * T[D3(b, i, j + 1, k)]))) *
                     tmask[D3(b, i, j, k)]) *
                    tmask[D3(b, i, j + 1, k)];
            }
    }
    orc_horz_div(b, flux_x, flux_y, wrk1_out);
}

/* OTA:2792-2824 */
void orc_vert_upwind(const orc_block *b, const double *T, const double *w, double *flux_z, double *wrk1_out)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    const double *tmask = b->tmask;
    double *ft1 = (double *)calloc((size_t)ni * nj, sizeof(double));
    memset(wrk1_out, 0, sizeof(double) * (size_t)NX1(b) * NY1(b) * nk);
    for (int k = 1; k <= nk; k++) {
        int kp1 = imin(k + 1, nk);
        for (int j = 1; j <= nj; j++)
            for (int i = 1; i <= ni; i++) {
                size_t c = (size_t)(i - 1) + (size_t)ni * (j - 1);
                double velocity = 0.5 * w[W3(b, i, j, k)];
                double wpos = velocity + fabs(velocity);
                double wneg = velocity - fabs(velocity);
                double ft2 = (((wneg * T[D3(b, i, j, k)]) + (wpos * T[D3(b, i, j, kp1)])) * tmask[D3(b, i, j, k)]) *
                             tmask[D3(b, i, j, kp1)];
                if (flux_z) flux_z[D3(b, i, j, k)] = b->dat[D2(b, i, j)] * ft2;
                wrk1_out[D3(b, i, j, k)] = -(tmask[D3(b, i, j, k)] * (ft1[c] - ft2));
                ft1[c] = ft2;
            }
    }
    free(ft1);
}

/* OTA:1990-1996 / 2162-2168 */
void orc_accumulate(const orc_block *b, const double *wrk1, double *th)
{
    for (int k = 1; k <= b->nk; k++)
        for (int j = 1; j <= b->nj; j++)
            for (int i = 1; i <= b->ni; i++) th[D3(b, i, j, k)] = th[D3(b, i, j, k)] + wrk1[D3(b, i, j, k)];
}

/* ocean_tracer.F90:2341-2350 */
void orc_tracer_update(const orc_block *b, double dtime, const double *rho_taum1, const double *rho_dztr_taup1,
                       const double *T_taum1, const double *th, double *T_taup1)
{
    for (int k = 1; k <= b->nk; k++)
        for (int j = 1; j <= b->nj; j++)
            for (int i = 1; i <= b->ni; i++) {
                size_t q = D3(b, i, j, k);
                T_taup1[q] = ((rho_taum1[q] * T_taum1[q]) + (dtime * th[q])) * rho_dztr_taup1[q];
            }
}

/* continuity on the T-grid:
 *   diverge_t(:,:,k) = tmask*(BDX_ET(uhrho_et) + BDY_NT(vhrho_nt))                   ocean_advection_velocity.F90:660
 *   BDX_ET(i,j) = (dyte(i,j)*a(i,j) - dyte(i-1,j)*a(i-1,j))*datr(i,j), 0 at i = isd   ocean_operators.F90:945-958
 *   BDY_NT(i,j) = (dxtn(i,j)*a(i,j) - dxtn(i,j-1)*a(i,j-1))*datr(i,j), 0 at j = jsd   ocean_operators.F90:1230-1243
 *   wrho_bt(:,:,k) = (tmp + diverge_t(:,:,k) + wrho_bt(:,:,k-1))*tmask,  tmp = rho_dzt_tendency - mass_source   :666-669
 * over the whole data domain; wrho_bt(:,:,0) is the caller's (-(pme+river), :664).  tend/src may be NULL (= zero arrays). */
void orc_continuity(const orc_block *b, const double *u, const double *v, const double *tend, const double *src,
                    double *w, double *diverge_t)
{
    const int ni = b->ni, nj = b->nj, nk = b->nk;
    for (int k = 1; k <= nk; k++)
        for (int j = 0; j <= nj + 1; j++)
            for (int i = 0; i <= ni + 1; i++) {
                double bdx = 0.0, bdy = 0.0;
                if (i >= 1)
                    bdx = ((b->dyte[D2(b, i, j)] * u[D3(b, i, j, k)]) - (b->dyte[D2(b, i - 1, j)] * u[D3(b, i - 1, j, k)])) * b->datr[D2(b, i, j)];
                if (j >= 1)
                    bdy = ((b->dxtn[D2(b, i, j)] * v[D3(b, i, j, k)]) - (b->dxtn[D2(b, i, j - 1)] * v[D3(b, i, j - 1, k)])) * b->datr[D2(b, i, j)];
                const double m = b->tmask[D3(b, i, j, k)];
                const double div = m * (bdx + bdy);
                const double tmp = (tend ? tend[D3(b, i, j, k)] : 0.0) - (src ? src[D3(b, i, j, k)] : 0.0);
                if (diverge_t) diverge_t[D3(b, i, j, k)] = div;
                w[W3(b, i, j, k)] = ((tmp + div) + w[W3(b, i, j, k - 1)]) * m;
            }
}

/* MPPI/mpp_chksum_int.h:20-38 + mpp_chksum.h:20-41: wrap-around sum of the int64 bit patterns */
int64_t orc_chksum(const double *a, int ni, int nj, int nk, int halo, const double *mask)
{
    uint64_t s = 0;
    size_t nx = (size_t)ni + 2 * halo, ny = (size_t)nj + 2 * halo;
    for (int k = 0; k < nk; k++)
        for (int j = 0; j < nj; j++)
            for (int i = 0; i < ni; i++) {
                size_t q = (size_t)(i + halo) + nx * ((size_t)(j + halo) + ny * k);
                double v = mask ? a[q] * mask[q] : a[q];
                uint64_t bits;
                memcpy(&bits, &v, 8);
                s += bits;
            }
    return (int64_t)s;
}

/* ocean_tracer_diag.F90:2405-2408, compute domain, conversion = 1 */
double orc_total_tracer(const orc_block *b, const double *rho, const double *T)
{
    double tot = 0.0;
    for (int j = 1; j <= b->nj; j++)
        for (int i = 1; i <= b->ni; i++) {
            double tk = 0.0;
            for (int k = 1; k <= b->nk; k++)
                tk = tk + (((b->tmask[D3(b, i, j, k)] * b->dat[D2(b, i, j)]) * rho[D3(b, i, j, k)]) * T[D3(b, i, j, k)]);
            tot += tk;
        }
    return tot;
}

/* ------------------------------------------------------------------------------------------------
 * multi-block driver = timed CPU baseline. One OpenMP thread per block stands in for one MPI rank;
 * orc_update_halo stands in for mpp_update_domains(XUPDATE / YUPDATE) (OTA:4213-4240, 4302-4343).
 * ---------------------------------------------------------------------------------------------- */
void orc_sweby_all_multiblock(const orc_layout *L, const orc_block *blocks, int ntr, double dtime,
                              const double *const *T, const double *const *u, const double *const *v,
                              const double *const *w, const double *const *rho, double *const *tm,
                              double *const *th, double *const *adv, int nthreads)
{
    int nb = L->px * L->py;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int b = 0; b < nb; b++)
        orc_sweby_all_z(&blocks[b], ntr, dtime, T + (size_t)b * ntr, w[b], rho[b], tm + (size_t)b * ntr, NULL, NULL);

    double **f = (double **)malloc(sizeof(double *) * nb);
    for (int n = 0; n < ntr; n++) {
        for (int b = 0; b < nb; b++) f[b] = tm[(size_t)b * ntr + n];
        orc_update_halo(L, f, blocks[0].nk, 2, ORC_XUPDATE);
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int b = 0; b < nb; b++)
        orc_sweby_all_x(&blocks[b], ntr, dtime, T + (size_t)b * ntr, u[b], rho[b], tm + (size_t)b * ntr, NULL, NULL);

    for (int n = 0; n < ntr; n++) {
        for (int b = 0; b < nb; b++) f[b] = tm[(size_t)b * ntr + n];
        orc_update_halo(L, f, blocks[0].nk, 2, ORC_YUPDATE);
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int b = 0; b < nb; b++)
        orc_sweby_all_y(&blocks[b], ntr, dtime, T + (size_t)b * ntr, u[b], v[b], w[b], rho[b], tm + (size_t)b * ntr,
                        th + (size_t)b * ntr, adv + (size_t)b * ntr, NULL, NULL);
    free(f);
}
