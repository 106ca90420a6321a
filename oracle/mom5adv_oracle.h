/*
 * mom5adv_oracle.h -- CPU ORACLE for the MOM5 tracer-advection hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (mom5_b200/, libmom5adv.so) never links, imports or calls anything in oracle/.
 *
 * It restates, operation for operation, the arithmetic of
 *   /root/reference/src/mom5/ocean_tracers/ocean_tracer_advect.F90   ("OTA")
 * for advect_tracer_sweby_all (OTA:4104-4511), advect_tracer_mdfl_sweby (OTA:3806-4066),
 * advect_tracer_mdfl_sweby_test (OTA:3469-3746),
 * horz/vert_advect_tracer_quicker (OTA:2538-2653, 2981-3031) + quicker_init (OTA:1442-1586),
 * horz/vert_advect_tracer_upwind (OTA:2238-2294, 2792-2824), and the FMS halo-update
 * semantics they rely on (src/shared/mpp/include/mpp_do_update.h:57-78,
 * mpp_domains_define.inc:4865-4885, test_mpp_domains.F90:3749-3766).
 *
 * Parity pinning: the reference ships no unit-level golden vectors for this path and its
 * Fortran cannot be compiled in this image (no Fortran compiler).  The oracle is pinned
 * instead against tests/golden/ *.npz fixtures produced by tests/golden/gen_from_reference.py,
 * which EXECUTES THE REFERENCE'S OWN SOURCE TEXT (the loop nests read from OTA at
 * generation time) through a small Fortran-statement interpreter in IEEE binary64.
 * See DESIGN.md "Oracle".
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (NO -march enabling FMA) -- see Makefile.
 *
 * Conventions
 *   - all reals are IEEE binary64 (reference builds with -fdefault-real-8 / -r8);
 *   - arrays are Fortran column-major, i fastest;
 *   - a "block" is one rank's local domain: compute domain 1..ni x 1..nj (isc=jsc=1),
 *     data domain (halo 1) 0..ni+1, "h2" scratch (halo 2) -1..ni+2, k = 1..nk;
 *   - wrho_bt is dimensioned (0:ni+1, 0:nj+1, 0:nk) as in ocean_advection_velocity.F90:342.
 *   - max(a,b)/min(a,b): first argument wins ties (see orc_max/orc_min).
 */
#ifndef MOM5ADV_ORACLE_H
#define MOM5ADV_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_block {
    int ni, nj, nk;
    /* data-domain (halo 1) static fields, dims (ni+2, nj+2) */
    const double *dat, *datr, *dxt, *dyt, *dxte, *dyte, *dxtn, *dytn;
    const double *tmask;      /* (ni+2, nj+2, nk)  Grd%tmask */
    const double *dzt;        /* (nk)              Grd%dzt   */
    /* halo-2 static fields, dims (ni+4, nj+4[, nk]); filled by *_init + halo update */
    double *tmask_h2;         /* tmask_mdfl == tmask_quick (OTA:1668-1675, 1478-1487) */
    double *dxt_h2, *dyt_h2;  /* dxt_quick, dyt_quick (OTA:1478-1511) */
    /* quicker weights on the data domain (OTA:1447-1455) */
    double *quick_x, *quick_y;                      /* (ni+2, nj+2, 2) */
    double *curv_xp, *curv_xn, *curv_yp, *curv_yn;  /* (ni+2, nj+2, 3) */
    double *quick_z, *curv_zp, *curv_zn;            /* (nk,2) (nk,3) (nk,3) */
} orc_block;

/* global decomposition descriptor for the halo "exchange" among in-process blocks */
typedef struct orc_layout {
    int ni_g, nj_g;           /* global compute extents */
    int px, py;               /* layout (idiv, jdiv) */
    const int *ibeg, *iend;   /* [px] global start/end (1-based, inclusive) per x-division */
    const int *jbeg, *jend;   /* [py] */
    int cyclic_x, cyclic_y, fold_north;
} orc_layout;

enum { ORC_XUPDATE = 1, ORC_YUPDATE = 2 };

/* ---- layout (mpp_domains_define.inc:28-55, 187-273) ---- */
void orc_define_layout(int ni_g, int nj_g, int ndivs, int *layout2);
int  orc_compute_extent(int isg, int ieg, int ndivs, int *ibegin, int *iend);

/* ---- halo update of halo-`halo` fields among blocks (one field per block) ----
 * fields[b] for block b = ix + px*iy, dims (ni_b+2*halo, nj_b+2*halo, nk).
 * flags: ORC_XUPDATE (E/W, j in compute), ORC_YUPDATE (N/S, i in compute), both = full incl. corners.
 * Halo points with no image (solid wall) are left untouched.                                        */
void orc_update_halo(const orc_layout *L, double *const *fields, int nk, int halo, int flags);

/* ---- sweby_all (OTA:4104-4511), per block, per phase ---- */
void orc_mdfl_init_mask(const orc_block *b);   /* tmask_h2 := 0; compute domain := Grd%tmask (OTA:1660-1674) */

void orc_sweby_all_z(const orc_block *b, int ntr, double dtime,
                     const double *const *T, const double *wrho_bt, const double *rho_dzt,
                     double *const *tm, double *const *flux_z, double *const *adv_z);
void orc_sweby_all_x(const orc_block *b, int ntr, double dtime,
                     const double *const *T, const double *uhrho_et, const double *rho_dzt,
                     double *const *tm, double *const *flux_x, double *const *adv_x);
void orc_sweby_all_y(const orc_block *b, int ntr, double dtime,
                     const double *const *T, const double *uhrho_et, const double *vhrho_nt,
                     const double *wrho_bt, const double *rho_dzt,
                     double *const *tm, double *const *th_tendency, double *const *adv_tendency,
                     double *const *flux_y, double *const *adv_y);

/* ---- advect_tracer_mdfl_sweby (OTA:3806-4066), one tracer; out = Tracer%wrk1 (caller's negation applied) ---- */
void orc_mdfl_sweby_z(const orc_block *b, double dtime, double sweby_limiter,
                      const double *T, const double *wrho_bt, const double *rho_dzt,
                      double *tm, double *flux_z);
void orc_mdfl_sweby_x(const orc_block *b, double dtime, double sweby_limiter,
                      const double *T, const double *uhrho_et, const double *rho_dzt,
                      double *tm, double *flux_x);
void orc_mdfl_sweby_y(const orc_block *b, double dtime, double sweby_limiter,
                      const double *T, const double *uhrho_et, const double *vhrho_nt,
                      const double *wrho_bt, const double *rho_dzt,
                      double *tm, double *flux_y, double *wrk1_out);

/* ---- advect_tracer_mdfl_sweby_test (OTA:3469-3746), one tracer; tr/tms/ms = tracer_mdfl, tracermass_mdfl, mass_mdfl
 *      (h2 scratch; XUPDATE of all three between z and x, YUPDATE between x and y) ---- */
void orc_sweby_test_z(const orc_block *b, double dtime, double sweby_limiter, const double *T, const double *wrho_bt,
                      const double *rho_dzt, double *tr, double *tms, double *ms, double *flux_z);
void orc_sweby_test_x(const orc_block *b, double dtime, double sweby_limiter, const double *uhrho_et,
                      double *tr, double *tms, double *ms, double *flux_x);
void orc_sweby_test_y(const orc_block *b, double dtime, double sweby_limiter, const double *T, const double *vhrho_nt,
                      const double *rho_dzt, double *tr, double *tms, double *ms, double *flux_y, double *wrk1_out);

/* ---- advect_tracer_mdppm (OTA:5990-6494) + ppm_limit_cw84/_ifc/_sh (OTA:6510-6657), one tracer.  m4 = tmask_mdppm and
 *      tr = tracer_mdppm on the halo-4 scratch, dims (ni+8, nj+8, nk); XUPDATE of tr (halo 4) between z and x, YUPDATE
 *      between x and y.  limiter = Tracer%ppm_hlimiter (1 cw84, 2 ifc, 3 sh; the reference uses it in all directions). ---- */
void orc_mdppm_z(const orc_block *b, double dtime, int limiter, const double *T, const double *wrho_bt, const double *rho_dzt,
                 const double *m4, double *tr, double *flux_z);
void orc_mdppm_x(const orc_block *b, double dtime, int limiter, const double *T, const double *uhrho_et, const double *rho_dzt,
                 const double *m4, double *tr, double *flux_x);
void orc_mdppm_y(const orc_block *b, double dtime, int limiter, const double *T, const double *uhrho_et, const double *vhrho_nt,
                 const double *wrho_bt, const double *rho_dzt, const double *m4, double *tr, double *flux_y, double *wrk1_out);

/* ---- compute_adv_diss (OTA:7547-7712): element-wise parts; the operators on T**2 are the arms above ---- */
void orc_square(const orc_block *b, const double *T, double *out);
void orc_adv_diss_final(const orc_block *b, double dtime, double conversion, const double *rho_tau, const double *rho_taup1,
                        const double *T_tau, const double *advect_tendency, const double *wrk2, const double *wrk3,
                        double *t2_tendency, double *diss);

/* ---- quicker (OTA:1442-1586, 2538-2653, 2981-3031) ---- */
void orc_quicker_init_pre(const orc_block *b);    /* fills tmask_h2/dxt_h2/dyt_h2 before their halo updates  */
void orc_quicker_init_edges(const orc_block *b);  /* OTA:1490-1509 edge replication (after tmask update)    */
void orc_quicker_init_weights(const orc_block *b);/* OTA:1516-1578 */
void orc_quicker_prep(const orc_block *b, const double *T_taum1, double *tq); /* OTA:2558-2565 */
void orc_horz_quicker_flux(const orc_block *b, const double *T_taum1, const double *T_tau, const double *tq,
                           const double *uhrho_et, const double *vhrho_nt,
                           const double *tmask_limit, int limit_with_upwind,
                           double *flux_x, double *flux_y);
void orc_fold_fix_flux(const orc_layout *L, double *const *flux_x, double *const *flux_y, int nk); /* OTA:2640 */
void orc_horz_div(const orc_block *b, const double *flux_x, const double *flux_y, double *wrk1_out); /* OTA:2642-2649, negated */
void orc_vert_quicker(const orc_block *b, const double *T_taum1, const double *T_tau, const double *wrho_bt,
                      const double *tmask_limit, double *flux_z, double *wrk1_out);

/* ---- upwind (OTA:2238-2294, 2792-2824) ---- */
void orc_horz_upwind(const orc_block *b, const double *T, const double *uhrho_et, const double *vhrho_nt,
                     double *flux_x, double *flux_y, double *wrk1_out);
void orc_vert_upwind(const orc_block *b, const double *T, const double *wrho_bt, double *flux_z, double *wrk1_out);

/* ---- dispatcher tail: th_tendency += wrk1 on the compute domain (OTA:1990-1996, 2162-2168) ---- */
void orc_accumulate(const orc_block *b, const double *wrk1, double *th_tendency);

/* ---- consumer (ocean_tracer.F90:2341-2350): field(taup1) = (rho_dzt(taum1)*T + dtime*th)*rho_dztr ---- */
void orc_tracer_update(const orc_block *b, double dtime, const double *rho_dzt_taum1, const double *rho_dztr_taup1,
                       const double *T_taum1, const double *th_tendency, double *T_taup1);

/* ---- continuity: diverge_t and wrho_bt from the horizontal transports (ocean_advection_velocity.F90:660-669) ---- */
void orc_continuity(const orc_block *b, const double *uhrho_et, const double *vhrho_nt, const double *rho_dzt_tendency,
                    const double *mass_source, double *wrho_bt, double *diverge_t);

/* ---- metrics ---- */
int64_t orc_chksum(const double *a, int ni, int nj, int nk, int halo, const double *mask_or_null);
    /* mpp_chksum_int.h:20-38 on the compute domain; with mask: chksum(a*mask) (ocean_tracer_util.F90:562-566) */
double  orc_total_tracer(const orc_block *b, const double *rho_dzt, const double *T);
    /* ocean_tracer_diag.F90:2405-2408 restricted to the compute domain (this block's share of mpp_global_sum) */

/* ---- multi-block drivers: whole sweby_all over px*py in-process blocks, OpenMP over blocks.
 * This is the timed CPU baseline ("blocks stand in for MPI ranks").                                        */
void orc_sweby_all_multiblock(const orc_layout *L, const orc_block *blocks, int ntr, double dtime,
                              const double *const *T,        /* [nb*ntr]  T[b*ntr+n] */
                              const double *const *u, const double *const *v, const double *const *w,
                              const double *const *rho,      /* [nb] */
                              double *const *tm,             /* [nb*ntr] h2 scratch */
                              double *const *th, double *const *adv, /* [nb*ntr] */
                              int nthreads);

#ifdef __cplusplus
}
#endif
#endif
