/*
 * mom5adv.h -- C ABI of the B200-native MOM5 tracer-advection path (libmom5adv.so).
 *
 * The reference (MOM5, Fortran 90) has no FFI boundary on this path; the boundary is created at the
 * call sites of ocean_tracer_advect_mod, cited per entry point below
 * (OTA = src/mom5/ocean_tracers/ocean_tracer_advect.F90 in the reference tree).  The Fortran side binds
 * these symbols through ISO_C_BINDING.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MOM5ADV_E* code on error; mom5adv_last_error()
 *     returns a message.  Nothing aborts: the Fortran shim maps nonzero to mpp_error(FATAL, ...).
 *   - all reals are IEEE binary64; arrays are Fortran column-major exactly as the reference dimensions them:
 *       2-D  (isd:ied, jsd:jed)            3-D  (isd:ied, jsd:jed, nk)
 *       wrho_bt (isd:ied, jsd:jed, 0:nk)   -- pass the address of the whole array (element (isd,jsd,0))
 *     with isd = isc-1, ied = iec+1, jsd = jsc-1, jed = jec+1 (ocean_domains_nml halo = 1, ocean_domains.F90:77).
 *   - "host" entry points take host pointers and are synchronous on return (H2D, kernels, D2H inside);
 *     "_dev" entry points take DEVICE pointers of the same layout plus a cudaStream_t (as void*), enqueue
 *     the work and return without synchronising.  The benchmarked path is the _dev one.
 *   - one handle per rank/GPU; a handle is not thread safe.
 *   - results are bit-identical to the reference arithmetic when the library is built with FMA
 *     contraction disabled (the default build: nvcc --fmad=false).
 */
#ifndef MOM5ADV_H
#define MOM5ADV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOM5ADV_VERSION 200

/* error codes */
#define MOM5ADV_OK 0
#define MOM5ADV_EINVAL (-1)   /* bad argument */
#define MOM5ADV_ECUDA (-2)    /* CUDA runtime error */
#define MOM5ADV_ENCCL (-3)    /* NCCL error */
#define MOM5ADV_ENOGPU (-4)   /* no CUDA device */
#define MOM5ADV_EUNSUP (-5)   /* configuration the path does not cover (e.g. open boundaries) */

/* scheme ids == ocean_parameters.F90:149-163 */
#define MOM5ADV_ADVECT_UPWIND 1
#define MOM5ADV_ADVECT_QUICKER 5
#define MOM5ADV_ADVECT_MDPPM 8
#define MOM5ADV_ADVECT_MDFL_SWEBY 9
#define MOM5ADV_ADVECT_DST_LINEAR 10
#define MOM5ADV_ADVECT_MDFL_SWEBY_TEST 12
#define MOM5ADV_ADVECT_DST_LINEAR_TEST 14

typedef struct mom5adv_ctx *mom5adv_handle;
typedef struct mom5adv_comm_s *mom5adv_comm; /* wraps an ncclComm_t; NULL = single rank */

/* Filled by the caller (the ISO_C_BINDING shim, from Grd/Dom) once.  All pointers are HOST pointers and are
 * only read during mom5adv_init.  Replaces the module state set up by ocean_tracer_advect_init
 * (OTA:507-756), mdfl_init (OTA:1644-1691) and quicker_init (OTA:1442-1586).                            */
typedef struct mom5adv_grid {
    int isc, iec, jsc, jec;       /* compute domain, GLOBAL indices (ocean_domains.F90:311-318) */
    int nk;
    int ni_global, nj_global;
    int layout_x, layout_y;       /* ocean_model_nml layout */
    const int *x_extent;          /* [layout_x] compute extents per x-division, or NULL = mpp_compute_extent rule */
    const int *y_extent;          /* [layout_y] */
    int cyclic_x, cyclic_y, tripolar; /* ocean_types.F90:751-753 */
    int have_obc;                 /* must be 0: open boundaries are not covered (-> MOM5ADV_EUNSUP) */
    const double *dat, *datr, *dxt, *dyt, *dxte, *dyte, *dxtn, *dytn; /* (isd:ied, jsd:jed) */
    const double *dzt;            /* (nk) Grd%dzt, for the quicker vertical weights */
    const double *tmask;          /* (isd:ied, jsd:jed, nk) Grd%tmask */
} mom5adv_grid;

const char *mom5adv_last_error(void);
int mom5adv_version(void);

/* ---- communicator (only needed when layout_x*layout_y > 1) ------------------------------------------
 * Replaces FMS mpp_update_domains for this path only.  Either let the library create its NCCL communicator
 * (rank 0 calls mom5adv_comm_unique_id, the caller broadcasts the 128 bytes -- MPI_Bcast in the Fortran
 * model, torch.distributed in bench.py -- then every rank calls mom5adv_comm_create), or wrap an existing
 * ncclComm_t.  Rank r owns block (r % layout_x, r / layout_x), the FMS pelist order.                     */
int mom5adv_comm_unique_id(char id_out[128]);
int mom5adv_comm_create(const char id[128], int rank, int nranks, mom5adv_comm *out);
int mom5adv_comm_from_nccl(void *nccl_comm, int rank, int nranks, mom5adv_comm *out);
int mom5adv_comm_destroy(mom5adv_comm c);

/* ---- device selection -----------------------------------------------------------------------------------
 * One rank drives one GPU.  A multi-rank host (the MPI Fortran model) calls mom5adv_set_device with its NODE-LOCAL rank
 * before mom5adv_comm_create / mom5adv_init: the rank is bound to device (rank mod device count), as FMS binds ranks to
 * cores (src/shared/mpp/affinity.c:67).  Without it every rank would land on device 0 and ncclCommInitRank would refuse
 * the duplicate GPU.  A host that already selected a device (torch.cuda.set_device) does not need to call it.        */
int mom5adv_device_count(int *count);
int mom5adv_set_device(int node_local_rank);

/* ---- lifetime ---------------------------------------------------------------------------------------- */
int mom5adv_init(const mom5adv_grid *grid, int ntracers_max, mom5adv_comm comm_or_null, mom5adv_handle *out);
int mom5adv_finalize(mom5adv_handle h);

/* ---- advect_tracer_sweby_all (OTA:4104-4511; called from horz_advect_tracer, OTA:2078-2080) -----------
 * For n = 0..ntr-1:   adv_tendency[n] (= T_prog(n)%wrk1) := rho_dzt*(tm - T)/dtime*tmask on the compute
 * domain, 0 on the halo ring;  th_tendency[n] += adv_tendency[n] on the compute domain.
 * Optional per-tracer diagnostics (array of ntr pointers, or NULL; individual entries may be NULL):
 *   flux_x (i=isc-1..iec), flux_y (j=jsc-1..jec), flux_z, and the per-direction tendencies adv_x/y/z
 *   (the reference's shared wrk1 sent to diag ids *_advection_x/y/z); points outside the reference's loop
 *   ranges are left untouched.                                                                           */
/* Host entry point: th_tendency may be NULL (single-rank layouts, no diagnostics requested): the call then returns adv_tendency
 * only and the caller forms th_tendency += adv_tendency itself.  When th_tendency is given, the library forms the sum ON THE
 * HOST, band by band while the next bands are on the link -- th_tendency never crosses PCIe (one IEEE add per point: the same
 * bits as on the device).  Caller arrays are page-locked on first use (cudaHostRegister, cached by address; MOM5ADV_PIN=0
 * disables it) so that the copies really overlap.                                                                        */
int mom5adv_sweby_all(mom5adv_handle h, int ntr, double dtime,
                      const double *const *T_taum1, double *const *th_tendency, double *const *adv_tendency,
                      const double *uhrho_et, const double *vhrho_nt, const double *wrho_bt,
                      const double *rho_dzt_tau,
                      double *const *flux_x, double *const *flux_y, double *const *flux_z,
                      double *const *adv_x, double *const *adv_y, double *const *adv_z);
int mom5adv_sweby_all_dev(mom5adv_handle h, int ntr, double dtime,
                          const double *const *T_taum1, double *const *th_tendency, double *const *adv_tendency,
                          const double *uhrho_et, const double *vhrho_nt, const double *wrho_bt,
                          const double *rho_dzt_tau,
                          double *const *flux_x, double *const *flux_y, double *const *flux_z,
                          double *const *adv_x, double *const *adv_y, double *const *adv_z, void *stream);

/* ---- horz_advect_tracer (OTA:1898-2083), one tracer, advect_sweby_all = .false. -----------------------
 * scheme: UPWIND (OTA:2238-2294), QUICKER (OTA:2538-2653), MDFL_SWEBY / DST_LINEAR (OTA:3806-4066),
 * MDFL_SWEBY_TEST / DST_LINEAR_TEST (OTA:3469-3746), MDPPM (OTA:5990-6494; limiter via mom5adv_set_ppm_limiters).
 * wrk1_out (= Tracer%wrk1) := -scheme(...) on the compute domain, 0 on the halo ring;
 * th_tendency += wrk1_out on the compute domain (OTA:1990-1996).  flux_x/flux_y/flux_z may be NULL.
 * T_tau and tmask_limit are read by QUICKER only (tmask_limit only if limit_with_upwind != 0).           */
int mom5adv_horz(mom5adv_handle h, int scheme, double dtime,
                 const double *T_taum1, const double *T_tau, const double *tmask_limit, int limit_with_upwind,
                 const double *uhrho_et, const double *vhrho_nt, const double *wrho_bt, const double *rho_dzt_tau,
                 double *th_tendency, double *wrk1_out, double *flux_x, double *flux_y, double *flux_z);
int mom5adv_horz_dev(mom5adv_handle h, int scheme, double dtime,
                     const double *T_taum1, const double *T_tau, const double *tmask_limit, int limit_with_upwind,
                     const double *uhrho_et, const double *vhrho_nt, const double *wrho_bt, const double *rho_dzt_tau,
                     double *th_tendency, double *wrk1_out, double *flux_x, double *flux_y, double *flux_z,
                     void *stream);

/* ---- vert_advect_tracer (OTA:2095-2226), one tracer ----------------------------------------------------
 * scheme: UPWIND (OTA:2792-2824), QUICKER (OTA:2981-3031); MDFL_SWEBY / DST_LINEAR are three-dimensional:
 * wrk1_out := 0 and th_tendency is unchanged (OTA:2147-2155).                                            */
int mom5adv_vert(mom5adv_handle h, int scheme,
                 const double *T_taum1, const double *T_tau, const double *tmask_limit,
                 const double *wrho_bt, double *th_tendency, double *wrk1_out, double *flux_z);
int mom5adv_vert_dev(mom5adv_handle h, int scheme,
                     const double *T_taum1, const double *T_tau, const double *tmask_limit,
                     const double *wrho_bt, double *th_tendency, double *wrk1_out, double *flux_z, void *stream);

/* ---- consumer of the tendencies -------------------------------------------------
 * T_taup1[n] := (rho_dzt_taum1*T_taum1[n] + dtime*th_tendency[n])*rho_dztr_taup1 on the compute domain
 * (ocean_tracer.F90:2341-2350), followed by the halo-1 update of T_taup1[n] that update_ocean_model performs
 * (mpp_update_domains(T_prog(n)%field(:,:,:,taup1), Dom%domain2d), ocean_model.F90:1903-1911): interior neighbours,
 * cyclic wrap, folded north edge; wall halos are left untouched.  Device pointers, data-domain layout.        */
int mom5adv_tracer_update_dev(mom5adv_handle h, int ntr, double dtime, const double *rho_dzt_taum1,
                              const double *rho_dztr_taup1, const double *const *T_taum1,
                              const double *const *th_tendency, double *const *T_taup1, void *stream);

/* The same consumer FUSED onto the advection pass: update_advection_only (ocean_tracer.F90:2618-2649) with
 * advect_tracer_sweby_all as the operator.  th_tendency := 0 + T_prog(n)%wrk1 and
 * T_taup1[n] := (rho_dzt_taum1*T_taum1[n] + dtime*th_tendency)*rho_dztr_taup1 are formed in the epilogue of the x/y pass,
 * followed by the halo-1 update of T_taup1[n].  th_out / adv_out (arrays of ntr pointers, or NULL, or NULL entries) receive
 * th_tendency / wrk1 on the compute domain when wanted; leaving them NULL saves their HBM traffic.                     */
int mom5adv_sweby_all_step_dev(mom5adv_handle h, int ntr, double dtime, const double *const *T_taum1,
                               const double *rho_dzt_taum1, const double *rho_dztr_taup1,
                               const double *uhrho_et, const double *vhrho_nt, const double *wrho_bt, const double *rho_dzt_tau,
                               double *const *T_taup1, double *const *th_out, double *const *adv_out, void *stream);

/* ---- producer of wrho_bt ----------------------------------------------------------
 * diverge_t(:,:,k) = tmask*(BDX_ET(uhrho_et) + BDY_NT(vhrho_nt)) and the continuity recurrence
 * wrho_bt(:,:,k) = ((rho_dzt_tendency - mass_source) + diverge_t(:,:,k) + wrho_bt(:,:,k-1))*tmask over the whole data
 * domain (ocean_advection_velocity.F90:660-669; ocean_operators.F90:945-958, 1230-1243).  wrho_bt(:,:,0) = -(pme+river)
 * must be set by the caller; rho_dzt_tendency / mass_source / diverge_t may be NULL (zero arrays / not wanted).     */
int mom5adv_continuity_dev(mom5adv_handle h, const double *uhrho_et, const double *vhrho_nt,
                           const double *rho_dzt_tendency, const double *mass_source, double *wrho_bt,
                           double *diverge_t, void *stream);

/* Tracer%ppm_hlimiter / Tracer%ppm_vlimiter of the tracer the next ADVECT_MDPPM calls advect (field-table entries,
 * ocean_tracer.F90:1006,1051; defaults 1): 1 = Colella-Woodward 1984, 2 = Lin's improved full constraint, 3 = Suresh-Huynh
 * 1997 (ppm_limit_cw84 / _ifc / _sh, OTA:6510-6657).  advect_tracer_mdppm reads ppm_hlimiter in all three directions
 * (OTA:6137,6263,6399); any other value makes the scheme fail as the reference does (OTA:6146-6148).                 */
int mom5adv_set_ppm_limiters(mom5adv_handle h, int ppm_hlimiter, int ppm_vlimiter);

/* ---- diagnostics producers ------------------------------------------------------
 * mom5adv_adv_diss_dev: compute_adv_diss (OTA:7547-7712), called from vert_advect_tracer's tail (OTA:2221-2223): the
 * tracer's own horizontal and vertical advection operators applied to field(tau)**2, then
 * adv_diss = -(conversion**2)/dtime * (adv*(2*rho_dzt(tau)*T + dtime*adv) - rho_dzt(taup1)*t2_tendency) on the compute
 * domain, 0 elsewhere.  advect_tendency = horz + vert Tracer%wrk1 (OTA:2000-2006, 2170-2176).  t2_tendency (the
 * reference's wrk1, diagnostic id_tracer2_advection before its conversion**2 factor) may be NULL.  Schemes as mom5adv_horz /
 * mom5adv_vert; ADVECT_MDFL_SWEBY_TEST has no arm in the reference's select, its horizontal operator is 0.
 * mom5adv_flux_int_z_dev: the z-integrated flux diagnostics *_xflux_adv_int_z / *_yflux_adv_int_z (OTA:4317-4326,
 * 4449-4458): out2d(i,j) = sum over k (in order) of flux3d(i,j,k) on the compute domain, 0 elsewhere.              */
int mom5adv_adv_diss_dev(mom5adv_handle h, int horz_scheme, int vert_scheme, double dtime, double conversion,
                         const double *T_tau, const double *tmask_limit, int limit_with_upwind,
                         const double *uhrho_et, const double *vhrho_nt, const double *wrho_bt,
                         const double *rho_dzt_tau, const double *rho_dzt_taup1, const double *advect_tendency,
                         double *adv_diss, double *t2_tendency, void *stream);
int mom5adv_adv_diss(mom5adv_handle h, int horz_scheme, int vert_scheme, double dtime, double conversion,
                     const double *T_tau, const double *tmask_limit, int limit_with_upwind,
                     const double *uhrho_et, const double *vhrho_nt, const double *wrho_bt,
                     const double *rho_dzt_tau, const double *rho_dzt_taup1, const double *advect_tendency,
                     double *adv_diss, double *t2_tendency);   /* host-pointer twin, synchronous */
int mom5adv_flux_int_z_dev(mom5adv_handle h, const double *flux3d, double *out2d, void *stream);

/* ---- metrics on device arrays ---------------------------------------------------------------------------
 * mom5adv_chksum_dev: mpp_chksum of the compute domain (wrap-around sum of the int64 bit patterns,
 * mpp_chksum_int.h:20-38), this rank's share; masked != 0 multiplies by tmask first
 * (ocean_tracer_util.F90:562-566).  mom5adv_total_tracer_dev: sum tmask*dat*rho_dzt*T over the compute
 * domain (ocean_tracer_diag.F90:2405-2408), this rank's share.                                           */
int mom5adv_chksum_dev(mom5adv_handle h, const double *field3d, int masked, int64_t *out, void *stream);
int mom5adv_total_tracer_dev(mom5adv_handle h, const double *rho_dzt, const double *T, double *out, void *stream);

/* ---- introspection ------------------------------------------------------------------------------------- */
/* device time of the last *_dev / host call per bucket, matching the reference's clocks
 * (OTA:1678-1688): [0] z sweep (cuk)  [1] x sweep (cui)  [2] y sweep (cuj)  [3] halo/mpi  [4] total.
 * Synchronises the handle's streams.                                                                     */
int mom5adv_last_timing_ms(mom5adv_handle h, float ms[5]);
/* number of kernels this library launched since init (all streams) */
int64_t mom5adv_kernel_launches(mom5adv_handle h);
/* bytes the last host-pointer call moved over the link: [0] host->device, [1] device->host (counted copy by copy) */
int mom5adv_last_transfer_bytes(mom5adv_handle h, int64_t bytes[2]);

#ifdef __cplusplus
}
#endif
#endif
