// capi.cu -- extern "C" entry points of libmom5adv.so (see include/mom5adv.h) and the host-side driver:
// domain bookkeeping, halo plans, launch configuration, stream/event handling.
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is dlopen'ed on first use (see nccl_api below)
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <set>
#include <thread>
#include <tuple>
#include <cmath>

#include "../../include/mom5adv.h"
#include "../../include/mpp_layout.h"
#include "halo.cuh"
#include "horz_vert_kernels.cuh"
#include "mom5adv_internal.cuh"
#include "sweby_kernels.cuh"
#include "sweby_fused.cuh"
#include "sweby_z_tma.cuh"
#include "sweby_fused_tma.cuh"
#include "sweby_test_kernels.cuh"
#include "mdppm_kernels.cuh"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
extern "C" const char *mom5adv_last_error(void) { return g_err; }
extern "C" int mom5adv_version(void) { return MOM5ADV_VERSION; }

// NCCL is bound at run time, not at link time: a host process that already carries an NCCL (PyTorch bundles its
// own libnccl.so.2) must keep exactly that one -- a DT_NEEDED on the system copy would shadow it when this
// library happens to be loaded first.  dlopen("libnccl.so.2") returns the already-loaded image if there is one.
struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId;
    decltype(&ncclCommInitRank) CommInitRank;
    decltype(&ncclCommDestroy) CommDestroy;
    decltype(&ncclGetErrorString) GetErrorString;
    decltype(&ncclGroupStart) GroupStart;
    decltype(&ncclGroupEnd) GroupEnd;
    decltype(&ncclSend) Send;
    decltype(&ncclRecv) Recv;
};
static NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        for (const char *nm : {"libnccl.so.2", "libnccl.so"}) {
            if ((api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
        }
        if (api.lib) {
#define NCCL_SYM(f) api.f = (decltype(api.f))dlsym(api.lib, "nccl" #f); if (!api.f) api.lib = nullptr
            NCCL_SYM(GetUniqueId); NCCL_SYM(CommInitRank); NCCL_SYM(CommDestroy); NCCL_SYM(GetErrorString);
            NCCL_SYM(GroupStart); NCCL_SYM(GroupEnd); NCCL_SYM(Send); NCCL_SYM(Recv);
#undef NCCL_SYM
        }
    }
    return api.lib ? &api : nullptr;
}
#define NCCL_NEED()                                                                                   \
    NcclApi *N = nccl_api();                                                                          \
    if (!N) { set_error("libnccl.so.2 could not be loaded: %s", dlerror()); return MOM5ADV_ENCCL; }

#define NCCL_TRY(x)                                                                                   \
    do {                                                                                              \
        ncclResult_t r_ = (x);                                                                        \
        if (r_ != ncclSuccess) {                                                                      \
            set_error("NCCL error %s at %s:%d (%s)", N->GetErrorString(r_), __FILE__, __LINE__, #x);  \
            return MOM5ADV_ENCCL;                                                                     \
        }                                                                                             \
    } while (0)

// ------------------------------------------------------------------------------------------------
// tensor maps (TMA descriptors).  The encoder lives in the driver library; it is fetched through the runtime so that the
// library carries no link-time dependency on libcuda.
// ------------------------------------------------------------------------------------------------
struct TmapKey {
    const void *base;
    unsigned long long n0, n1, n2;
    unsigned b0, b1, b2;
    bool operator<(const TmapKey &o) const
    {
        return std::tie(base, n0, n1, n2, b0, b1, b2) < std::tie(o.base, o.n0, o.n1, o.n2, o.b0, o.b1, o.b2);
    }
};
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tma_encoder()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct mom5adv_ctx {
    Geom g;
    int isc_g, jsc_g, ni_g, nj_g, px, py, ix, iy, rank;
    int cyclic_x, cyclic_y, tripolar;
    std::vector<int> ibeg, iend, jbeg, jend;
    mom5adv_comm comm = nullptr;
    int ntr_max = 0;
    // static device data
    double *dat = 0, *datr = 0, *dxte = 0, *dyte = 0, *dxtn = 0, *dytn = 0, *tmask = 0;
    uint8_t *mask = 0;                 // tmask_mdfl == tmask_quick as u8, halo 2
    uint8_t *nibx = 0, *niby = 0;      // neighbourhood nibbles of the x and y sweeps, data-domain layout
    uint8_t *nibx_p = 0, *niby_p = 0;  // the same with rows padded to nib_pitch bytes (a multiple of 16) for the tensor maps of the fused pass
    int nib_pitch = 0;
    QuickW qw{};                       // quicker weights (device)
    std::vector<double *> tmA, tmB;    // h2 scratch per tracer
    double *st_ms = 0;                 // mass_mdfl of advect_tracer_mdfl_sweby_test (h2 scratch, allocated on first use)
    double *pp_fz = 0;                 // MDPPM: flux_z work array when the caller does not want it
    double *pp_tr = 0, *pp_m4 = 0, *pp_da = 0;   // MDPPM: tracer_mdppm, tmask_mdppm, slope scratch (h4, allocated on first use)
    bool pp_ready = false;             // tmask_mdppm filled and halo-updated
    HaloPlan plan4[4];                 // halo-4 updates (Dom_mdppm, OTA:1706), indexed by flags
    int ppm_hlimiter = 1, ppm_vlimiter = 1;      // Tracer%ppm_hlimiter / ppm_vlimiter defaults (ocean_types.F90:1018-1019)
    // halo machinery
    HaloPlan plan[4];                  // indexed by flags (1 = X, 2 = Y, 3 = XY)
    HaloPlan plan1;                    // halo-1 full update of data-domain arrays (field(taup1), OM:1903-1911)
    double *sendbuf = 0, *recvbuf = 0;
    size_t bufcap = 0;
    // host-pointer mode mirrors
    std::vector<double *> hm;          // generic pool of data-domain device arrays
    double *hm_w = 0;
    cudaStream_t stream = 0;           // library-owned stream for the host-pointer entry points
    cudaStream_t s_up = 0, s_down = 0; // copy streams of the pipelined host-pointer path
    cudaStream_t s_comm = 0;           // halo exchange stream (overlapped with interior tiles)
    cudaStream_t s_edge = 0;           // edge tiles / chunks of the fused driver: they run NEXT TO the interior launch, not after it
    cudaEvent_t ev_edge[4] = {};
    cudaEvent_t ev_sync[4] = {};
    int overlap = 1;                   // MOM5ADV_OVERLAP=0 disables the comm/compute overlap
    int y_rows = 32;
    int fuse = 1;                      // MOM5ADV_FUSE=0: three separate sweeps instead of z + fused x/y
    int horz_fused = 1;                // MOM5ADV_HORZ_FUSED=0: quicker / upwind horizontal arm as flux kernel + fold fix + divergence kernel
    int tma = 3;                       // MOM5ADV_TMA: bit 0 = TMA-staged z sweep, bit 1 = TMA-staged fused x/y pass (default 3 = both); 0 = per-thread
                                       // LDGSTS staging everywhere (also what blocks with an odd ni+2 or unaligned bases get)
    int f_rows = 64;
    std::vector<cudaEvent_t> ev_up, ev_done;
    int banded = 1;                    // MOM5ADV_BANDED=0: host-pointer sweby_all pipelines over tracers instead of j-bands
    cudaEvent_t ev[6] = {};
    bool ev_valid = false;
    int64_t launches = 0;
    std::set<const void *> smem_ok;    // kernels whose dynamic shared-memory limit has been raised on this handle's device
    long long h2d_bytes = 0, d2h_bytes = 0;   // bytes the last host-pointer call moved over the link (mom5adv_last_transfer_bytes)
    std::map<const void *, size_t> pinned;    // caller arrays page-locked by the library (cudaHostRegister), by base address
    int pin = 1;                       // MOM5ADV_PIN=0: never page-lock caller arrays
    int host_threads = 8;              // MOM5ADV_HOST_THREADS: threads of the host-side th_tendency += adv_tendency
    std::vector<double *> hstage;      // pinned staging for adv_tendency bands when the caller does not want adv back
    size_t hstage_elems = 0;
    std::vector<cudaEvent_t> ev_dl;    // download-complete events of the banded pipeline
    double *met_ring = 0, *met_y = 0;  // packed 2-D metrics for the TMA-staged fused pass: (dyte, datr) and (dxtn, dytn) planes
    unsigned *zbits = 0;               // per-column mask bit strings for the z sweep (k_build_zbits), nzw words per column
    int nzw = 0;
    std::map<TmapKey, CUtensorMap> tmaps;   // tensor maps by (base pointer, extents, box): encoded once per caller array
};

#ifndef YROWS_MAX
#define YROWS_MAX 32
#endif
#ifndef FROWS_MAX
#define FROWS_MAX 256
#endif

#define LAUNCH(h, kern, grid, block, smem, st, ...)  \
    do {                                             \
        kern<<<grid, block, smem, st>>>(__VA_ARGS__); \
        (h)->launches++;                             \
    } while (0)

static size_t n3(const mom5adv_ctx *h) { return (size_t)h->g.slab * h->g.nk; }
static bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }
static size_t nh2(const mom5adv_ctx *h) { return (size_t)h->g.tslab * h->g.nk; }

// ------------------------------------------------------------------------------------------------
// device selection (one rank per GPU: the Fortran shim calls this with its node-local rank before mom5adv_init)
// ------------------------------------------------------------------------------------------------
extern "C" int mom5adv_device_count(int *count)
{
    if (!count) { set_error("mom5adv_device_count: null argument"); return MOM5ADV_EINVAL; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { *count = 0; set_error("mom5adv_device_count: no CUDA device"); return MOM5ADV_ENOGPU; }
    *count = n;
    return 0;
}
extern "C" int mom5adv_set_device(int node_local_rank)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { set_error("mom5adv_set_device: no CUDA device"); return MOM5ADV_ENOGPU; }
    if (node_local_rank < 0) { set_error("mom5adv_set_device: negative rank"); return MOM5ADV_EINVAL; }
    CUDA_TRY(cudaSetDevice(node_local_rank % n));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// layout helpers (mpp_compute_extent, MPPI/mpp_domains_define.inc:187-273)
// ------------------------------------------------------------------------------------------------
static int compute_extent(int isg, int ieg, int ndivs, std::vector<int> &ibegin, std::vector<int> &iend)
{
    ibegin.assign(ndivs, 0);
    iend.assign(ndivs, 0);
    return mpp_compute_extent_c(isg, ieg, ndivs, ibegin.data(), iend.data());   // include/mpp_layout.h (shared with the oracle)
}

static int find_div(const std::vector<int> &b, const std::vector<int> &e, int gidx)
{
    for (size_t d = 0; d < b.size(); d++)
        if (gidx >= b[d] && gidx <= e[d]) return (int)d;
    return -1;
}

// Build the message lists of one update (same algorithm as mom5_b200/domain.py:Decomposition.exchange_plan;
// tests/test_plan.py checks the two against each other through mom5adv_debug_plan).
struct DomInfo {
    int ni_g, nj_g, px, py, cyclic_x, cyclic_y, tripolar;
    const std::vector<int> *ibeg, *iend, *jbeg, *jend;
};

static void recv_strips(const DomInfo &D, int rank, int flags, int halo, std::vector<Msg> &recv, std::vector<Msg> &send_on_peer)
{
    const int bx = rank % D.px, by = rank / D.px;
    const int i0g = (*D.ibeg)[bx], j0g = (*D.jbeg)[by];
    const int ni = (*D.iend)[bx] - i0g + 1, nj = (*D.jend)[by] - j0g + 1;
    struct R { int ia, ib, ja, jb; };
    std::vector<R> regs;
    if (flags & 1) { regs.push_back({1 - halo, 0, 1, nj}); regs.push_back({ni + 1, ni + halo, 1, nj}); }
    if (flags & 2) { regs.push_back({1, ni, 1 - halo, 0}); regs.push_back({1, ni, nj + 1, nj + halo}); }
    if ((flags & 1) && (flags & 2)) {
        regs.push_back({1 - halo, 0, 1 - halo, 0});
        regs.push_back({ni + 1, ni + halo, 1 - halo, 0});
        regs.push_back({1 - halo, 0, nj + 1, nj + halo});
        regs.push_back({ni + 1, ni + halo, nj + 1, nj + halo});
    }
    for (const R &r : regs) {
        const bool beyond_n = (r.ja + j0g - 1) > D.nj_g;
        const bool fold = D.tripolar && beyond_n;
        const int nxr = r.ib - r.ia + 1, nyr = r.jb - r.ja + 1;
        std::vector<int> sx(nxr), ox(nxr), sy(nyr), oy(nyr);
        for (int p = 0; p < nxr; p++) {
            int ig = r.ia + p + i0g - 1;
            if (fold) ig = D.ni_g + 1 - ig;
            bool ok = true;
            if (ig < 1) { if (D.cyclic_x) ig += D.ni_g; else ok = false; }
            else if (ig > D.ni_g) { if (D.cyclic_x) ig -= D.ni_g; else ok = false; }
            if (ok && (ig < 1 || ig > D.ni_g)) ok = false;
            sx[p] = ig;
            ox[p] = ok ? find_div(*D.ibeg, *D.iend, ig) : -1;
        }
        for (int q = 0; q < nyr; q++) {
            int jg = r.ja + q + j0g - 1;
            bool ok = true;
            if (jg > D.nj_g) {
                if (D.tripolar) jg = 2 * D.nj_g + 1 - jg;
                else if (D.cyclic_y) jg -= D.nj_g;
                else ok = false;
            } else if (jg < 1) {
                if (D.cyclic_y) jg += D.nj_g; else ok = false;
            }
            if (ok && (jg < 1 || jg > D.nj_g)) ok = false;
            sy[q] = jg;
            oy[q] = ok ? find_div(*D.jbeg, *D.jend, jg) : -1;
        }
        // maximal runs of constant valid owner
        auto runs = [](const std::vector<int> &o) {
            std::vector<std::pair<int, int>> out;
            int a = -1;
            for (int q = 0; q < (int)o.size(); q++) {
                if (o[q] >= 0 && a < 0) a = q;
                if (a >= 0 && (q == (int)o.size() - 1 || o[q + 1] != o[a])) {
                    out.push_back({a, q});
                    a = -1;
                }
            }
            return out;
        };
        for (auto rx : runs(ox))
            for (auto ry : runs(oy)) {
                const int peer = ox[rx.first] + D.px * oy[ry.first];
                const int p_i0g = (*D.ibeg)[ox[rx.first]], p_j0g = (*D.jbeg)[oy[ry.first]];
                int si_a = sx[rx.first] - p_i0g + 1, si_b = sx[rx.second] - p_i0g + 1;
                int sj_a = sy[ry.first] - p_j0g + 1, sj_b = sy[ry.second] - p_j0g + 1;
                Msg mr{peer, r.ia + rx.first, r.ia + rx.second, r.ja + ry.first, r.ja + ry.second, fold ? 1 : 0};
                Msg ms{rank, std::min(si_a, si_b), std::max(si_a, si_b), std::min(sj_a, sj_b), std::max(sj_a, sj_b), fold ? 1 : 0};
                recv.push_back(mr);
                send_on_peer.push_back(ms);
            }
    }
}

static void build_plan(const DomInfo &D, int rank, int flags, int halo, HaloPlan &P)
{
    P.sends.clear();
    P.recvs.clear();
    for (int r = 0; r < D.px * D.py; r++) {
        std::vector<Msg> rv, sd;
        recv_strips(D, r, flags, halo, rv, sd);
        for (size_t q = 0; q < rv.size(); q++) {
            if (r == rank) P.recvs.push_back(rv[q]);
            if (rv[q].peer == rank) {
                Msg s = sd[q];
                s.peer = r;
                P.sends.push_back(s);
            }
        }
    }
}

// debug / test hook: flat dump of a plan without any GPU. out: rows of 7 ints
// (is_send, peer, i0, i1, j0, j1, flip); returns the number of rows (<= max_rows) or a negative error.
extern "C" int mom5adv_debug_plan(int ni_g, int nj_g, int px, int py, int cyclic_x, int cyclic_y, int tripolar, int rank,
                                  int flags, int halo, int *out, int max_rows)
{
    std::vector<int> ib, ie, jb, je;
    if (compute_extent(1, ni_g, px, ib, ie) || compute_extent(1, nj_g, py, jb, je)) {
        set_error("mom5adv_debug_plan: bad extents");
        return MOM5ADV_EINVAL;
    }
    DomInfo D{ni_g, nj_g, px, py, cyclic_x, cyclic_y, tripolar, &ib, &ie, &jb, &je};
    HaloPlan P;
    build_plan(D, rank, flags, halo, P);
    int n = 0;
    for (int s = 1; s >= 0; s--)
        for (const Msg &m : (s ? P.sends : P.recvs)) {
            if (n >= max_rows) return n;
            int *o = out + 7 * n++;
            o[0] = s; o[1] = m.peer; o[2] = m.i0; o[3] = m.i1; o[4] = m.j0; o[5] = m.j1; o[6] = m.flip;
        }
    return n;
}

extern "C" int mom5adv_debug_extent(int isg, int ieg, int ndivs, int *ibegin, int *iend)
{
    std::vector<int> b, e;
    int rc = compute_extent(isg, ieg, ndivs, b, e);
    if (rc) return MOM5ADV_EINVAL;
    for (int d = 0; d < ndivs; d++) { ibegin[d] = b[d]; iend[d] = e[d]; }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// communicator
// ------------------------------------------------------------------------------------------------
extern "C" int mom5adv_comm_unique_id(char id_out[128])
{
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    NCCL_NEED();
    NCCL_TRY(N->GetUniqueId(&id));
    memcpy(id_out, &id, 128);
    return 0;
}
extern "C" int mom5adv_comm_create(const char id[128], int rank, int nranks, mom5adv_comm *out)
{
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    ncclComm_t c;
    NCCL_NEED();
    NCCL_TRY(N->CommInitRank(&c, nranks, uid, rank));
    *out = new mom5adv_comm_s{(void *)c, rank, nranks, true};
    return 0;
}
extern "C" int mom5adv_comm_from_nccl(void *nccl_comm, int rank, int nranks, mom5adv_comm *out)
{
    if (!nccl_comm) { set_error("mom5adv_comm_from_nccl: null communicator"); return MOM5ADV_EINVAL; }
    *out = new mom5adv_comm_s{nccl_comm, rank, nranks, false};
    return 0;
}
extern "C" int mom5adv_comm_destroy(mom5adv_comm c)
{
    if (!c) return 0;
    if (c->owned && nccl_api()) nccl_api()->CommDestroy((ncclComm_t)c->nccl);
    delete c;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// halo update of nf h2 fields
// ------------------------------------------------------------------------------------------------
template <int LAYOUT>
static int halo_update_l(mom5adv_ctx *h, double *const *fields, int nf, int flags, cudaStream_t st)
{
    const HaloPlan &P = (LAYOUT == 0) ? h->plan[flags & 3] : (LAYOUT == 1) ? h->plan1 : h->plan4[flags & 3];
    const int nk = h->g.nk;
    for (int f0 = 0; f0 < nf; f0 += HALO_MAXF) {
        const int nfc = std::min(HALO_MAXF, nf - f0);
        // ---- local (self) strips ----
        CopyArgs L{};
        L.nf = nfc; L.nk = nk;
        for (int n = 0; n < nfc; n++) L.f[n] = fields[f0 + n];
        // match each self-recv with the self-send of the same ordinal
        std::vector<const Msg *> rs, ss;
        for (const Msg &m : P.recvs) if (m.peer == h->rank) rs.push_back(&m);
        for (const Msg &m : P.sends) if (m.peer == h->rank) ss.push_back(&m);
        if (rs.size() != ss.size()) { set_error("halo plan: self send/recv mismatch"); return MOM5ADV_EINVAL; }
        auto flush_local = [&](CopyArgs &A) {
            if (A.nmsg == 0) return;
            A.total = A.start[A.nmsg];
            const int nb = (int)std::min<long long>((A.total + 255) / 256, 148 * 16);
            LAUNCH(h, (k_halo<0, LAYOUT>), nb, 256, 0, st, h->g, A);
            A.nmsg = 0;
        };
        L.nmsg = 0; L.start[0] = 0;
        for (size_t q = 0; q < rs.size(); q++) {
            const Msg &r = *rs[q], &s = *ss[q];
            CopyDesc d{};
            d.si0 = s.i0; d.sj0 = s.j0; d.si1 = s.i1; d.sj1 = s.j1;
            d.di0 = r.i0; d.dj0 = r.j0;
            d.w = r.i1 - r.i0 + 1; d.h = r.j1 - r.j0 + 1; d.flip = r.flip;
            L.d[L.nmsg] = d;
            L.start[L.nmsg + 1] = L.start[L.nmsg] + (long long)d.w * d.h * nk * nfc;
            if (++L.nmsg == HALO_MAXMSG) { flush_local(L); L.start[0] = 0; }
        }
        flush_local(L);

        // ---- remote strips ----
        std::vector<const Msg *> rr, sr;
        for (const Msg &m : P.recvs) if (m.peer != h->rank) rr.push_back(&m);
        for (const Msg &m : P.sends) if (m.peer != h->rank) sr.push_back(&m);
        if (rr.empty() && sr.empty()) continue;
        if (!h->comm) { set_error("halo update needs a communicator (layout %dx%d)", h->px, h->py); return MOM5ADV_EINVAL; }
        // group by peer, keeping plan order within a peer
        auto by_peer = [](std::vector<const Msg *> &v) {
            std::stable_sort(v.begin(), v.end(), [](const Msg *a, const Msg *b) { return a->peer < b->peer; });
        };
        by_peer(rr); by_peer(sr);
        if (rr.size() > HALO_MAXMSG || sr.size() > HALO_MAXMSG) { set_error("halo plan too large"); return MOM5ADV_EINVAL; }
        CopyArgs S{}, R{};
        S.nf = R.nf = nfc; S.nk = R.nk = nk;
        for (int n = 0; n < nfc; n++) S.f[n] = R.f[n] = fields[f0 + n];
        S.start[0] = R.start[0] = 0;
        for (const Msg *m : sr) {
            CopyDesc d{};
            d.si0 = m->i0; d.sj0 = m->j0; d.si1 = m->i1; d.sj1 = m->j1;
            d.w = m->i1 - m->i0 + 1; d.h = m->j1 - m->j0 + 1; d.flip = m->flip;
            d.off = S.start[S.nmsg];
            S.d[S.nmsg] = d;
            S.start[S.nmsg + 1] = S.start[S.nmsg] + (long long)d.w * d.h * nk * nfc;
            S.nmsg++;
        }
        for (const Msg *m : rr) {
            CopyDesc d{};
            d.di0 = m->i0; d.dj0 = m->j0;
            d.w = m->i1 - m->i0 + 1; d.h = m->j1 - m->j0 + 1; d.flip = m->flip;
            d.off = R.start[R.nmsg];
            R.d[R.nmsg] = d;
            R.start[R.nmsg + 1] = R.start[R.nmsg] + (long long)d.w * d.h * nk * nfc;
            R.nmsg++;
        }
        S.total = S.start[S.nmsg];
        R.total = R.start[R.nmsg];
        const size_t need = (size_t)std::max(S.total, R.total);
        if (need > h->bufcap) {
            if (h->sendbuf) cudaFree(h->sendbuf);
            if (h->recvbuf) cudaFree(h->recvbuf);
            CUDA_TRY(cudaStreamSynchronize(st));
            CUDA_TRY(cudaMalloc(&h->sendbuf, need * sizeof(double)));
            CUDA_TRY(cudaMalloc(&h->recvbuf, need * sizeof(double)));
            h->bufcap = need;
        }
        S.buf = h->sendbuf;
        R.buf = h->recvbuf;
        if (S.total) {
            const int nb = (int)std::min<long long>((S.total + 255) / 256, 148 * 16);
            LAUNCH(h, (k_halo<1, LAYOUT>), nb, 256, 0, st, h->g, S);
        }
        ncclComm_t comm = (ncclComm_t)h->comm->nccl;
        NCCL_NEED();
        NCCL_TRY(N->GroupStart());
        for (int m = 0; m < S.nmsg;) {   // one send per peer
            int e = m;
            while (e < S.nmsg && sr[e]->peer == sr[m]->peer) e++;
            NCCL_TRY(N->Send(h->sendbuf + S.start[m], (size_t)(S.start[e] - S.start[m]), ncclDouble, sr[m]->peer, comm, st));
            m = e;
        }
        for (int m = 0; m < R.nmsg;) {
            int e = m;
            while (e < R.nmsg && rr[e]->peer == rr[m]->peer) e++;
            NCCL_TRY(N->Recv(h->recvbuf + R.start[m], (size_t)(R.start[e] - R.start[m]), ncclDouble, rr[m]->peer, comm, st));
            m = e;
        }
        NCCL_TRY(N->GroupEnd());
        if (R.total) {
            const int nb = (int)std::min<long long>((R.total + 255) / 256, 148 * 16);
            LAUNCH(h, (k_halo<2, LAYOUT>), nb, 256, 0, st, h->g, R);
        }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int halo_update(mom5adv_ctx *h, double *const *fields, int nf, int flags, cudaStream_t st)
{
    return halo_update_l<0>(h, fields, nf, flags, st);
}

// ------------------------------------------------------------------------------------------------
// init / finalize
// ------------------------------------------------------------------------------------------------
template <class T>
static int upload(T **dst, const T *src, size_t n)
{
    CUDA_TRY(cudaMalloc(dst, n * sizeof(T)));
    CUDA_TRY(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

static int quicker_setup(mom5adv_ctx *h, const mom5adv_grid *G, cudaStream_t st);
static int mirror(mom5adv_ctx *h, size_t idx, double **out);
static int init_body(mom5adv_ctx *h, const mom5adv_grid *G, int ntracers_max, mom5adv_comm comm);

extern "C" int mom5adv_init(const mom5adv_grid *G, int ntracers_max, mom5adv_comm comm, mom5adv_handle *out)
{
    if (!G || !out || ntracers_max < 1) { set_error("mom5adv_init: bad arguments"); return MOM5ADV_EINVAL; }
    if (G->have_obc) { set_error("mom5adv_init: open boundaries (have_obc) are not covered by the GPU path"); return MOM5ADV_EUNSUP; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("mom5adv_init: no CUDA device"); return MOM5ADV_ENOGPU; }
    if (G->tripolar && G->cyclic_y) { set_error("mom5adv_init: tripolar and cyclic_y are exclusive"); return MOM5ADV_EINVAL; }
    if (G->tripolar && (G->ni_global % 2)) { set_error("mom5adv_init: tripolar fold needs an even ni_global"); return MOM5ADV_EINVAL; }
    mom5adv_ctx *h = new mom5adv_ctx();
    // every failure below goes through ONE cleanup: mom5adv_finalize copes with a partially built context
    const int rc = init_body(h, G, ntracers_max, comm);
    if (rc) { mom5adv_finalize(h); return rc; }
    *out = h;
    return 0;
}

static int init_body(mom5adv_ctx *h, const mom5adv_grid *G, int ntracers_max, mom5adv_comm comm)
{
    h->ntr_max = ntracers_max;
    h->ni_g = G->ni_global; h->nj_g = G->nj_global; h->px = G->layout_x; h->py = G->layout_y;
    h->cyclic_x = G->cyclic_x; h->cyclic_y = G->cyclic_y; h->tripolar = G->tripolar;
    h->isc_g = G->isc; h->jsc_g = G->jsc;
    auto from_extent = [](const int *ext, int n, int total, std::vector<int> &b, std::vector<int> &e) {
        b.resize(n); e.resize(n);
        int s = 1;
        for (int d = 0; d < n; d++) { b[d] = s; e[d] = s + ext[d] - 1; s += ext[d]; }
        return (s - 1 == total) ? 0 : 1;
    };
    int rc = G->x_extent ? from_extent(G->x_extent, h->px, h->ni_g, h->ibeg, h->iend) : compute_extent(1, h->ni_g, h->px, h->ibeg, h->iend);
    rc |= G->y_extent ? from_extent(G->y_extent, h->py, h->nj_g, h->jbeg, h->jend) : compute_extent(1, h->nj_g, h->py, h->jbeg, h->jend);
    if (rc) { set_error("mom5adv_init: invalid domain extents"); return MOM5ADV_EINVAL; }
    h->ix = find_div(h->ibeg, h->iend, G->isc);
    h->iy = find_div(h->jbeg, h->jend, G->jsc);
    if (h->ix < 0 || h->iy < 0 || h->ibeg[h->ix] != G->isc || h->iend[h->ix] != G->iec || h->jbeg[h->iy] != G->jsc ||
        h->jend[h->iy] != G->jec) {
        set_error("mom5adv_init: compute domain (%d:%d,%d:%d) does not match layout %dx%d", G->isc, G->iec, G->jsc, G->jec, h->px, h->py);
        return MOM5ADV_EINVAL;
    }
    h->rank = h->ix + h->px * h->iy;
    if (h->px * h->py > 1) {
        if (!comm || comm->nranks != h->px * h->py || comm->rank != h->rank) {
            set_error("mom5adv_init: layout %dx%d needs a communicator with %d ranks and rank == ix + px*iy", h->px, h->py, h->px * h->py);
            return MOM5ADV_EINVAL;
        }
        h->comm = comm;
    }
    Geom &g = h->g;
    g.ni = G->iec - G->isc + 1; g.nj = G->jec - G->jsc + 1; g.nk = G->nk;
    g.nxd = g.ni + 2; g.nyd = g.nj + 2; g.slab = (long long)g.nxd * g.nyd;
    g.tpitch = ((g.ni + TOFF + 3) + 15) / 16 * 16; g.tslab = (long long)g.tpitch * (g.nj + 4);
    g.mpitch = g.tpitch; g.mslab = g.tslab;
    g.p4 = g.ni + 8; g.s4 = (long long)g.p4 * (g.nj + 8);
    {   // a halo strip must come from ONE neighbour (FMS refuses halos wider than the compute domain as well)
        const bool src_x = h->cyclic_x || h->px > 1, src_y = h->cyclic_y || h->py > 1 || h->tripolar;
        if ((src_x && g.ni < 2) || (src_y && g.nj < 2)) {
            set_error("mom5adv_init: local block %dx%d is narrower than the width-2 halo it has to fill", g.ni, g.nj);
            return MOM5ADV_EUNSUP;
        }
    }
    if ((unsigned long long)g.tslab * (unsigned long long)g.nk >= (1ull << 32) || (unsigned long long)g.slab * (unsigned long long)(g.nk + 2) >= (1ull << 32)) {
        set_error("mom5adv_init: local block too large (the kernels index with 32-bit element offsets: < 2^32 elements per array)");
        return MOM5ADV_EUNSUP;
    }
    const size_t n2 = (size_t)g.slab;
    if (upload(&h->dat, G->dat, n2) || upload(&h->datr, G->datr, n2) || upload(&h->dxte, G->dxte, n2) ||
        upload(&h->dyte, G->dyte, n2) || upload(&h->dxtn, G->dxtn, n2) || upload(&h->dytn, G->dytn, n2) ||
        upload(&h->tmask, G->tmask, n3(h))) return MOM5ADV_ECUDA;
    CUDA_TRY(cudaMalloc(&h->met_ring, 2 * n2 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&h->met_y, 2 * n2 * sizeof(double)));
    CUDA_TRY(cudaMemcpy(h->met_ring, h->dyte, n2 * sizeof(double), cudaMemcpyDeviceToDevice));
    CUDA_TRY(cudaMemcpy(h->met_ring + n2, h->datr, n2 * sizeof(double), cudaMemcpyDeviceToDevice));
    CUDA_TRY(cudaMemcpy(h->met_y, h->dxtn, n2 * sizeof(double), cudaMemcpyDeviceToDevice));
    CUDA_TRY(cudaMemcpy(h->met_y + n2, h->dytn, n2 * sizeof(double), cudaMemcpyDeviceToDevice));
    CUDA_TRY(cudaMalloc(&h->mask, (size_t)g.mslab * g.nk));
    CUDA_TRY(cudaMemset(h->mask, 0, (size_t)g.mslab * g.nk));
    h->tmA.assign(ntracers_max, nullptr); h->tmB.assign(ntracers_max, nullptr);
    for (int n = 0; n < ntracers_max; n++) {
        CUDA_TRY(cudaMalloc(&h->tmA[n], nh2(h) * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->tmB[n], nh2(h) * sizeof(double)));
        CUDA_TRY(cudaMemset(h->tmA[n], 0, nh2(h) * sizeof(double)));   // wall halos stay 0 forever (OTA:1660-1666)
        CUDA_TRY(cudaMemset(h->tmB[n], 0, nh2(h) * sizeof(double)));
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_up, cudaStreamNonBlocking));
    {   // the halo stream outranks the compute stream: its pack / NCCL / unpack kernels are scheduled ahead of the (tens of
        // thousands of) interior blocks already queued, so the exchange starts at once instead of behind them
        int prio_lo = 0, prio_hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&h->s_comm, cudaStreamNonBlocking, prio_hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&h->s_edge, cudaStreamNonBlocking, prio_hi));
        for (int e = 0; e < 4; e++) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_edge[e], cudaEventDisableTiming));
    }
    for (int e = 0; e < 4; e++) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_sync[e], cudaEventDisableTiming));
    if (const char *ov = getenv("MOM5ADV_OVERLAP")) h->overlap = atoi(ov);
    if (const char *fu = getenv("MOM5ADV_FUSE")) h->fuse = atoi(fu);
    if (const char *tm = getenv("MOM5ADV_TMA")) h->tma = atoi(tm);
    if (const char *hf = getenv("MOM5ADV_HORZ_FUSED")) h->horz_fused = atoi(hf);
    if (const char *bd = getenv("MOM5ADV_BANDED")) h->banded = atoi(bd);
    if (const char *pn = getenv("MOM5ADV_PIN")) h->pin = atoi(pn);
    h->host_threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency() / 2));
    if (const char *ht = getenv("MOM5ADV_HOST_THREADS")) h->host_threads = std::max(1, atoi(ht));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_down, cudaStreamNonBlocking));
    h->ev_up.assign(ntracers_max, nullptr); h->ev_done.assign(ntracers_max, nullptr);
    for (int n = 0; n < ntracers_max; n++) {
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_up[n], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_done[n], cudaEventDisableTiming));
    }
    for (int e = 0; e < 6; e++) CUDA_TRY(cudaEventCreate(&h->ev[e]));
    DomInfo D{h->ni_g, h->nj_g, h->px, h->py, h->cyclic_x, h->cyclic_y, h->tripolar, &h->ibeg, &h->iend, &h->jbeg, &h->jend};
    for (int f = 1; f <= 3; f++) build_plan(D, h->rank, f, 2, h->plan[f]);
    build_plan(D, h->rank, 3, 1, h->plan1);
    for (int f = 1; f <= 3; f++) build_plan(D, h->rank, f, 4, h->plan4[f]);

    // tmask_mdfl == tmask_quick: compute domain := Grd%tmask, full halo-2 update (OTA:1668-1675, 1478-1487), stored as u8
    cudaStream_t st = h->stream;
    dim3 gb((g.ni + 127) / 128, g.nj, g.nk);
    LAUNCH(h, k_d1_to_h2, gb, 128, 0, st, g, h->tmask, h->tmA[0]);
    double *f0[1] = {h->tmA[0]};
    if ((rc = halo_update(h, f0, 1, 3, st))) { return rc; }
    dim3 gm((g.ni + 4 + 127) / 128, g.nj + 4, g.nk);
    LAUNCH(h, k_h2_to_mask, gm, 128, 0, st, g, h->tmA[0], h->mask);
    h->nib_pitch = (g.nxd + 15) / 16 * 16;
    const size_t nibp = (size_t)h->nib_pitch * g.nyd * g.nk;
    CUDA_TRY(cudaMalloc(&h->nibx, n3(h)));
    CUDA_TRY(cudaMalloc(&h->niby, n3(h)));
    CUDA_TRY(cudaMalloc(&h->nibx_p, nibp));
    CUDA_TRY(cudaMalloc(&h->niby_p, nibp));
    CUDA_TRY(cudaMemsetAsync(h->nibx_p, 0, nibp, st));
    CUDA_TRY(cudaMemsetAsync(h->niby_p, 0, nibp, st));
    LAUNCH(h, k_build_nibbles, dim3((g.ni + 2 + 127) / 128, g.nj + 2, g.nk), 128, 0, st, g, h->mask, h->nibx, h->niby, h->nibx_p, h->niby_p, h->nib_pitch);
    h->nzw = (g.nk + 3 + 31) / 32;
    CUDA_TRY(cudaMalloc(&h->zbits, (size_t)h->nzw * g.slab * sizeof(unsigned)));
    LAUNCH(h, k_build_zbits, dim3((g.ni + 2 + 127) / 128, g.nj + 2), 128, 0, st, g, h->mask, h->zbits, h->nzw);
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaMemset(h->tmA[0], 0, nh2(h) * sizeof(double)));
    if ((rc = quicker_setup(h, G, st))) return rc;
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int mom5adv_finalize(mom5adv_handle h)
{
    if (!h) return 0;
    cudaDeviceSynchronize();
    cudaGetLastError();   // a failed init leaves a sticky-free error behind; the frees below must not trip over it
    for (double *p : {h->dat, h->datr, h->dxte, h->dyte, h->dxtn, h->dytn, h->tmask, h->sendbuf, h->recvbuf, h->hm_w, h->st_ms, h->pp_tr, h->pp_m4, h->pp_da, h->pp_fz,
                      h->met_ring, h->met_y})
        if (p) cudaFree(p);
    for (uint8_t *p : {h->mask, h->nibx, h->niby, h->nibx_p, h->niby_p})
        if (p) cudaFree(p);
    if (h->zbits) cudaFree(h->zbits);
    for (auto &kv : h->pinned) if (kv.second) cudaHostUnregister(const_cast<void *>(kv.first));
    for (double *p : h->hstage) if (p) cudaFreeHost(p);
    for (cudaEvent_t e : h->ev_dl) if (e) cudaEventDestroy(e);
    for (double *p : h->tmA) if (p) cudaFree(p);
    for (double *p : h->tmB) if (p) cudaFree(p);
    for (double *p : h->hm) if (p) cudaFree(p);
    free_quickw(h->qw);
    for (int e = 0; e < 6; e++) if (h->ev[e]) cudaEventDestroy(h->ev[e]);
    for (cudaStream_t s : {h->stream, h->s_up, h->s_comm, h->s_down, h->s_edge}) if (s) cudaStreamDestroy(s);
    for (int e = 0; e < 4; e++) if (h->ev_edge[e]) cudaEventDestroy(h->ev_edge[e]);
    for (int e = 0; e < 4; e++) if (h->ev_sync[e]) cudaEventDestroy(h->ev_sync[e]);
    for (cudaEvent_t e : h->ev_up) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_done) if (e) cudaEventDestroy(e);
    delete h;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Sweby driver
// ------------------------------------------------------------------------------------------------
static int z_tiles(int ni);
static int pick_kchunk(const Geom &g, int per_level_threads)
{
    // aim for >= ~4 resident waves of threads; each extra chunk costs one redundant face per column
    const long long want = 148LL * 2048 * 2;
    int nch = (int)std::min<long long>((want + per_level_threads - 1) / per_level_threads, std::max(1, g.nk / 8));
    nch = std::max(nch, 1);
    return (g.nk + nch - 1) / nch;
}

// A launch covers `count` tiles first, first+step, ... of the sweep's tiled dimension (z: i-tiles of ZBX columns,
// x: i-tiles of 31 cells, y / fused xy: j-chunks); count < 0 = all of them.  x additionally takes a row range.
struct Part {
    int first = 0, step = 1, count = -1;
    int row_first = 1, row_last = -1;   // x only; row_last < 0 = nj
};
enum { PH_Z = 0, PH_X = 1, PH_Y = 2, PH_XY = 3 };

// 3-D tensor map over an array dimensioned (n0, n1, n2) Fortran-style (n0 contiguous), box b0 x b1 x b2; ESZ = 8 (FP64) or 1 (u8)
template <int ESZ>
static int tmap3_t(mom5adv_ctx *h, const void *base, unsigned long long n0, unsigned long long n1, unsigned long long n2, unsigned b0,
                   unsigned b1, unsigned b2, CUtensorMap *out)
{
    const TmapKey key{base, n0, n1, n2, b0, b1, b2};
    auto it = h->tmaps.find(key);
    if (it != h->tmaps.end()) { *out = it->second; return 0; }
    EncodeTiledFn enc = tma_encoder();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return MOM5ADV_ECUDA; }
    const cuuint64_t dims[3] = {n0, n1, n2};
    const cuuint64_t strides[2] = {n0 * ESZ, n0 * n1 * ESZ};
    const cuuint32_t box[3] = {b0, b1, b2}, estr[3] = {1, 1, 1};
    CUtensorMap m;
    const CUresult r = enc(&m, ESZ == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for a %llux%llux%llu array, box %ux%ux%u", (int)r, n0, n1, n2, b0, b1, b2); return MOM5ADV_ECUDA; }
    if (h->tmaps.size() > 4096) h->tmaps.clear();   // callers that hand in fresh allocations every step must not grow the cache forever
    h->tmaps[key] = m;
    *out = m;
    return 0;
}

static int tmap3(mom5adv_ctx *h, const double *base, unsigned long long n0, unsigned long long n1, unsigned long long n2, unsigned b0,
                 unsigned b1, unsigned b2, CUtensorMap *out)
{
    return tmap3_t<8>(h, base, n0, n1, n2, b0, b1, b2, out);
}

// TMA staging needs 16-byte aligned bases and row strides that are multiples of 16 bytes (an even ni+2)
static bool tma_ok(const mom5adv_ctx *h, int which) { return (h->tma & which) && (h->g.nxd % 2 == 0) && tma_encoder() != nullptr; }

template <int NT, int VAR, bool DIAG>
static int launch_z_tma(mom5adv_ctx *h, const SwebyArgs<NT> &b, dim3 grid, cudaStream_t st)
{
    const Geom &g = h->g;
    ZMaps<NT> m;
    int rc;
    for (int n = 0; n < NT; n++)
        if ((rc = tmap3(h, b.T[n], g.nxd, g.nyd, g.nk, ZBX, 1, ZT_KC, &m.T[n]))) return rc;
    if ((rc = tmap3(h, b.w, g.nxd, g.nyd, g.nk + 1, ZBX, 1, ZT_KC, &m.w)) || (rc = tmap3(h, b.rho, g.nxd, g.nyd, g.nk, ZBX, 1, ZT_KC, &m.rho))) return rc;
    if (h->smem_ok.insert((const void *)k_sweby_z_tma<NT, VAR, DIAG>).second)
        CUDA_TRY(cudaFuncSetAttribute(k_sweby_z_tma<NT, VAR, DIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZTLayout<NT>::BYTES));
    LAUNCH(h, (k_sweby_z_tma<NT, VAR, DIAG>), grid, ZBX, ZTLayout<NT>::BYTES, st, g, b, m, h->zbits, h->nzw);
    return 0;
}

template <int NT, int VAR, bool DIAG, bool UPD>
static int launch_xy_tma(mom5adv_ctx *h, const SwebyArgs<NT> &b, unsigned nblk, int nxb, int nxt, cudaStream_t st)
{
    const Geom &g = h->g;
    typedef FusedTmaLayout<NT, UPD> LY;
    FusedMaps<NT, UPD> m;
    memset(&m, 0, sizeof m);
    int rc;
    const unsigned long long nx = g.nxd, ny = g.nyd, nk = g.nk;
    for (int n = 0; n < NT; n++) {
        if ((rc = tmap3(h, b.T[n], nx, ny, nk, FT_RW, 1, 1, &m.T[n])) ||
            (rc = tmap3(h, b.tm_in[n], g.tpitch, g.nj + 4, nk, FT_RW, 1, 1, &m.tm[n]))) return rc;
        if (!UPD && b.accumulate && (rc = tmap3(h, b.th[n], nx, ny, nk, FT_RW, 1, 1, &m.th[n]))) return rc;
    }
    if ((rc = tmap3(h, b.u, nx, ny, nk, FT_RW, 1, 1, &m.u)) || (rc = tmap3(h, b.v, nx, ny, nk, FT_RW, 1, 1, &m.v)) ||
        (rc = tmap3(h, b.w, nx, ny, nk + 1, FT_RW, 1, 2, &m.w)) || (rc = tmap3(h, b.rho, nx, ny, nk, FT_RW, 1, 1, &m.rho)) ||
        (rc = tmap3(h, h->met_ring, nx, ny, 2, FT_RW, 1, 2, &m.met_ring)) || (rc = tmap3(h, h->dxte, nx, ny, 1, FT_RW, 1, 1, &m.dxte)) ||
        (rc = tmap3(h, h->met_y, nx, ny, 2, FT_RW, 1, 2, &m.met_y))) return rc;
    if (UPD && ((rc = tmap3(h, b.rho_m1, nx, ny, nk, FT_RW, 1, 1, &m.rho_m1)) || (rc = tmap3(h, b.rho_r, nx, ny, nk, FT_RW, 1, 1, &m.rho_r)))) return rc;
    if ((rc = tmap3_t<1>(h, h->nibx_p, h->nib_pitch, ny, nk, FT_NW, 1, 1, &m.nibx)) || (rc = tmap3_t<1>(h, h->niby_p, h->nib_pitch, ny, nk, FT_NW, 1, 1, &m.niby))) return rc;
    if (h->smem_ok.insert((const void *)k_sweby_xy_tma<NT, VAR, DIAG, UPD>).second) {
        CUDA_TRY(cudaFuncSetAttribute(k_sweby_xy_tma<NT, VAR, DIAG, UPD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LY::BYTES));
        // three 53 KB blocks per SM: ask for the largest shared-memory carveout
        CUDA_TRY(cudaFuncSetAttribute(k_sweby_xy_tma<NT, VAR, DIAG, UPD>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    }
    LAUNCH(h, (k_sweby_xy_tma<NT, VAR, DIAG, UPD>), nblk, 32 * FWARPS, LY::BYTES, st, g, b, m, nxb, nxt);
    return 0;
}

template <int NT>
static bool xy_tma_ok(const mom5adv_ctx *h, const SwebyArgs<NT> &b, bool upd)
{
    bool ok = tma_ok(h, 2) && aligned16(b.u) && aligned16(b.v) && aligned16(b.w) && aligned16(b.rho);
    for (int n = 0; n < NT; n++) ok = ok && aligned16(b.T[n]) && (upd || !b.accumulate || aligned16(b.th[n]));
    if (upd) ok = ok && aligned16(b.rho_m1) && aligned16(b.rho_r);
    return ok;
}

template <int NT, int VAR, bool DIAG>
static int launch_group(mom5adv_ctx *h, int phase, const Part &pt, const SwebyArgs<NT> &a, cudaStream_t st)
{
    const Geom &g = h->g;
    SwebyArgs<NT> b = a;
    b.tile_first = pt.first; b.tile_step = pt.step;
    b.row_first = pt.row_first; b.row_last = pt.row_last < 0 ? g.nj : pt.row_last;
    if (pt.count == 0) return 0;
    if (phase == PH_Z) {
        b.kc = pick_kchunk(g, g.ni * g.nj);
        const int nzt = z_tiles(g.ni);
        const int nrows = b.row_last - b.row_first + 1;
        if (nrows <= 0) return 0;
        dim3 grid(pt.count < 0 ? nzt : pt.count, nrows, (g.nk + b.kc - 1) / b.kc);
        bool tma = tma_ok(h, 1);
        for (int n = 0; n < NT; n++) tma = tma && aligned16(b.T[n]);
        tma = tma && aligned16(b.w) && aligned16(b.rho);
        b.zbits = h->zbits; b.nzw = h->nzw;
        if (tma) return launch_z_tma<NT, VAR, DIAG>(h, b, grid, st);
        LAUNCH(h, (k_sweby_z<NT, VAR, DIAG>), grid, ZBX, 0, st, g, b);
    } else if (phase == PH_X) {
        b.kc = pick_kchunk(g, g.ni * g.nj);
        const int nxt = (g.ni + 30) / 31;
        const int nrows = b.row_last - b.row_first + 1;
        if (nrows <= 0) return 0;
        dim3 grid(pt.count < 0 ? nxt : pt.count, (nrows + XWARPS - 1) / XWARPS, (g.nk + b.kc - 1) / b.kc);
        LAUNCH(h, (k_sweby_x<NT, VAR, DIAG>), grid, dim3(32, XWARPS), 0, st, g, b);
    } else if (phase == PH_Y) {
        const int YBX = 32 * YWARPS;
        const int nxt = (g.ni + YBX - 1) / YBX;
        b.kc = h->y_rows;
        const int njc = (g.nj + b.kc - 1) / b.kc;
        LAUNCH(h, (k_sweby_y<NT, VAR, DIAG>), (unsigned)(g.nk * nxt * (pt.count < 0 ? njc : pt.count)), YBX, 0, st, g, b, nxt);
    } else {
        const int nxt = (g.ni + 30) / 31, nxb = (nxt + FWARPS - 1) / FWARPS;
        b.kc = h->f_rows;
        const int njc = (g.nj + b.kc - 1) / b.kc;
        const unsigned nblk = (unsigned)(g.nk * nxb * (pt.count < 0 ? njc : pt.count));
        if (VAR == VAR_ALL && !DIAG && b.Tnew[0]) {   // time update in the epilogue (mom5adv_sweby_all_step_dev)
            if (xy_tma_ok<NT>(h, b, true)) return launch_xy_tma<NT, VAR_ALL, false, true>(h, b, nblk, nxb, nxt, st);
            typedef FusedLayout<NT, true> LYU;
            if (h->smem_ok.insert((const void *)k_sweby_xy<NT, VAR_ALL, false, true>).second)   // once per handle (= per device)
                CUDA_TRY(cudaFuncSetAttribute(k_sweby_xy<NT, VAR_ALL, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LYU::BYTES));
            LAUNCH(h, (k_sweby_xy<NT, VAR_ALL, false, true>), nblk, 32 * FWARPS, LYU::BYTES, st, g, b, nxb, nxt);
            return 0;
        }
        if (xy_tma_ok<NT>(h, b, false)) return launch_xy_tma<NT, VAR, DIAG, false>(h, b, nblk, nxb, nxt, st);
        if (h->smem_ok.insert((const void *)k_sweby_xy<NT, VAR, DIAG>).second)
            CUDA_TRY(cudaFuncSetAttribute(k_sweby_xy<NT, VAR, DIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FusedLayout<NT>::BYTES));
        LAUNCH(h, (k_sweby_xy<NT, VAR, DIAG>), nblk, 32 * FWARPS, FusedLayout<NT>::BYTES, st, g, b, nxb, nxt);
    }
    return 0;
}

struct SwebyCall {
    int ntr, var;
    double dtime, sl;
    const double *const *T;
    double *const *th, *const *adv;
    const double *u, *v, *w, *rho;
    double *const *fx, *const *fy, *const *fz, *const *ax, *const *ay, *const *az;
    int accumulate;
    // time update fused into the x/y pass (mom5adv_sweby_all_step_dev); th / adv entries may then be null
    double *const *Tnew = nullptr;
    const double *rho_m1 = nullptr, *rho_r = nullptr;
};

// diag_ok = false: never write diagnostics from this launch (edge-row x sweep ahead of the fused pass, which writes them)
template <int NT>
static int run_phase(mom5adv_ctx *h, const SwebyCall &c, int n0, int phase, const Part &pt, bool diag_ok, cudaStream_t st)
{
    SwebyArgs<NT> a{};
    bool diag = false;
    double *const *fl = phase == PH_Z ? c.fz : phase == PH_Y ? c.fy : c.fx;
    double *const *da = phase == PH_Z ? c.az : phase == PH_Y ? c.ay : c.ax;
    for (int n = 0; n < NT; n++) {
        a.T[n] = c.T[n0 + n];
        if (phase == PH_Z) { a.tm_in[n] = h->tmA[n0 + n]; }
        if (phase == PH_X || phase == PH_XY) { a.tm_in[n] = h->tmA[n0 + n]; a.tm_out[n] = h->tmB[n0 + n]; }
        if (phase == PH_Y) { a.tm_in[n] = h->tmB[n0 + n]; }
        if (phase == PH_Y || phase == PH_XY) { a.th[n] = c.th ? c.th[n0 + n] : nullptr; a.adv[n] = c.adv ? c.adv[n0 + n] : nullptr; }
        if (phase == PH_XY && c.Tnew) a.Tnew[n] = c.Tnew[n0 + n];
        if (diag_ok) {
            a.flux[n] = fl ? fl[n0 + n] : nullptr;
            a.dadv[n] = da ? da[n0 + n] : nullptr;
            if (phase == PH_XY) {
                a.flux2[n] = c.fy ? c.fy[n0 + n] : nullptr;
                a.dadv2[n] = c.ay ? c.ay[n0 + n] : nullptr;
            }
        }
        diag |= a.flux[n] || a.dadv[n] || a.flux2[n] || a.dadv2[n];
    }
    a.u = c.u; a.v = c.v; a.w = c.w; a.rho = c.rho;
    a.nib = phase == PH_Y ? h->niby : h->nibx;      // (the z sweep reads zbits)
    a.nib2 = h->niby;
    a.dat = h->dat; a.datr = h->datr; a.dxte = h->dxte; a.dyte = h->dyte; a.dxtn = h->dxtn; a.dytn = h->dytn;
    a.dtime = c.dtime; a.sl = c.sl; a.accumulate = c.accumulate;
    a.rho_m1 = c.rho_m1; a.rho_r = c.rho_r;
    if (c.var == VAR_ALL) {
        if (diag) return launch_group<NT, VAR_ALL, true>(h, phase, pt, a, st);
        return launch_group<NT, VAR_ALL, false>(h, phase, pt, a, st);
    }
    if (diag) return launch_group<NT, VAR_ONE, true>(h, phase, pt, a, st);
    return launch_group<NT, VAR_ONE, false>(h, phase, pt, a, st);
}

static int run_phase_all(mom5adv_ctx *h, const SwebyCall &c, int phase, const Part &pt, cudaStream_t st, bool diag_ok = true)
{
    int rc = 0;
    for (int n0 = 0; n0 < c.ntr && !rc;) {
        const int left = c.ntr - n0;
        // groups of <= MAXNT tracers; prefer an even split (e.g. 10 = 4 + 3 + 3) over 4 + 4 + 2
        const int ngroups = (left + MAXNT - 1) / MAXNT;
        const int nt = (left + ngroups - 1) / ngroups;
        switch (nt) {
        case 1: rc = run_phase<1>(h, c, n0, phase, pt, diag_ok, st); break;
        case 2: rc = run_phase<2>(h, c, n0, phase, pt, diag_ok, st); break;
        case 3: rc = run_phase<3>(h, c, n0, phase, pt, diag_ok, st); break;
        default: rc = run_phase<4>(h, c, n0, phase, pt, diag_ok, st); break;
        }
        n0 += nt;
    }
    return rc;
}

static int zero_rings(mom5adv_ctx *h, double *const *arrs, int n, cudaStream_t st)
{
    for (int n0 = 0; n0 < n; n0 += MAXNT) {
        RingArgs<MAXNT> r{};
        for (int q = 0; q < MAXNT && n0 + q < n; q++) r.p[q] = arrs[n0 + q];
        LAUNCH(h, k_zero_ring<MAXNT>, 64, 256, 0, st, h->g, r);
    }
    return 0;
}

static bool plan_has_remote(const mom5adv_ctx *h, int flags)
{
    for (const Msg &m : h->plan[flags & 3].recvs) if (m.peer != h->rank) return true;
    for (const Msg &m : h->plan[flags & 3].sends) if (m.peer != h->rank) return true;
    return false;
}

static Part part_of(int first, int step, int count)
{
    Part p;
    p.first = first; p.step = step; p.count = count;
    return p;
}

// z tiles: tile t covers the data-domain columns t*ZBX .. t*ZBX + ZBX-1 (column 0, the west halo, is an idle lane of tile 0).
// Starting the tiles at an EVEN column makes every TMA box of the z sweep start on a 16-byte boundary, which the tensor-map
// copies of FP64 data require (measured on B200: an odd first coordinate raises "illegal instruction", tests/cuda/tma_probe.cu).
static int z_tiles(int ni) { return (ni + ZBX) / ZBX; }

// Number of trailing tiles of a dimension of n points cut into nt tiles of `size` that hold the last TWO points (the width
// of a halo strip): 1, or 2 when the last tile holds a single point.  These tiles (and tile 0) form the "edge set" of the
// comm/compute overlap: everything the strip pack reads, and every tile that reads what the strip unpack writes.
static int edge_tiles(int n, int size, int nt) { return (n - (nt - 1) * size >= 2 || nt < 2) ? 1 : 2; }

static int f_rows_for(int ni, int nj, int nk)
{   // rows per j-chunk of the fused pass: every chunk redoes the x arithmetic of 4 rows, so as large as >= ~4 waves of blocks allow
    const int nxt = (ni + 30) / 31, nxb = (nxt + FWARPS - 1) / FWARPS;
    int rows = FROWS_MAX;
    while (rows > 8 && (long long)nk * nxb * ((nj + rows - 1) / rows) < 148LL * FMINB * 4) rows /= 2;
    return rows;
}
static void pick_f_rows(mom5adv_ctx *h) { h->f_rows = f_rows_for(h->g.ni, h->g.nj, h->g.nk); }

static int y_rows_for(int ni, int nj, int nk)
{   // y-sweep chunking of the three-sweep driver (rows per j-chunk): >= ~4 waves of threads, at most YROWS_MAX rows
    const int YBX = 32 * YWARPS;
    const long long per_chunk = (long long)((ni + YBX - 1) / YBX) * YBX * nk;
    int rows = YROWS_MAX;
    while (rows > 8 && per_chunk * ((nj + rows - 1) / rows) < 148LL * 2048 * 2) rows /= 2;
    return rows;
}

// The tile sets of the comm/compute overlap, as pure functions of the block shape (mom5adv_debug_overlap_sets exposes them to the
// CPU tests, which check the two invariants the overlap rests on for thousands of shapes: every cell a strip PACK reads has been
// produced by an edge tile, and no interior tile reads a cell a strip UNPACK writes).
struct OverlapSets {
    int nzt, z_last;            // z tiles; trailing tiles in the edge set (tile 0 always is)
    int f_rows, njc_f, c_hi;    // fused pass: rows per chunk, chunks, interior chunks are 1..c_hi
    int nxt, x_last;            // three-sweep driver, x tiles of 31 cells
    int y_rows, njc_y, y_last;  // three-sweep driver, y chunks
};
static OverlapSets overlap_sets(int ni, int nj, int nk)
{
    OverlapSets o;
    o.nzt = z_tiles(ni);
    // the last two columns span two tiles when ni % ZBX == 0: z tile t covers the data-domain columns t*ZBX .. t*ZBX + ZBX-1
    o.z_last = (ni - (o.nzt - 1) * ZBX + 1 >= 2 || o.nzt < 2) ? 1 : 2;
    o.f_rows = f_rows_for(ni, nj, nk);
    o.njc_f = (nj + o.f_rows - 1) / o.f_rows;
    // chunks jc in [1, c_hi] read no halo row of the x-updated tracer (rows js-2 .. je+2 lie inside 1..nj)
    o.c_hi = std::min((nj - 2) / o.f_rows - 1, o.njc_f - 1);
    o.nxt = (ni + 30) / 31;
    o.x_last = edge_tiles(ni, 31, o.nxt);
    o.y_rows = y_rows_for(ni, nj, nk);
    o.njc_y = (nj + o.y_rows - 1) / o.y_rows;
    o.y_last = edge_tiles(nj, o.y_rows, o.njc_y);
    return o;
}

extern "C" int mom5adv_debug_overlap_sets(int ni, int nj, int nk, int out[12])
{
    if (ni < 1 || nj < 1 || nk < 1 || !out) { set_error("mom5adv_debug_overlap_sets: bad arguments"); return MOM5ADV_EINVAL; }
    const OverlapSets o = overlap_sets(ni, nj, nk);
    const int v[12] = {o.nzt, o.z_last, o.f_rows, o.njc_f, o.c_hi, o.nxt, o.x_last, o.y_rows, o.njc_y, o.y_last, ZBX, 31};
    for (int q = 0; q < 12; q++) out[q] = v[q];
    return 0;
}

// Three separate sweeps (z, x, y), the running tracer materialised between them as the reference does.
static int sweby_dev_unfused(mom5adv_ctx *h, const SwebyCall &c, cudaStream_t st)
{
    int rc;
    const Geom &g = h->g;
    const OverlapSets os = overlap_sets(g.ni, g.nj, g.nk);
    h->y_rows = os.y_rows;
    const int nxt = os.nxt, njc = os.njc_y;
    // Overlap the NCCL strip exchange with the interior tiles of the sweep that consumes it (the reference's only
    // overlap is tracer n's exchange with tracer n+1's compute, OTA:4214-4216): the exchange runs on the library's comm
    // stream while the x (y) sweep works on the tiles (j-chunks) that read no halo; the edge tiles follow.
    // The tiles that run under the exchange must touch no halo cell (the unpack writes them concurrently): x tile t reads
    // tm(31t-1 .. 31t+33), y chunk c reads rows cR-1 .. (c+1)R+2.  When the last tile (chunk) holds a single column (row)
    // the one before it reaches into the halo as well and joins the edge set.
    const int x_last = os.x_last, y_last = os.y_last;
    const bool ovx = h->overlap && plan_has_remote(h, 1) && nxt >= 3 + x_last;
    const bool ovy = h->overlap && plan_has_remote(h, 2) && njc >= 3 + y_last;
    cudaStream_t sc = h->s_comm;
    const Part all;
    CUDA_TRY(cudaEventRecord(h->ev[0], st));
    zero_rings(h, c.adv, c.ntr, st);
    if ((rc = run_phase_all(h, c, PH_Z, all, st))) return rc;
    CUDA_TRY(cudaEventRecord(h->ev[1], st));
    if (ovx) {
        CUDA_TRY(cudaEventRecord(h->ev_sync[0], st));
        CUDA_TRY(cudaStreamWaitEvent(sc, h->ev_sync[0], 0));
        if ((rc = halo_update(h, h->tmA.data(), c.ntr, 1, sc))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_sync[1], sc));
        CUDA_TRY(cudaEventRecord(h->ev[2], st));
        if ((rc = run_phase_all(h, c, PH_X, part_of(1, 1, nxt - 1 - x_last), st))) return rc;
        CUDA_TRY(cudaStreamWaitEvent(st, h->ev_sync[1], 0));
        if ((rc = run_phase_all(h, c, PH_X, part_of(0, 1, 1), st))) return rc;
        if ((rc = run_phase_all(h, c, PH_X, part_of(nxt - x_last, 1, x_last), st))) return rc;
    } else {
        if ((rc = halo_update(h, h->tmA.data(), c.ntr, 1, st))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev[2], st));
        if ((rc = run_phase_all(h, c, PH_X, all, st))) return rc;
    }
    CUDA_TRY(cudaEventRecord(h->ev[3], st));
    if (ovy) {
        CUDA_TRY(cudaEventRecord(h->ev_sync[2], st));
        CUDA_TRY(cudaStreamWaitEvent(sc, h->ev_sync[2], 0));
        if ((rc = halo_update(h, h->tmB.data(), c.ntr, 2, sc))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_sync[3], sc));
        CUDA_TRY(cudaEventRecord(h->ev[4], st));
        if ((rc = run_phase_all(h, c, PH_Y, part_of(1, 1, njc - 1 - y_last), st))) return rc;
        CUDA_TRY(cudaStreamWaitEvent(st, h->ev_sync[3], 0));
        if ((rc = run_phase_all(h, c, PH_Y, part_of(0, 1, 1), st))) return rc;
        if ((rc = run_phase_all(h, c, PH_Y, part_of(njc - y_last, 1, y_last), st))) return rc;
    } else {
        if ((rc = halo_update(h, h->tmB.data(), c.ntr, 2, st))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev[4], st));
        if ((rc = run_phase_all(h, c, PH_Y, all, st))) return rc;
    }
    CUDA_TRY(cudaEventRecord(h->ev[5], st));
    h->ev_valid = true;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// z sweep, then x and y in one pass (k_sweby_xy): the x-updated tracer exists in HBM only on the four edge rows whose
// north/south halo images the fused pass reads back.  With remote neighbours the work is spread over three streams so that
// neither the exchanges nor the small edge launches sit on the critical path:
//   st (caller)  : z interior tiles ............ | x on rows 1,2,nj-1,nj | xy interior chunks ....................... | join
//   se (edge)    : z edge tiles |                                        :            | xy edge chunks (first, last) |
//   sc (comm)    :              | E/W strip update |                     | N/S strip update |
// The edge tiles hold the columns the E/W pack reads, the edge chunks read what the N/S unpack writes; both run CONCURRENTLY
// with the interior launch (se outranks st), so their partial waves fill up with interior blocks instead of ending in a tail.
static int sweby_dev_fused(mom5adv_ctx *h, const SwebyCall &c, cudaStream_t st)
{
    int rc;
    const Geom &g = h->g;
    const OverlapSets os = overlap_sets(g.ni, g.nj, g.nk);
    h->f_rows = os.f_rows;
    const int njc = os.njc_f, nzt = os.nzt;
    // chunks jc in [1, c_hi] read no halo row of the x-updated tracer; the z tiles holding the columns 1, 2, ni-1, ni the E/W strip
    // pack reads (tile 0 and the last z_last tiles) run BEFORE the exchange starts (overlap_sets)
    const int c_hi = os.c_hi, z_last = os.z_last;
    const bool need_y = !h->plan[2].recvs.empty();
    const bool ovx = h->overlap && plan_has_remote(h, 1) && nzt >= 3 + z_last;
    const bool ovy = h->overlap && plan_has_remote(h, 2) && c_hi >= 1;
    cudaStream_t sc = h->s_comm, se = h->s_edge;
    const Part all;
    CUDA_TRY(cudaEventRecord(h->ev[0], st));
    if (c.adv) zero_rings(h, c.adv, c.ntr, st);
    if (ovx) {
        CUDA_TRY(cudaEventRecord(h->ev_edge[0], st));                      // everything the caller queued before this call
        CUDA_TRY(cudaStreamWaitEvent(se, h->ev_edge[0], 0));
        if ((rc = run_phase_all(h, c, PH_Z, part_of(0, 1, 1), se))) return rc;
        if ((rc = run_phase_all(h, c, PH_Z, part_of(nzt - z_last, 1, z_last), se))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_sync[0], se));
        CUDA_TRY(cudaStreamWaitEvent(sc, h->ev_sync[0], 0));
        if ((rc = halo_update(h, h->tmA.data(), c.ntr, 1, sc))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_sync[1], sc));
        if ((rc = run_phase_all(h, c, PH_Z, part_of(1, 1, nzt - 1 - z_last), st))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev[1], st));
        CUDA_TRY(cudaStreamWaitEvent(st, h->ev_sync[1], 0));              // (implies the edge tiles: the exchange waited for them)
    } else {
        if ((rc = run_phase_all(h, c, PH_Z, all, st))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev[1], st));
        if ((rc = halo_update(h, h->tmA.data(), c.ntr, 1, st))) return rc;
    }
    CUDA_TRY(cudaEventRecord(h->ev[2], st));
    if (need_y) {
        Part south, north;
        south.row_first = 1; south.row_last = std::min(2, g.nj);
        north.row_first = std::max(3, g.nj - 1); north.row_last = g.nj;
        if ((rc = run_phase_all(h, c, PH_X, south, st, false))) return rc;
        if ((rc = run_phase_all(h, c, PH_X, north, st, false))) return rc;
    }
    CUDA_TRY(cudaEventRecord(h->ev[3], st));
    if (need_y && ovy) {
        CUDA_TRY(cudaEventRecord(h->ev_sync[2], st));
        CUDA_TRY(cudaStreamWaitEvent(sc, h->ev_sync[2], 0));
        if ((rc = halo_update(h, h->tmB.data(), c.ntr, 2, sc))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_sync[3], sc));
        CUDA_TRY(cudaEventRecord(h->ev[4], st));
        if ((rc = run_phase_all(h, c, PH_XY, part_of(1, 1, c_hi), st))) return rc;
        CUDA_TRY(cudaStreamWaitEvent(se, h->ev_sync[3], 0));              // N/S strips in place (and, through ev_sync[2], all of z and x)
        if ((rc = run_phase_all(h, c, PH_XY, part_of(0, 1, 1), se))) return rc;
        if ((rc = run_phase_all(h, c, PH_XY, part_of(c_hi + 1, 1, njc - 1 - c_hi), se))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_edge[1], se));
        CUDA_TRY(cudaStreamWaitEvent(st, h->ev_edge[1], 0));              // join
    } else {
        if (need_y && (rc = halo_update(h, h->tmB.data(), c.ntr, 2, st))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev[4], st));
        if ((rc = run_phase_all(h, c, PH_XY, all, st))) return rc;
    }
    CUDA_TRY(cudaEventRecord(h->ev[5], st));
    h->ev_valid = true;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int sweby_dev(mom5adv_ctx *h, const SwebyCall &c, cudaStream_t st)
{
    if (c.ntr < 1 || c.ntr > h->ntr_max) { set_error("sweby: ntr=%d outside 1..%d", c.ntr, h->ntr_max); return MOM5ADV_EINVAL; }
    return h->fuse ? sweby_dev_fused(h, c, st) : sweby_dev_unfused(h, c, st);
}

extern "C" int mom5adv_sweby_all_dev(mom5adv_handle h, int ntr, double dtime, const double *const *T, double *const *th,
                                     double *const *adv, const double *u, const double *v, const double *w, const double *rho,
                                     double *const *fx, double *const *fy, double *const *fz, double *const *ax,
                                     double *const *ay, double *const *az, void *stream)
{
    if (!h || !T || !th || !adv || !u || !v || !w || !rho) { set_error("mom5adv_sweby_all_dev: null argument"); return MOM5ADV_EINVAL; }
    SwebyCall c{ntr, VAR_ALL, dtime, 1.0, T, th, adv, u, v, w, rho, fx, fy, fz, ax, ay, az, 1};
    return sweby_dev(h, c, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// host-pointer mode: device mirrors + H2D / D2H around the _dev path
// ------------------------------------------------------------------------------------------------
static int mirror(mom5adv_ctx *h, size_t idx, double **out)
{
    while (h->hm.size() <= idx) {
        double *p = nullptr;
        CUDA_TRY(cudaMalloc(&p, n3(h) * sizeof(double)));
        h->hm.push_back(p);
    }
    *out = h->hm[idx];
    return 0;
}
static int mirror_w(mom5adv_ctx *h, double **out)
{
    if (!h->hm_w) CUDA_TRY(cudaMalloc(&h->hm_w, ((size_t)h->g.slab * (h->g.nk + 1)) * sizeof(double)));
    *out = h->hm_w;
    return 0;
}
#define H2D(dst, src, n) do { CUDA_TRY(cudaMemcpyAsync(dst, src, (n) * sizeof(double), cudaMemcpyHostToDevice, st)); h->h2d_bytes += (long long)((n) * sizeof(double)); } while (0)
#define D2H(dst, src, n) do { CUDA_TRY(cudaMemcpyAsync(dst, src, (n) * sizeof(double), cudaMemcpyDeviceToHost, st)); h->d2h_bytes += (long long)((n) * sizeof(double)); } while (0)

// Page-lock a caller array once (the model's arrays live for the whole run): with pageable memory every cudaMemcpyAsync is
// staged through a driver bounce buffer and effectively synchronous, so neither copy engine overlaps anything.  Failure to
// register (e.g. the array is already registered, or memory limits) is not an error: the copy then simply runs staged.
static void pin_host(mom5adv_ctx *h, const void *p, size_t bytes)
{
    if (!h->pin || !p || h->pinned.count(p)) return;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) { h->pinned[p] = 0; return; }
    cudaGetLastError();
    if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) h->pinned[p] = bytes;
    else { cudaGetLastError(); h->pinned[p] = 0; }
}

// Host-pointer sweby_all as a pipeline over j-BANDS (one band = one j-chunk of the fused pass): while band b is on its way
// up the PCIe link, the SMs run the z sweep of band b-1 and the fused x/y pass of the chunk below it, and the results of
// finished chunks travel down -- both copy engines stay busy for the whole call instead of idling during the upload of
// the velocity fields (tracer-wise pipeline) and the download of the last tracer.  Single-rank layouts only: the E/W
// strip update (cyclic wrap) is redone after every band, which across ranks would need matching band counts.
//   chunk c (rows c*R+1 .. (c+1)*R) reads the z-updated tracer on rows <= (c+1)*R + 2  ->  it runs after band c+1;
//   chunks touching a halo row of the x-updated tracer (the first, the last one or two) run at the end, after the edge-row
//   x sweep and the N/S strip update, exactly as in sweby_dev_fused.
static int band_copy(mom5adv_ctx *h, double *dst, const double *src, int ja, int jb, int nlev, cudaMemcpyKind kind, cudaStream_t st)
{
    const Geom &g = h->g;
    const size_t pitch = (size_t)g.slab * sizeof(double), ofs = (size_t)ja * g.nxd, width = (size_t)(jb - ja + 1) * g.nxd * sizeof(double);
    CUDA_TRY(cudaMemcpy2DAsync(dst + ofs, pitch, src + ofs, pitch, width, (size_t)nlev, kind, st));
    (kind == cudaMemcpyHostToDevice ? h->h2d_bytes : h->d2h_bytes) += (long long)(width * nlev);
    return 0;
}

// th_tendency(i,j,k) += adv(i,j,k) on the compute-domain points of rows ja..jb, on the host (OTA:4420-4424: the reference touches
// th_tendency on the compute domain only; halo points must keep their bits, so they are not even added a zero)
static void host_accumulate(const Geom &g, double *th, const double *adv, int ja, int jb, int nthreads)
{
    ja = std::max(ja, 1); jb = std::min(jb, g.nj);
    if (jb < ja) return;
    const int nrow = (jb - ja + 1) * g.nk;                 // (row, level) pairs
    nthreads = std::max(1, std::min(nthreads, nrow));
    auto work = [&](int t) {
        for (int q = t; q < nrow; q += nthreads) {
            const int k = q / (jb - ja + 1) + 1, j = ja + q % (jb - ja + 1);
            double *a = th + d3(g, 1, j, k);
            const double *b = adv + d3(g, 1, j, k);
            for (int i = 0; i < g.ni; i++) a[i] = a[i] + b[i];
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto &th_ : pool) th_.join();
}

static int sweby_all_banded(mom5adv_ctx *h, int ntr, double dtime, const double *const *T, double *const *th, double *const *adv,
                            const double *u, const double *v, const double *w, const double *rho, double *du, double *dv,
                            double *dw, double *dr, double *const *dT, double *const *dadv)
{
    const Geom &g = h->g;
    int rc;
    pick_f_rows(h);
    const int R = h->f_rows, njc = (g.nj + R - 1) / R;
    const int c_hi = std::min((g.nj - 2) / R - 1, njc - 1);
    cudaStream_t st = h->stream, up = h->s_up, down = h->s_down;
    while ((int)h->ev_up.size() < njc + 1) {
        cudaEvent_t e1, e2;
        CUDA_TRY(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
        h->ev_up.push_back(e1); h->ev_done.push_back(e2);
    }
    while ((int)h->ev_dl.size() < njc + 2) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->ev_dl.push_back(e);
    }
    // where the downloaded adv_tendency lands on the host: the caller's array, or pinned staging owned by the library
    std::vector<double *> hadv(ntr);
    for (int n = 0; n < ntr; n++) {
        if (adv && adv[n]) { hadv[n] = adv[n]; continue; }
        if ((int)h->hstage.size() <= n || h->hstage_elems < n3(h)) {
            for (double *p : h->hstage) cudaFreeHost(p);
            h->hstage.assign(ntr, nullptr);
            for (int q = 0; q < ntr; q++) CUDA_TRY(cudaMallocHost(&h->hstage[q], n3(h) * sizeof(double)));
            h->hstage_elems = n3(h);
        }
        hadv[n] = h->hstage[n];
    }
    std::vector<const double *> cT(dT, dT + ntr);
    // the device forms adv_tendency only (accumulate = 0: th_tendency is neither read nor written on the device, and never crosses the link)
    SwebyCall c{ntr, VAR_ALL, dtime, 1.0, cT.data(), nullptr, dadv, du, dv, dw, dr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    zero_rings(h, dadv, ntr, st);
    struct Dl { int ja, jb, ev; };
    std::vector<Dl> dls;
    auto download = [&](int ja, int jb, int evi) -> int {          // rows ja..jb of adv once the compute stream got here
        CUDA_TRY(cudaEventRecord(h->ev_done[evi], st));
        CUDA_TRY(cudaStreamWaitEvent(down, h->ev_done[evi], 0));
        for (int n = 0; n < ntr; n++)
            if ((rc = band_copy(h, hadv[n], dadv[n], ja, jb, g.nk, cudaMemcpyDeviceToHost, down))) return rc;
        const int e = (int)dls.size();
        CUDA_TRY(cudaEventRecord(h->ev_dl[e], down));
        dls.push_back({ja, jb, e});
        return 0;
    };
    for (int b = 0; b < njc; b++) {
        const int ja = (b == 0) ? 0 : b * R + 1, jb = (b == njc - 1) ? g.nj + 1 : (b + 1) * R;
        const cudaMemcpyKind H = cudaMemcpyHostToDevice;
        if ((rc = band_copy(h, du, u, ja, jb, g.nk, H, up)) || (rc = band_copy(h, dv, v, ja, jb, g.nk, H, up)) ||
            (rc = band_copy(h, dr, rho, ja, jb, g.nk, H, up)) || (rc = band_copy(h, dw, w, ja, jb, g.nk + 1, H, up))) return rc;
        for (int n = 0; n < ntr; n++)
            if ((rc = band_copy(h, dT[n], T[n], ja, jb, g.nk, H, up))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_up[b], up));
        CUDA_TRY(cudaStreamWaitEvent(st, h->ev_up[b], 0));
        Part zp;
        zp.row_first = std::max(ja, 1); zp.row_last = std::min(jb, g.nj);
        if ((rc = run_phase_all(h, c, PH_Z, zp, st))) return rc;
        if ((rc = halo_update(h, h->tmA.data(), ntr, 1, st))) return rc;   // E/W strips (all rows; rows of later bands are redone)
        const int ch = b - 1;                                             // the chunk that has just become computable
        if (ch >= 1 && ch <= c_hi) {
            if ((rc = run_phase_all(h, c, PH_XY, part_of(ch, 1, 1), st))) return rc;
            if ((rc = download(ch * R + 1, (ch + 1) * R, b))) return rc;
        }
    }
    // edge chunks: x sweep on the four edge rows, N/S strip update of the x-updated tracer, then the chunks that read it
    if (!h->plan[2].recvs.empty()) {
        Part south, north;
        south.row_first = 1; south.row_last = std::min(2, g.nj);
        north.row_first = std::max(3, g.nj - 1); north.row_last = g.nj;
        if ((rc = run_phase_all(h, c, PH_X, south, st, false)) || (rc = run_phase_all(h, c, PH_X, north, st, false))) return rc;
        if ((rc = halo_update(h, h->tmB.data(), ntr, 2, st))) return rc;
    }
    const int first_tail = std::max(c_hi + 1, 1);
    if ((rc = run_phase_all(h, c, PH_XY, part_of(0, 1, 1), st))) return rc;
    if (njc > first_tail && (rc = run_phase_all(h, c, PH_XY, part_of(first_tail, 1, njc - first_tail), st))) return rc;
    if (njc > 1) {
        if ((rc = download(0, R, njc))) return rc;                         // chunk 0 with the south halo row
        if ((rc = download(first_tail * R + 1, g.nj + 1, njc))) return rc; // remaining rows up to the north halo row, in one go
    } else if ((rc = download(0, g.nj + 1, njc))) return rc;
    h->ev_valid = false;
    CUDA_TRY(cudaGetLastError());
    // everything is enqueued; this thread now follows the downloads and accumulates band by band while the link and the SMs go on
    for (const Dl &d : dls) {
        CUDA_TRY(cudaEventSynchronize(h->ev_dl[d.ev]));
        if (th)
            for (int n = 0; n < ntr; n++)
                if (th[n]) host_accumulate(g, th[n], hadv[n], d.ja, d.jb, h->host_threads);
    }
    CUDA_TRY(cudaStreamSynchronize(down));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int mom5adv_sweby_all(mom5adv_handle h, int ntr, double dtime, const double *const *T, double *const *th,
                                 double *const *adv, const double *u, const double *v, const double *w, const double *rho,
                                 double *const *fx, double *const *fy, double *const *fz, double *const *ax, double *const *ay,
                                 double *const *az)
{
    if (!h || !T || !u || !v || !w || !rho) { set_error("mom5adv_sweby_all: null argument"); return MOM5ADV_EINVAL; }
    if (!th && !adv) { set_error("mom5adv_sweby_all: neither th_tendency nor adv_tendency is wanted"); return MOM5ADV_EINVAL; }
    if (ntr < 1 || ntr > h->ntr_max) { set_error("mom5adv_sweby_all: ntr=%d outside 1..%d", ntr, h->ntr_max); return MOM5ADV_EINVAL; }
    cudaStream_t st = h->stream;
    const size_t N = n3(h);
    int rc;
    h->h2d_bytes = h->d2h_bytes = 0;
    {   // page-lock the caller's arrays (once per array: the cache is keyed by address)
        const size_t B = N * sizeof(double);
        pin_host(h, u, B); pin_host(h, v, B); pin_host(h, rho, B); pin_host(h, w, (size_t)h->g.slab * (h->g.nk + 1) * sizeof(double));
        for (int n = 0; n < ntr; n++) {
            pin_host(h, T[n], B);
            if (th && th[n]) pin_host(h, th[n], B);
            if (adv && adv[n]) pin_host(h, adv[n], B);
        }
    }
    size_t slot = 0;
    double *du, *dv, *dw, *dr;
    if ((rc = mirror(h, slot++, &du)) || (rc = mirror(h, slot++, &dv)) || (rc = mirror(h, slot++, &dr)) || (rc = mirror_w(h, &dw))) return rc;
    std::vector<double *> dT(ntr), dth(ntr), dadv(ntr);
    std::vector<const double *> cT(ntr);
    for (int n = 0; n < ntr; n++) {
        if ((rc = mirror(h, slot++, &dT[n])) || (rc = mirror(h, slot++, &dth[n])) || (rc = mirror(h, slot++, &dadv[n]))) return rc;
        cT[n] = dT[n];
    }
    // optional diagnostics share per-kind mirrors
    double *const *hd[6] = {fx, fy, fz, ax, ay, az};
    std::vector<std::vector<double *>> dd(6);
    for (int q = 0; q < 6; q++)
        if (hd[q]) {
            dd[q].assign(ntr, nullptr);
            for (int n = 0; n < ntr; n++)
                if (hd[q][n]) {
                    if ((rc = mirror(h, slot++, &dd[q][n]))) return rc;
                    H2D(dd[q][n], hd[q][n], N);   // points outside the loop ranges keep the caller's values
                }
        }
    bool any_diag = false;
    for (int q = 0; q < 6; q++) any_diag |= (hd[q] != nullptr);
    pick_f_rows(h);
    if (!any_diag && h->banded && h->fuse && !plan_has_remote(h, 1) && !plan_has_remote(h, 2) &&
        (h->g.nj + h->f_rows - 1) / h->f_rows >= 4)
        return sweby_all_banded(h, ntr, dtime, T, th, adv, u, v, w, rho, du, dv, dw, dr, dT.data(), dadv.data());
    if (!th) { set_error("mom5adv_sweby_all: th_tendency may only be omitted on the banded single-rank pipeline"); return MOM5ADV_EUNSUP; }
    if (!any_diag && ntr > 1) {
        // Pipelined over tracers: the copy engines run in both directions while the SMs work on another tracer.
        //   up stream  : u, v, w, rho, then (T_n, th_n) for n = 0, 1, ...
        //   compute    : tracer n as soon as its inputs have landed (results do not depend on the grouping of tracers)
        //   down stream: th_n, adv_n as soon as tracer n is done -- overlapping the upload of tracer n+1
        cudaStream_t up = h->s_up, down = h->s_down;
        {
            cudaStream_t st = up;
            H2D(du, u, N); H2D(dv, v, N); H2D(dr, rho, N);
            H2D(dw, w, (size_t)h->g.slab * (h->g.nk + 1));
        }
        for (int n = 0; n < ntr; n++) {
            {
                cudaStream_t st = up;
                H2D(dT[n], T[n], N); H2D(dth[n], th[n], N);
            }
            CUDA_TRY(cudaEventRecord(h->ev_up[n], up));
            CUDA_TRY(cudaStreamWaitEvent(st, h->ev_up[n], 0));
            const double *Tn[1] = {dT[n]};
            double *thn[1] = {dth[n]}, *advn[1] = {dadv[n]};
            SwebyCall c1{1, VAR_ALL, dtime, 1.0, Tn, thn, advn, du, dv, dw, dr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1};
            if ((rc = sweby_dev(h, c1, st))) return rc;
            CUDA_TRY(cudaEventRecord(h->ev_done[n], st));
            CUDA_TRY(cudaStreamWaitEvent(down, h->ev_done[n], 0));
            {
                cudaStream_t st = down;
                D2H(th[n], dth[n], N);
                if (adv && adv[n]) D2H(adv[n], dadv[n], N);
            }
        }
        CUDA_TRY(cudaStreamSynchronize(down));
        CUDA_TRY(cudaStreamSynchronize(st));
        return 0;
    }
    H2D(du, u, N); H2D(dv, v, N); H2D(dr, rho, N);
    H2D(dw, w, (size_t)h->g.slab * (h->g.nk + 1));
    for (int n = 0; n < ntr; n++) { H2D(dT[n], T[n], N); H2D(dth[n], th[n], N); }
    auto arr = [&](int q) -> double *const * { return hd[q] ? dd[q].data() : nullptr; };
    SwebyCall c{ntr, VAR_ALL, dtime, 1.0, cT.data(), dth.data(), dadv.data(), du, dv, dw, dr, arr(0), arr(1), arr(2), arr(3), arr(4), arr(5), 1};
    if ((rc = sweby_dev(h, c, st))) return rc;
    for (int n = 0; n < ntr; n++) {
        D2H(th[n], dth[n], N);
        if (adv && adv[n]) D2H(adv[n], dadv[n], N);
        for (int q = 0; q < 6; q++)
            if (hd[q] && hd[q][n]) D2H(hd[q][n], dd[q][n], N);
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// fold-line fix of the quicker fluxes (OTA:2640), possibly across ranks
// ------------------------------------------------------------------------------------------------
static int fold_line_fix(mom5adv_ctx *h, double *fy, cudaStream_t st)
{
    if (h->iy != h->py - 1) return 0;                      // only the top row of ranks touches the fold
    const int middle = (1 + h->ni_g) / 2 + 1;
    const int nk = h->g.nk;
    // enumerate, for every top-row rank r, the destination segments (global i >= middle) split by the owner of the mirror image
    struct Seg { int dst_rank, src_rank, dst_i1g, src_i0g, w; };
    std::vector<Seg> segs;
    for (int bx = 0; bx < h->px; bx++) {
        int a = std::max(h->ibeg[bx], middle), b = h->iend[bx];
        while (a <= b) {
            // destination run [a, e] whose mirror [ni_g+1-e, ni_g+1-a] lies in ONE source block
            const int sx = find_div(h->ibeg, h->iend, h->ni_g + 1 - a);
            const int e = std::min(b, h->ni_g + 1 - h->ibeg[sx]);
            segs.push_back({bx + h->px * h->iy, sx + h->px * h->iy, e, h->ni_g + 1 - e, e - a + 1});
            a = e + 1;
        }
    }
    FoldArgs L{}, S{}, R{};
    std::vector<int> speer, rpeer;
    L.fy = S.fy = R.fy = fy;
    auto add = [&](FoldArgs &A, int src_i0, int dst_i1, int w) {
        if (A.nseg >= FOLD_MAXSEG) return false;
        A.s[A.nseg] = FoldSeg{src_i0, dst_i1, w, A.start[A.nseg]};
        A.start[A.nseg + 1] = A.start[A.nseg] + (long long)w * nk;
        A.nseg++;
        return true;
    };
    const int my_i0g = h->ibeg[h->ix];
    bool ok = true;
    // deterministic order on both sides: by (dst_rank, src_rank, position)
    for (const Seg &sg : segs) {
        if (sg.dst_rank == h->rank && sg.src_rank == h->rank) ok &= add(L, sg.src_i0g - my_i0g + 1, sg.dst_i1g - my_i0g + 1, sg.w);
        else if (sg.src_rank == h->rank) { ok &= add(S, sg.src_i0g - my_i0g + 1, 0, sg.w); speer.push_back(sg.dst_rank); }
        else if (sg.dst_rank == h->rank) { ok &= add(R, 0, sg.dst_i1g - my_i0g + 1, sg.w); rpeer.push_back(sg.src_rank); }
    }
    if (!ok) { set_error("fold-line fix: too many segments"); return MOM5ADV_EINVAL; }
    auto nblk = [](long long n) { return (int)std::min<long long>((n + 255) / 256, 148 * 8); };
    if (L.nseg) LAUNCH(h, k_fold_line<0>, nblk(L.start[L.nseg]), 256, 0, st, h->g, L);
    if (S.nseg || R.nseg) {
        if (!h->comm) { set_error("fold-line fix needs a communicator"); return MOM5ADV_EINVAL; }
        const size_t need = (size_t)std::max(S.start[S.nseg], R.start[R.nseg]);
        if (need > h->bufcap) {
            CUDA_TRY(cudaStreamSynchronize(st));
            if (h->sendbuf) cudaFree(h->sendbuf);
            if (h->recvbuf) cudaFree(h->recvbuf);
            CUDA_TRY(cudaMalloc(&h->sendbuf, need * sizeof(double)));
            CUDA_TRY(cudaMalloc(&h->recvbuf, need * sizeof(double)));
            h->bufcap = need;
        }
        S.buf = h->sendbuf;
        R.buf = h->recvbuf;
        if (S.nseg) LAUNCH(h, k_fold_line<1>, nblk(S.start[S.nseg]), 256, 0, st, h->g, S);
        NCCL_NEED();
        ncclComm_t comm = (ncclComm_t)h->comm->nccl;
        NCCL_TRY(N->GroupStart());
        for (int m = 0; m < S.nseg; m++)
            NCCL_TRY(N->Send(h->sendbuf + S.start[m], (size_t)(S.start[m + 1] - S.start[m]), ncclDouble, speer[m], comm, st));
        for (int m = 0; m < R.nseg; m++)
            NCCL_TRY(N->Recv(h->recvbuf + R.start[m], (size_t)(R.start[m + 1] - R.start[m]), ncclDouble, rpeer[m], comm, st));
        NCCL_TRY(N->GroupEnd());
        if (R.nseg) LAUNCH(h, k_fold_line<2>, nblk(R.start[R.nseg]), 256, 0, st, h->g, R);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// horz_advect_tracer / vert_advect_tracer (one tracer)
// ------------------------------------------------------------------------------------------------
// quicker / upwind horizontal arm as ONE pass (k_horz_fused); tq = h->tmA[0] must already hold tracer_quick for quicker
static int horz_fused_dev(mom5adv_ctx *h, bool quicker, const double *Tm1, const double *Tt, const double *tlimit, int limit,
                          const double *u, const double *v, double *th, double *wrk1, double *fx, double *fy, cudaStream_t st)
{
    const Geom &g = h->g;
    HorzIn a{h->tmask, h->dyte, h->dxtn, Tm1, Tt, h->tmA[0], tlimit, u, v, h->mask, limit};
    const bool flux = fx || fy;
    // flux_x = flux_y = 0 outside the loop ranges (OTA:2568-2569) -- only when the caller wants the arrays at all
    if (fx) CUDA_TRY(cudaMemsetAsync(fx, 0, n3(h) * sizeof(double), st));
    if (fy) CUDA_TRY(cudaMemsetAsync(fy, 0, n3(h) * sizeof(double), st));
    double *wz[1] = {wrk1};
    zero_rings(h, wz, 1, st);                                  // Tracer%wrk1 = 0 on the data domain (OTA:1925-1931)
    // the folded top row is on this rank as a whole (layout_x = 1): its eastern half takes minus the mirror column's north flux
    const int fold_mid = (h->tripolar && h->iy == h->py - 1 && h->px == 1) ? (1 + h->ni_g) / 2 + 1 : 0;
    const dim3 grid((g.ni + 30) / 31, (g.nj + HFW - 1) / HFW, g.nk), block(32, HFW);
    if (quicker) {
        if (flux) LAUNCH(h, (k_horz_fused<true, true>), grid, block, 0, st, g, h->qw, a, h->datr, fold_mid, th, wrk1, fx, fy);
        else LAUNCH(h, (k_horz_fused<true, false>), grid, block, 0, st, g, h->qw, a, h->datr, fold_mid, th, wrk1, fx, fy);
    } else {
        if (flux) LAUNCH(h, (k_horz_fused<false, true>), grid, block, 0, st, g, h->qw, a, h->datr, 0, th, wrk1, fx, fy);
        else LAUNCH(h, (k_horz_fused<false, false>), grid, block, 0, st, g, h->qw, a, h->datr, 0, th, wrk1, fx, fy);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int mom5adv_horz_dev(mom5adv_handle h, int scheme, double dtime, const double *Tm1, const double *Tt,
                                const double *tlimit, int limit_with_upwind, const double *u, const double *v, const double *w,
                                const double *rho, double *th, double *wrk1, double *fx, double *fy, double *fz, void *stream)
{
    if (!h || !Tm1 || !u || !v || !th || !wrk1) { set_error("mom5adv_horz_dev: null argument"); return MOM5ADV_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const Geom &g = h->g;
    switch (scheme) {
    case MOM5ADV_ADVECT_MDFL_SWEBY:
    case MOM5ADV_ADVECT_DST_LINEAR: {
        if (!w || !rho) { set_error("mom5adv_horz_dev: sweby needs wrho_bt and rho_dzt"); return MOM5ADV_EINVAL; }
        // flux_x = flux_y = 0 (OTA:3837-3838)
        if (fx) CUDA_TRY(cudaMemsetAsync(fx, 0, n3(h) * sizeof(double), st));
        if (fy) CUDA_TRY(cudaMemsetAsync(fy, 0, n3(h) * sizeof(double), st));
        const double *Ts[1] = {Tm1};
        double *ths[1] = {th}, *advs[1] = {wrk1}, *fxs[1] = {fx}, *fys[1] = {fy}, *fzs[1] = {fz};
        SwebyCall c{1, VAR_ONE, dtime, scheme == MOM5ADV_ADVECT_MDFL_SWEBY ? 1.0 : 0.0, Ts, ths, advs, u, v, w, rho,
                    fx ? fxs : nullptr, fy ? fys : nullptr, fz ? fzs : nullptr, nullptr, nullptr, nullptr, 1};
        return sweby_dev(h, c, st);
    }
    case MOM5ADV_ADVECT_MDFL_SWEBY_TEST:
    case MOM5ADV_ADVECT_DST_LINEAR_TEST: {
        if (!w || !rho) { set_error("mom5adv_horz_dev: sweby_test needs wrho_bt and rho_dzt"); return MOM5ADV_EINVAL; }
        int rc;
        if (!h->st_ms) {
            CUDA_TRY(cudaMalloc(&h->st_ms, nh2(h) * sizeof(double)));
            CUDA_TRY(cudaMemset(h->st_ms, 0, nh2(h) * sizeof(double)));   // wall halos stay 0, as tmA / tmB (OTA:3500-3502)
        }
        STArgs a{};
        a.T = Tm1; a.u = u; a.v = v; a.w = w; a.rho = rho; a.mask = h->mask;
        a.dat = h->dat; a.datr = h->datr; a.dyte = h->dyte; a.dxtn = h->dxtn;
        a.tr = h->tmA[0]; a.tms = h->tmB[0]; a.ms = h->st_ms;
        a.fx = fx; a.fy = fy; a.fz = fz; a.th = th; a.wrk1 = wrk1;
        a.dtime = dtime; a.sl = (scheme == MOM5ADV_ADVECT_MDFL_SWEBY_TEST) ? 1.0 : 0.0;
        if (!a.fx && (rc = mirror(h, 0, &a.fx))) return rc;    // the reference's module-level flux_x / flux_y work arrays
        if (!a.fy && (rc = mirror(h, 1, &a.fy))) return rc;
        CUDA_TRY(cudaMemsetAsync(a.fx, 0, n3(h) * sizeof(double), st));   // flux_x = flux_y = 0 (OTA:3503-3504)
        CUDA_TRY(cudaMemsetAsync(a.fy, 0, n3(h) * sizeof(double), st));
        double *wz[1] = {wrk1};
        zero_rings(h, wz, 1, st);                                          // Tracer%wrk1 = 0 on the data domain (OTA:1925-1931)
        double *f3[3] = {a.tr, a.tms, a.ms};
        const int nbx = (g.ni + 127) / 128, nbx1 = (g.ni + 1 + 127) / 128;
        LAUNCH(h, k_st_z, dim3(nbx, g.nj), 128, 0, st, g, a);
        if ((rc = halo_update(h, f3, 3, 1, st))) return rc;                // XUPDATE of the three fields (OTA:3582-3584)
        LAUNCH(h, k_st_xflux, dim3(nbx1, g.nj, g.nk), 128, 0, st, g, a);
        LAUNCH(h, k_st_xupd, dim3(nbx, g.nj, g.nk), 128, 0, st, g, a);
        if ((rc = halo_update(h, f3, 3, 2, st))) return rc;                // YUPDATE (OTA:3651-3653)
        LAUNCH(h, k_st_yflux, dim3(nbx, g.nj + 1, g.nk), 128, 0, st, g, a);
        LAUNCH(h, k_st_yupd, dim3(nbx, g.nj, g.nk), 128, 0, st, g, a);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    case MOM5ADV_ADVECT_MDPPM: {
        if (!w || !rho) { set_error("mom5adv_horz_dev: mdppm needs wrho_bt and rho_dzt"); return MOM5ADV_EINVAL; }
        if (h->ppm_hlimiter < 1 || h->ppm_hlimiter > 3) {   // OTA:6146-6148
            set_error("mom5adv_horz_dev: must choose ppm_hlimiter=1,2 or 3");
            return MOM5ADV_EINVAL;
        }
        int rc;
        {
            const bool src_x = h->cyclic_x || h->px > 1, src_y = h->cyclic_y || h->py > 1 || h->tripolar;
            if ((src_x && g.ni < 4) || (src_y && g.nj < 4)) {
                set_error("mom5adv_horz_dev: mdppm needs a local block of at least 4x4 cells for its width-4 halo (have %dx%d)", g.ni, g.nj);
                return MOM5ADV_EUNSUP;
            }
        }
        const size_t n4 = (size_t)g.s4 * g.nk;
        if (!h->pp_ready) {
            if (!h->pp_tr) CUDA_TRY(cudaMalloc(&h->pp_tr, n4 * sizeof(double)));
            if (!h->pp_da) CUDA_TRY(cudaMalloc(&h->pp_da, n4 * sizeof(double)));
            if (!h->pp_m4) CUDA_TRY(cudaMalloc(&h->pp_m4, n4 * sizeof(double)));
            // mdppm_init (OTA:1714-1726): tmask_mdppm = 0; compute domain := Grd%tmask; full halo-4 update
            CUDA_TRY(cudaMemsetAsync(h->pp_m4, 0, n4 * sizeof(double), st));
            LAUNCH(h, k_d1_to_h4, dim3((g.ni + 127) / 128, g.nj, g.nk), 128, 0, st, g, h->tmask, h->pp_m4);
            double *m1[1] = {h->pp_m4};
            if ((rc = halo_update_l<2>(h, m1, 1, 3, st))) return rc;
            h->pp_ready = true;   // only now: a failed mask update must not leave later calls with an unfilled tmask_mdppm
        }
        PPMArgs a{};
        a.T = Tm1; a.u = u; a.v = v; a.w = w; a.rho = rho; a.m4 = h->pp_m4; a.tmask = h->tmask;
        a.dat = h->dat; a.datr = h->datr; a.dxte = h->dxte; a.dyte = h->dyte; a.dxtn = h->dxtn; a.dytn = h->dytn;
        a.tr = h->pp_tr; a.da = h->pp_da; a.fx = fx; a.fy = fy; a.fz = fz; a.th = th; a.wrk1 = wrk1;
        a.dtime = dtime; a.limiter = h->ppm_hlimiter;
        if (!a.fx && (rc = mirror(h, 0, &a.fx))) return rc;    // the reference's module-level flux work arrays
        if (!a.fy && (rc = mirror(h, 1, &a.fy))) return rc;
        if (!a.fz) {
            if (!h->pp_fz) CUDA_TRY(cudaMalloc(&h->pp_fz, n3(h) * sizeof(double)));
            a.fz = h->pp_fz;
        }
        CUDA_TRY(cudaMemsetAsync(a.tr, 0, n4 * sizeof(double), st));       // tracer_mdppm = 0.0 (OTA:6022)
        CUDA_TRY(cudaMemsetAsync(a.fx, 0, n3(h) * sizeof(double), st));    // flux_x = flux_y = 0.0 (OTA:6023-6024)
        CUDA_TRY(cudaMemsetAsync(a.fy, 0, n3(h) * sizeof(double), st));
        double *wz[1] = {wrk1}, *t1[1] = {a.tr};
        zero_rings(h, wz, 1, st);                                           // Tracer%wrk1 = 0 on the data domain (OTA:1925-1931)
        const int nbx = (g.ni + 127) / 128;
        const dim3 cells(nbx, g.nj, g.nk);
        LAUNCH(h, k_ppm_zslope, cells, 128, 0, st, g, a);
        LAUNCH(h, k_ppm_zflux, cells, 128, 0, st, g, a);
        LAUNCH(h, k_ppm_zupd, cells, 128, 0, st, g, a);
        if ((rc = halo_update_l<2>(h, t1, 1, 1, st))) return rc;           // XUPDATE, halo 4 (OTA:6201)
        LAUNCH(h, k_ppm_hslope<0>, dim3((g.ni + 4 + 127) / 128, g.nj, g.nk), 128, 0, st, g, a);
        LAUNCH(h, k_ppm_hflux<0>, dim3((g.ni + 1 + 127) / 128, g.nj, g.nk), 128, 0, st, g, a);
        LAUNCH(h, k_ppm_xupd, cells, 128, 0, st, g, a);
        if ((rc = halo_update_l<2>(h, t1, 1, 2, st))) return rc;           // YUPDATE (OTA:6333)
        LAUNCH(h, k_ppm_hslope<1>, dim3(nbx, g.nj + 4, g.nk), 128, 0, st, g, a);
        LAUNCH(h, k_ppm_hflux<1>, dim3(nbx, g.nj + 1, g.nk), 128, 0, st, g, a);
        LAUNCH(h, k_ppm_yupd, cells, 128, 0, st, g, a);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    case MOM5ADV_ADVECT_UPWIND: {
        if (h->horz_fused) return horz_fused_dev(h, false, Tm1, nullptr, nullptr, 0, u, v, th, wrk1, fx, fy, st);
        double *tfx = fx, *tfy = fy;   // the reference's module-level flux_x / flux_y work arrays
        int rc;
        if (!tfx && (rc = mirror(h, 0, &tfx))) return rc;
        if (!tfy && (rc = mirror(h, 1, &tfy))) return rc;
        return horz_upwind_dev(h->g, h->tmask, h->dyte, h->dxtn, h->datr, Tm1, u, v, th, wrk1, tfx, tfy, st, &h->launches);
    }
    case MOM5ADV_ADVECT_QUICKER: {
        if (!Tt) { set_error("mom5adv_horz_dev: quicker needs T_tau"); return MOM5ADV_EINVAL; }
        if (limit_with_upwind && !tlimit) { set_error("mom5adv_horz_dev: limit_with_upwind needs tmask_limit"); return MOM5ADV_EINVAL; }
        // tracer_quick = 0; compute domain := T(taum1); full halo-2 update (OTA:2558-2566).  Wall halos of the
        // scratch are never written and stay 0.
        dim3 gb((g.ni + 127) / 128, g.nj, g.nk);
        LAUNCH(h, k_d1_to_h2, gb, 128, 0, st, g, Tm1, h->tmA[0]);
        double *f0[1] = {h->tmA[0]};
        int rc = halo_update(h, f0, 1, 3, st);
        if (rc) return rc;
        // flux, fold line and divergence in one pass unless the fold line is split over ranks (then: flux arrays + strip exchange)
        if (h->horz_fused && !(h->tripolar && h->px > 1))
            return horz_fused_dev(h, true, Tm1, Tt, tlimit, limit_with_upwind, u, v, th, wrk1, fx, fy, st);
        double *tfx = fx, *tfy = fy;   // the reference's module-level flux_x / flux_y work arrays
        if (!tfx && (rc = mirror(h, 0, &tfx))) return rc;
        if (!tfy && (rc = mirror(h, 1, &tfy))) return rc;
        if ((rc = horz_quicker_flux_dev(h->g, h->qw, h->tmask, h->mask, h->dyte, h->dxtn, Tm1, Tt, h->tmA[0], tlimit,
                                        limit_with_upwind, u, v, tfx, tfy, st, &h->launches))) { set_error("quicker flux kernel failed"); return MOM5ADV_ECUDA; }
        if (h->tripolar && (rc = fold_line_fix(h, tfy, st))) return rc;
        if ((rc = horz_div_dev(h->g, h->tmask, h->datr, tfx, tfy, th, wrk1, st, &h->launches))) { set_error("quicker divergence kernel failed"); return MOM5ADV_ECUDA; }
        return 0;
    }
    default:
        set_error("mom5adv_horz_dev: chose invalid horz advection scheme %d", scheme);   // OTA:1983-1985
        return MOM5ADV_EINVAL;
    }
}

extern "C" int mom5adv_vert_dev(mom5adv_handle h, int scheme, const double *Tm1, const double *Tt, const double *tlimit,
                                const double *w, double *th, double *wrk1, double *fz, void *stream)
{
    if (!h || !wrk1 || !th) { set_error("mom5adv_vert_dev: null argument"); return MOM5ADV_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    switch (scheme) {
    case MOM5ADV_ADVECT_MDFL_SWEBY:
    case MOM5ADV_ADVECT_DST_LINEAR:
    case MOM5ADV_ADVECT_MDPPM:
    case MOM5ADV_ADVECT_MDFL_SWEBY_TEST:
    case MOM5ADV_ADVECT_DST_LINEAR_TEST:   // three-dimensional schemes: wrk1 = 0, th unchanged (OTA:2116-2122, 2147-2155)
        CUDA_TRY(cudaMemsetAsync(wrk1, 0, n3(h) * sizeof(double), st));
        return 0;
    case MOM5ADV_ADVECT_UPWIND:
        if (!Tm1 || !w) { set_error("mom5adv_vert_dev: null argument"); return MOM5ADV_EINVAL; }
        return vert_dev(h->g, h->qw, h->tmask, h->dat, Tm1, Tm1, nullptr, w, th, wrk1, fz, 0, st, &h->launches);
    case MOM5ADV_ADVECT_QUICKER:
        if (!Tm1 || !Tt || !tlimit || !w) { set_error("mom5adv_vert_dev: quicker needs T_taum1, T_tau, tmask_limit, wrho_bt"); return MOM5ADV_EINVAL; }
        return vert_dev(h->g, h->qw, h->tmask, h->dat, Tm1, Tt, tlimit, w, th, wrk1, fz, 1, st, &h->launches);
    default:
        set_error("mom5adv_vert_dev: invalid advection scheme chosen %d", scheme);   // OTA:2157-2159
        return MOM5ADV_EINVAL;
    }
}

extern "C" int mom5adv_horz(mom5adv_handle h, int scheme, double dtime, const double *Tm1, const double *Tt, const double *tlimit,
                            int limit_with_upwind, const double *u, const double *v, const double *w, const double *rho,
                            double *th, double *wrk1, double *fx, double *fy, double *fz)
{
    if (!h || !Tm1 || !u || !v || !th || !wrk1) { set_error("mom5adv_horz: null argument"); return MOM5ADV_EINVAL; }
    cudaStream_t st = h->stream;
    const size_t N = n3(h);
    int rc;
    // slots 0,1 are reserved as flux scratch by the quicker path
    size_t slot = 2;
    double *dTm1, *dTt = 0, *dtl = 0, *du, *dv, *dw = 0, *dr = 0, *dth, *dwrk, *dfx = 0, *dfy = 0, *dfz = 0;
    if ((rc = mirror(h, slot++, &dTm1)) || (rc = mirror(h, slot++, &du)) || (rc = mirror(h, slot++, &dv)) ||
        (rc = mirror(h, slot++, &dth)) || (rc = mirror(h, slot++, &dwrk))) return rc;
    H2D(dTm1, Tm1, N); H2D(du, u, N); H2D(dv, v, N); H2D(dth, th, N);
    if (Tt) { if ((rc = mirror(h, slot++, &dTt))) return rc; H2D(dTt, Tt, N); }
    if (tlimit) { if ((rc = mirror(h, slot++, &dtl))) return rc; H2D(dtl, tlimit, N); }
    if (rho) { if ((rc = mirror(h, slot++, &dr))) return rc; H2D(dr, rho, N); }
    if (w) { if ((rc = mirror_w(h, &dw))) return rc; H2D(dw, w, (size_t)h->g.slab * (h->g.nk + 1)); }
    if (fx) { if ((rc = mirror(h, slot++, &dfx))) return rc; H2D(dfx, fx, N); }
    if (fy) { if ((rc = mirror(h, slot++, &dfy))) return rc; H2D(dfy, fy, N); }
    if (fz) { if ((rc = mirror(h, slot++, &dfz))) return rc; H2D(dfz, fz, N); }
    if ((rc = mom5adv_horz_dev(h, scheme, dtime, dTm1, dTt, dtl, limit_with_upwind, du, dv, dw, dr, dth, dwrk, dfx, dfy, dfz, st))) return rc;
    D2H(th, dth, N); D2H(wrk1, dwrk, N);
    if (fx) D2H(fx, dfx, N);
    if (fy) D2H(fy, dfy, N);
    if (fz) D2H(fz, dfz, N);
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int mom5adv_vert(mom5adv_handle h, int scheme, const double *Tm1, const double *Tt, const double *tlimit, const double *w,
                            double *th, double *wrk1, double *fz)
{
    if (!h || !th || !wrk1) { set_error("mom5adv_vert: null argument"); return MOM5ADV_EINVAL; }
    cudaStream_t st = h->stream;
    const size_t N = n3(h);
    int rc;
    size_t slot = 2;
    double *dTm1 = 0, *dTt = 0, *dtl = 0, *dw = 0, *dth, *dwrk, *dfz = 0;
    if ((rc = mirror(h, slot++, &dth)) || (rc = mirror(h, slot++, &dwrk))) return rc;
    H2D(dth, th, N);
    if (Tm1) { if ((rc = mirror(h, slot++, &dTm1))) return rc; H2D(dTm1, Tm1, N); }
    if (Tt) { if ((rc = mirror(h, slot++, &dTt))) return rc; H2D(dTt, Tt, N); }
    if (tlimit) { if ((rc = mirror(h, slot++, &dtl))) return rc; H2D(dtl, tlimit, N); }
    if (w) { if ((rc = mirror_w(h, &dw))) return rc; H2D(dw, w, (size_t)h->g.slab * (h->g.nk + 1)); }
    if (fz) { if ((rc = mirror(h, slot++, &dfz))) return rc; H2D(dfz, fz, N); }
    if ((rc = mom5adv_vert_dev(h, scheme, dTm1, dTt, dtl, dw, dth, dwrk, dfz, st))) return rc;
    D2H(th, dth, N); D2H(wrk1, dwrk, N);
    if (fz) D2H(fz, dfz, N);
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// diagnostics producers
// ------------------------------------------------------------------------------------------------
extern "C" int mom5adv_adv_diss_dev(mom5adv_handle h, int horz_scheme, int vert_scheme, double dtime, double conversion,
                                    const double *T_tau, const double *tlimit, int limit_with_upwind, const double *u,
                                    const double *v, const double *w, const double *rho_tau, const double *rho_taup1,
                                    const double *advect_tendency, double *adv_diss, double *t2_tendency, void *stream)
{
    if (!h || !T_tau || !u || !v || !w || !rho_tau || !rho_taup1 || !advect_tendency || !adv_diss) {
        set_error("mom5adv_adv_diss_dev: null argument");
        return MOM5ADV_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const Geom &g = h->g;
    const size_t N = n3(h);
    int rc;
    double *sq, *w2, *w3, *thd;   // wrk1 (squared tracer), wrk2, wrk3 of the reference; thd absorbs the dispatchers' th += wrk1
    if ((rc = mirror(h, 2, &sq)) || (rc = mirror(h, 3, &w2)) || (rc = mirror(h, 4, &w3)) || (rc = mirror(h, 5, &thd))) return rc;
    CUDA_TRY(cudaMemsetAsync(thd, 0, N * sizeof(double), st));
    LAUNCH(h, k_square, 148 * 8, 256, 0, st, N, T_tau, sq);
    switch (horz_scheme) {   // OTA:7583-7626: the operator acts on the squared tracer at BOTH time levels
    case MOM5ADV_ADVECT_UPWIND: case MOM5ADV_ADVECT_QUICKER: case MOM5ADV_ADVECT_MDFL_SWEBY: case MOM5ADV_ADVECT_DST_LINEAR:
    case MOM5ADV_ADVECT_DST_LINEAR_TEST: case MOM5ADV_ADVECT_MDPPM:
        if ((rc = mom5adv_horz_dev(h, horz_scheme, dtime, sq, sq, tlimit, limit_with_upwind, u, v, w, rho_tau, thd, w2, nullptr, nullptr, nullptr, st))) return rc;
        break;
    case MOM5ADV_ADVECT_MDFL_SWEBY_TEST:   // has no arm in compute_adv_diss's select: wrk2 stays 0
        CUDA_TRY(cudaMemsetAsync(w2, 0, N * sizeof(double), st));
        break;
    default:
        set_error("mom5adv_adv_diss_dev: horz advection scheme %d is not covered by the GPU path", horz_scheme);
        return MOM5ADV_EINVAL;
    }
    switch (vert_scheme) {   // OTA:7628-7662
    case MOM5ADV_ADVECT_UPWIND: case MOM5ADV_ADVECT_QUICKER: case MOM5ADV_ADVECT_MDFL_SWEBY: case MOM5ADV_ADVECT_DST_LINEAR:
    case MOM5ADV_ADVECT_MDFL_SWEBY_TEST: case MOM5ADV_ADVECT_DST_LINEAR_TEST: case MOM5ADV_ADVECT_MDPPM:
        if ((rc = mom5adv_vert_dev(h, vert_scheme, sq, sq, tlimit, w, thd, w3, nullptr, st))) return rc;
        break;
    default:
        set_error("mom5adv_adv_diss_dev: vert advection scheme %d is not covered by the GPU path", vert_scheme);
        return MOM5ADV_EINVAL;
    }
    DissArgs a{rho_tau, rho_taup1, T_tau, advect_tendency, w2, w3, t2_tendency, adv_diss, dtime, 1.0 / dtime, conversion};
    LAUNCH(h, k_adv_diss, dim3((g.ni + 2 + 127) / 128, g.nj + 2, g.nk), 128, 0, st, g, a);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// host-pointer twin (what the Fortran shim binds at OTA:2221-2223): H2D of the operands, the device path, D2H of the results
extern "C" int mom5adv_adv_diss(mom5adv_handle h, int horz_scheme, int vert_scheme, double dtime, double conversion,
                                const double *T_tau, const double *tlimit, int limit_with_upwind, const double *u, const double *v,
                                const double *w, const double *rho_tau, const double *rho_taup1, const double *advect_tendency,
                                double *adv_diss, double *t2_tendency)
{
    if (!h || !T_tau || !u || !v || !w || !rho_tau || !rho_taup1 || !advect_tendency || !adv_diss) {
        set_error("mom5adv_adv_diss: null argument");
        return MOM5ADV_EINVAL;
    }
    cudaStream_t st = h->stream;
    const size_t N = n3(h);
    int rc;
    size_t slot = 6;   // slots 0,1 are the flux work arrays of the dispatchers, 2..5 the work arrays of mom5adv_adv_diss_dev
    double *dT, *dtl = 0, *du, *dv, *dw, *dr0, *dr1, *dadv, *ddiss, *dt2 = 0;
    if ((rc = mirror(h, slot++, &dT)) || (rc = mirror(h, slot++, &du)) || (rc = mirror(h, slot++, &dv)) || (rc = mirror(h, slot++, &dr0)) ||
        (rc = mirror(h, slot++, &dr1)) || (rc = mirror(h, slot++, &dadv)) || (rc = mirror(h, slot++, &ddiss)) || (rc = mirror_w(h, &dw))) return rc;
    H2D(dT, T_tau, N); H2D(du, u, N); H2D(dv, v, N); H2D(dr0, rho_tau, N); H2D(dr1, rho_taup1, N); H2D(dadv, advect_tendency, N);
    H2D(dw, w, (size_t)h->g.slab * (h->g.nk + 1));
    if (tlimit) { if ((rc = mirror(h, slot++, &dtl))) return rc; H2D(dtl, tlimit, N); }
    if (t2_tendency && (rc = mirror(h, slot++, &dt2))) return rc;
    if ((rc = mom5adv_adv_diss_dev(h, horz_scheme, vert_scheme, dtime, conversion, dT, dtl, limit_with_upwind, du, dv, dw, dr0, dr1, dadv,
                                   ddiss, dt2, st))) return rc;
    D2H(adv_diss, ddiss, N);
    if (t2_tendency) D2H(t2_tendency, dt2, N);
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int mom5adv_set_ppm_limiters(mom5adv_handle h, int ppm_hlimiter, int ppm_vlimiter)
{
    if (!h) { set_error("mom5adv_set_ppm_limiters: null handle"); return MOM5ADV_EINVAL; }
    h->ppm_hlimiter = ppm_hlimiter;   // validated where the reference validates it: inside the scheme (OTA:6137-6148)
    h->ppm_vlimiter = ppm_vlimiter;
    return 0;
}

extern "C" int mom5adv_flux_int_z_dev(mom5adv_handle h, const double *flux3d, double *out2d, void *stream)
{
    if (!h || !flux3d || !out2d) { set_error("mom5adv_flux_int_z_dev: null argument"); return MOM5ADV_EINVAL; }
    const Geom &g = h->g;
    LAUNCH(h, k_flux_int_z, dim3((g.ni + 2 + 127) / 128, g.nj + 2), 128, 0, (cudaStream_t)stream, g, flux3d, out2d);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// consumer of the tendencies: tracer time update + halo-1 update (ocean_tracer.F90:2341-2350, ocean_model.F90:1903-1911)
// ------------------------------------------------------------------------------------------------
extern "C" int mom5adv_tracer_update_dev(mom5adv_handle h, int ntr, double dtime, const double *rho_dzt_taum1,
                                         const double *rho_dztr_taup1, const double *const *T_taum1,
                                         const double *const *th_tendency, double *const *T_taup1, void *stream)
{
    if (!h || !rho_dzt_taum1 || !rho_dztr_taup1 || !T_taum1 || !th_tendency || !T_taup1 || ntr < 1) {
        set_error("mom5adv_tracer_update_dev: bad argument");
        return MOM5ADV_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const Geom &g = h->g;
    for (int n0 = 0; n0 < ntr; n0 += MAXNT) {
        UpdArgs<MAXNT> a{};
        for (int q = 0; q < MAXNT && n0 + q < ntr; q++) { a.T[q] = T_taum1[n0 + q]; a.th[q] = th_tendency[n0 + q]; a.Tnew[q] = T_taup1[n0 + q]; }
        a.rho_m1 = rho_dzt_taum1; a.rho_r = rho_dztr_taup1; a.dtime = dtime;
        LAUNCH(h, k_tracer_update<MAXNT>, dim3((g.ni + 127) / 128, g.nj, g.nk), 128, 0, st, g, a);
    }
    std::vector<double *> f(T_taup1, T_taup1 + ntr);
    return halo_update_l<1>(h, f.data(), ntr, 3, st);
}

// update_advection_only (ocean_tracer.F90:2618-2649) with advect_tracer_sweby_all as the advection operator, in two
// kernels per tracer group: z sweep, then the fused x/y pass whose epilogue forms th_tendency = 0 + wrk1 and
// field(taup1) = (rho_dzt(taum1)*field(taum1) + dtime*th_tendency)*rho_dztr; then the halo-1 update of field(taup1)
// (ocean_model.F90:1903-1911).  Saves, per cell and tracer, the read of th_tendency, the writes of th_tendency and wrk1 (when
// the caller passes NULL for them) and the whole stand-alone update pass.
extern "C" int mom5adv_sweby_all_step_dev(mom5adv_handle h, int ntr, double dtime, const double *const *T_taum1,
                                          const double *rho_dzt_taum1, const double *rho_dztr_taup1, const double *u, const double *v,
                                          const double *w, const double *rho_tau, double *const *T_taup1, double *const *th_out,
                                          double *const *adv_out, void *stream)
{
    if (!h || !T_taum1 || !rho_dzt_taum1 || !rho_dztr_taup1 || !u || !v || !w || !rho_tau || !T_taup1) {
        set_error("mom5adv_sweby_all_step_dev: null argument");
        return MOM5ADV_EINVAL;
    }
    if (!h->fuse) { set_error("mom5adv_sweby_all_step_dev needs the fused driver (MOM5ADV_FUSE=0 is set)"); return MOM5ADV_EUNSUP; }
    cudaStream_t st = (cudaStream_t)stream;
    SwebyCall c{ntr, VAR_ALL, dtime, 1.0, T_taum1, th_out, adv_out, u, v, w, rho_tau, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    c.Tnew = T_taup1; c.rho_m1 = rho_dzt_taum1; c.rho_r = rho_dztr_taup1;
    int rc = sweby_dev(h, c, st);
    if (rc) return rc;
    std::vector<double *> f(T_taup1, T_taup1 + ntr);
    return halo_update_l<1>(h, f.data(), ntr, 3, st);
}

// ------------------------------------------------------------------------------------------------
// producer of wrho_bt: continuity (ocean_advection_velocity.F90:660-669)
// ------------------------------------------------------------------------------------------------
extern "C" int mom5adv_continuity_dev(mom5adv_handle h, const double *uhrho_et, const double *vhrho_nt,
                                      const double *rho_dzt_tendency, const double *mass_source, double *wrho_bt,
                                      double *diverge_t, void *stream)
{
    if (!h || !uhrho_et || !vhrho_nt || !wrho_bt) { set_error("mom5adv_continuity_dev: null argument"); return MOM5ADV_EINVAL; }
    const Geom &g = h->g;
    LAUNCH(h, k_continuity, dim3((g.ni + 2 + 127) / 128, g.nj + 2), 128, 0, (cudaStream_t)stream, g, h->tmask, h->dyte, h->dxtn,
           h->datr, uhrho_et, vhrho_nt, rho_dzt_tendency, mass_source, wrho_bt, diverge_t);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// metrics
// ------------------------------------------------------------------------------------------------
__global__ void k_chksum(const Geom g, const double *__restrict__ f, const double *__restrict__ tmask, unsigned long long *out)
{
    unsigned long long s = 0;
    const long long tot = (long long)g.ni * g.nj * g.nk;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % g.ni) + 1;
        const int j = (int)((e / g.ni) % g.nj) + 1;
        const int k = (int)(e / ((long long)g.ni * g.nj)) + 1;
        double v = f[d3(g, i, j, k)];
        if (tmask) v = v * tmask[d3(g, i, j, k)];
        s += (unsigned long long)__double_as_longlong(v);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// per-column partial sums in the reference's order, then an (order-free up to round-off) column reduction
__global__ void k_total_tracer(const Geom g, const double *__restrict__ tmask, const double *__restrict__ dat,
                               const double *__restrict__ rho, const double *__restrict__ T, double *out)
{
    double s = 0.0;
    const long long tot = (long long)g.ni * g.nj;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % g.ni) + 1, j = (int)(e / g.ni) + 1;
        double tk = 0.0;
        for (int k = 1; k <= g.nk; k++)
            tk = tk + (((tmask[d3(g, i, j, k)] * dat[d2(g, i, j)]) * rho[d3(g, i, j, k)]) * T[d3(g, i, j, k)]);
        s += tk;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

extern "C" int mom5adv_chksum_dev(mom5adv_handle h, const double *f, int masked, int64_t *out, void *stream)
{
    if (!h || !f || !out) { set_error("mom5adv_chksum_dev: null argument"); return MOM5ADV_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *d;
    CUDA_TRY(cudaMalloc(&d, 8));
    CUDA_TRY(cudaMemsetAsync(d, 0, 8, st));
    LAUNCH(h, k_chksum, 148 * 8, 256, 0, st, h->g, f, masked ? h->tmask : nullptr, d);
    CUDA_TRY(cudaMemcpyAsync(out, d, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaFree(d);
    return 0;
}

extern "C" int mom5adv_total_tracer_dev(mom5adv_handle h, const double *rho, const double *T, double *out, void *stream)
{
    if (!h || !rho || !T || !out) { set_error("mom5adv_total_tracer_dev: null argument"); return MOM5ADV_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    double *d;
    CUDA_TRY(cudaMalloc(&d, 8));
    CUDA_TRY(cudaMemsetAsync(d, 0, 8, st));
    LAUNCH(h, k_total_tracer, 148 * 4, 256, 0, st, h->g, h->tmask, h->dat, rho, T, d);
    CUDA_TRY(cudaMemcpyAsync(out, d, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaFree(d);
    return 0;
}

extern "C" int mom5adv_last_timing_ms(mom5adv_handle h, float ms[5])
{
    if (!h || !ms) { set_error("mom5adv_last_timing_ms: null argument"); return MOM5ADV_EINVAL; }
    for (int q = 0; q < 5; q++) ms[q] = 0.f;
    if (!h->ev_valid) return 0;
    CUDA_TRY(cudaEventSynchronize(h->ev[5]));
    float hx = 0, hy = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms[0], h->ev[0], h->ev[1]));
    CUDA_TRY(cudaEventElapsedTime(&hx, h->ev[1], h->ev[2]));
    CUDA_TRY(cudaEventElapsedTime(&ms[1], h->ev[2], h->ev[3]));
    CUDA_TRY(cudaEventElapsedTime(&hy, h->ev[3], h->ev[4]));
    CUDA_TRY(cudaEventElapsedTime(&ms[2], h->ev[4], h->ev[5]));
    CUDA_TRY(cudaEventElapsedTime(&ms[4], h->ev[0], h->ev[5]));
    ms[3] = hx + hy;
    return 0;
}

extern "C" int64_t mom5adv_kernel_launches(mom5adv_handle h) { return h ? h->launches : 0; }

extern "C" int mom5adv_last_transfer_bytes(mom5adv_handle h, int64_t bytes[2])
{
    if (!h || !bytes) { set_error("mom5adv_last_transfer_bytes: null argument"); return MOM5ADV_EINVAL; }
    bytes[0] = h->h2d_bytes; bytes[1] = h->d2h_bytes;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// quicker_init on device (OTA:1442-1586)
// ------------------------------------------------------------------------------------------------
static int quicker_setup(mom5adv_ctx *h, const mom5adv_grid *G, cudaStream_t st)
{
    // dxt_quick / dyt_quick: compute domain := dxt, dyt; edge replication (OTA:1490-1509); full halo-2 update where an
    // image exists (OTA:1510-1511).  Done on the host for the replication part (2-D, once), on the device for the update.
    const Geom &g = h->g;
    const int nx2 = g.ni + 4, ny2 = g.nj + 4;
    if (!G->dxt || !G->dyt || !G->dzt) { set_error("mom5adv_init: dxt, dyt, dzt are required"); return MOM5ADV_EINVAL; }
    std::vector<double> dx((size_t)nx2 * ny2, 0.0), dy((size_t)nx2 * ny2, 0.0);
    auto H = [&](int i, int j) { return (size_t)(i + 1) + (size_t)nx2 * (j + 1); };
    auto D = [&](int i, int j) { return (size_t)i + (size_t)g.nxd * j; };
    for (int j = 1; j <= g.nj; j++)
        for (int i = 1; i <= g.ni; i++) { dx[H(i, j)] = G->dxt[D(i, j)]; dy[H(i, j)] = G->dyt[D(i, j)]; }
    for (int i = -1; i <= 0; i++) {
        for (int j = 1; j <= g.nj; j++) dx[H(i, j)] = dx[H(1, j)];
        for (int j = -1; j <= g.nj + 2; j++) dy[H(i, j)] = dy[H(1, j)];
    }
    for (int i = g.ni + 1; i <= g.ni + 2; i++) {
        for (int j = 1; j <= g.nj; j++) dx[H(i, j)] = dx[H(g.ni, j)];
        for (int j = -1; j <= g.nj + 2; j++) dy[H(i, j)] = dy[H(g.ni, j)];
    }
    for (int j = -1; j <= 0; j++)
        for (int i = -1; i <= g.ni + 2; i++) { dx[H(i, j)] = dx[H(i, 1)]; dy[H(i, j)] = dy[H(i, 1)]; }
    for (int j = g.nj + 1; j <= g.nj + 2; j++)
        for (int i = -1; i <= g.ni + 2; i++) { dx[H(i, j)] = dx[H(i, g.nj)]; dy[H(i, j)] = dy[H(i, g.nj)]; }
    // stage into one-level h2 fields on the device, update, read back
    Geom g1 = g;
    g1.nk = 1;
    std::vector<double> sx((size_t)g.tslab, 0.0), sy((size_t)g.tslab, 0.0);
    for (int j = -1; j <= g.nj + 2; j++)
        for (int i = -1; i <= g.ni + 2; i++) { sx[t3(g1, i, j, 1)] = dx[H(i, j)]; sy[t3(g1, i, j, 1)] = dy[H(i, j)]; }
    double *dsx = h->tmA[0], *dsy = h->tmB[0];
    CUDA_TRY(cudaMemcpyAsync(dsx, sx.data(), sx.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dsy, sy.data(), sy.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    const int nk_save = h->g.nk;
    h->g.nk = 1;
    double *f2[2] = {dsx, dsy};
    int rc = halo_update(h, f2, 2, 3, st);
    h->g.nk = nk_save;
    if (rc) return rc;
    if ((rc = alloc_quickw(h->qw, g))) { set_error("quicker weights: allocation failed"); return MOM5ADV_ECUDA; }
    LAUNCH(h, k_quicker_weights, dim3((g.ni + 1 + 127) / 128, g.nj + 1), 128, 0, st, g, dsx, dsy, h->qw);
    // vertical weights on the host (nk values), OTA:1566-1578
    std::vector<double> qz(2 * g.nk), zp(3 * g.nk), zn(3 * g.nk);
    for (int k = 1; k <= g.nk; k++) {
        const int kp2 = std::min(k + 2, g.nk), kp1 = std::min(k + 1, g.nk), km1 = std::max(k - 1, 1);
        const double zm = G->dzt[km1 - 1], z0 = G->dzt[k - 1], z1 = G->dzt[kp1 - 1], z2 = G->dzt[kp2 - 1];
        qz[(k - 1)] = z1 / (z1 + z0);
        qz[(k - 1) + g.nk] = z0 / (z1 + z0);
        zp[(k - 1)] = (z0 * z1) / (((zm + (2.0 * z0)) + z1) * (z0 + z1));
        zp[(k - 1) + g.nk] = -((z0 * z1) / ((z0 + z1) * (zm + z0)));
        zp[(k - 1) + 2 * g.nk] = (z0 * z1) / (((zm + (2.0 * z0)) + z1) * (zm + z0));
        zn[(k - 1)] = (z0 * z1) / (((z0 + (2.0 * z1)) + z2) * (z1 + z2));
        zn[(k - 1) + g.nk] = -((z0 * z1) / ((z1 + z2) * (z0 + z1)));
        zn[(k - 1) + 2 * g.nk] = (z0 * z1) / (((z0 + (2.0 * z1)) + z2) * (z0 + z1));
    }
    CUDA_TRY(cudaMemcpyAsync(h->qw.quick_z, qz.data(), qz.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->qw.curv_zp, zp.data(), zp.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->qw.curv_zn, zn.data(), zn.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaMemsetAsync(dsx, 0, (size_t)g.tslab * sizeof(double), st));
    CUDA_TRY(cudaMemsetAsync(dsy, 0, (size_t)g.tslab * sizeof(double), st));
    return 0;
}
