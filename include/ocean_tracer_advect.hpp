// ocean_tracer_advect.hpp -- C++ host-side mirror of the reference's operator interface for the tracer
// advection path, on top of the C ABI of mom5adv.h.
//
// The reference is compiled Fortran; its module ocean_tracer_advect_mod exposes
//     ocean_tracer_advect_init   (OTA:507)        horz_advect_tracer (OTA:1898)
//     vert_advect_tracer         (OTA:2095)       ocean_tracer_advect_end
// (OTA = src/mom5/ocean_tracers/ocean_tracer_advect.F90) working on the derived types of
// src/mom5/ocean_core/ocean_types.F90.  This header restates that interface -- same names, same argument order and
// meaning, same error behaviour (an invalid scheme is fatal, OTA:1983-1985 / 2157-2159 -> std::runtime_error in place of
// mpp_error(FATAL)) -- for C++ callers; the Fortran model itself binds the same C ABI through
// mom5_b200/fortran/ocean_tracer_advect_gpu.F90.  Header-only, no dependency besides mom5adv.h / libmom5adv.so.
//
// Arrays are the caller's, Fortran column-major, dimensioned exactly as in the reference:
//   3-D (isd:ied, jsd:jed, nk), field (isd:ied, jsd:jed, nk, 3), wrho_bt (isd:ied, jsd:jed, 0:nk); HOST pointers.
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "mom5adv.h"

namespace mom5 {

// scheme ids, ocean_parameters.F90:149-163
enum : int { ADVECT_UPWIND = 1, ADVECT_QUICKER = 5, ADVECT_MDPPM = 8, ADVECT_MDFL_SWEBY = 9, ADVECT_DST_LINEAR = 10,
             ADVECT_MDFL_SWEBY_TEST = 12, ADVECT_DST_LINEAR_TEST = 14 };

// ocean_time_type (ocean_types.F90:937-947): 1-based time-level indices into field(:,:,:,1:3)
struct ocean_time_type {
    int taum1 = 1, tau = 2, taup1 = 3;
};

// ocean_domain_type (ocean_types.F90:922-935)
struct ocean_domain_type {
    int isc, iec, jsc, jec;   // compute domain, global indices
    int isd, ied, jsd, jed;   // data domain (halo 1)
    int layout[2] = {1, 1};
};

// ocean_grid_type (ocean_types.F90:747-920), the members this path reads
struct ocean_grid_type {
    int ni, nj, nk;
    bool cyclic_x = false, cyclic_y = false, tripolar = false;
    const double *dat, *datr, *dxt, *dyt, *dxte, *dyte, *dxtn, *dytn;   // (isd:ied, jsd:jed)
    const double *dzt;                                                  // (nk)
    const double *tmask;                                                // (isd:ied, jsd:jed, nk)
};

// ocean_adv_vel_type (ocean_types.F90:949-958)
struct ocean_adv_vel_type {
    const double *uhrho_et, *vhrho_nt;   // (isd:ied, jsd:jed, nk)
    const double *wrho_bt;               // (isd:ied, jsd:jed, 0:nk)
};

// ocean_thickness_type (ocean_types.F90:674-744)
struct ocean_thickness_type {
    const double *rho_dzt;               // (isd:ied, jsd:jed, nk, 3)
};

// ocean_prog_tracer_type (ocean_types.F90:1004-1077), the members this path touches
struct ocean_prog_tracer_type {
    std::string name;
    double *field;                        // (isd:ied, jsd:jed, nk, 3)
    double *th_tendency;                  // (isd:ied, jsd:jed, nk)
    double *wrk1;                         // (isd:ied, jsd:jed, nk)
    const double *tmask_limit = nullptr;  // (isd:ied, jsd:jed, nk)
    int horz_advect_scheme = ADVECT_MDFL_SWEBY, vert_advect_scheme = ADVECT_MDFL_SWEBY;
    int ppm_hlimiter = 1, ppm_vlimiter = 1;   // ocean_types.F90:1018-1019 (ADVECT_MDPPM)
};

// ocean_tracer_advect_nml (OTA:483-495), the switches that change this path
struct ocean_tracer_advect_nml {
    bool advect_sweby_all = false;
    bool limit_with_upwind = false;
    bool zero_tracer_advect_horz = false, zero_tracer_advect_vert = false;
    bool have_obc = false;
};

class ocean_tracer_advect {
public:
    // ocean_tracer_advect_init (OTA:507-756) + mdfl_init (OTA:1644-1691) + quicker_init (OTA:1442-1586)
    ocean_tracer_advect(const ocean_grid_type &Grid, const ocean_domain_type &Domain, int num_prog_tracers,
                        const ocean_tracer_advect_nml &nml = {}, mom5adv_comm comm = nullptr)
        : nml_(nml), num_prog_tracers_(num_prog_tracers)
    {
        slab_ = (size_t)(Domain.ied - Domain.isd + 1) * (size_t)(Domain.jed - Domain.jsd + 1);
        n3_ = slab_ * (size_t)Grid.nk;
        mom5adv_grid g{};
        g.isc = Domain.isc; g.iec = Domain.iec; g.jsc = Domain.jsc; g.jec = Domain.jec; g.nk = Grid.nk;
        g.ni_global = Grid.ni; g.nj_global = Grid.nj; g.layout_x = Domain.layout[0]; g.layout_y = Domain.layout[1];
        g.cyclic_x = Grid.cyclic_x; g.cyclic_y = Grid.cyclic_y; g.tripolar = Grid.tripolar; g.have_obc = nml.have_obc;
        g.dat = Grid.dat; g.datr = Grid.datr; g.dxt = Grid.dxt; g.dyt = Grid.dyt; g.dxte = Grid.dxte; g.dyte = Grid.dyte;
        g.dxtn = Grid.dxtn; g.dytn = Grid.dytn; g.dzt = Grid.dzt; g.tmask = Grid.tmask;
        check(mom5adv_init(&g, num_prog_tracers, comm, &h_), "ocean_tracer_advect_init");
    }
    ~ocean_tracer_advect() { mom5adv_finalize(h_); }   // ocean_tracer_advect_end
    ocean_tracer_advect(const ocean_tracer_advect &) = delete;
    ocean_tracer_advect &operator=(const ocean_tracer_advect &) = delete;

    // subroutine horz_advect_tracer(Time, Adv_vel, Thickness, Dens, T_prog, Tracer, ntracer, dtime)   OTA:1898-2083
    // (Dens is only used by the water-mass diagnostics, out of scope; ntracer is 1-based as in the reference)
    void horz_advect_tracer(const ocean_time_type &Time, const ocean_adv_vel_type &Adv_vel,
                            const ocean_thickness_type &Thickness, std::vector<ocean_prog_tracer_type> &T_prog,
                            ocean_prog_tracer_type &Tracer, int ntracer, double dtime)
    {
        if (nml_.zero_tracer_advect_horz) return;                                  // OTA:1915
        const double *rho_tau = Thickness.rho_dzt + n3_ * (size_t)(Time.tau - 1);
        if (!nml_.advect_sweby_all) {                                              // OTA:1923-2075
            switch (Tracer.horz_advect_scheme) {
            case ADVECT_UPWIND: case ADVECT_QUICKER: case ADVECT_MDFL_SWEBY: case ADVECT_DST_LINEAR:
            case ADVECT_MDFL_SWEBY_TEST: case ADVECT_DST_LINEAR_TEST: case ADVECT_MDPPM: break;
            default: throw std::runtime_error("==>Error from ocean_tracer_advect_mod (horz_advect_tracer): chose invalid horz advection scheme");
            }
            if (Tracer.horz_advect_scheme == ADVECT_MDPPM)
                check(mom5adv_set_ppm_limiters(h_, Tracer.ppm_hlimiter, Tracer.ppm_vlimiter), "ppm limiters");
            check(mom5adv_horz(h_, Tracer.horz_advect_scheme, dtime, level(Tracer.field, Time.taum1), level(Tracer.field, Time.tau),
                               Tracer.tmask_limit, nml_.limit_with_upwind, Adv_vel.uhrho_et, Adv_vel.vhrho_nt, Adv_vel.wrho_bt,
                               rho_tau, Tracer.th_tendency, Tracer.wrk1, nullptr, nullptr, nullptr),
                  "horz_advect_tracer");
        }
        if (nml_.advect_sweby_all && ntracer == 1) {                               // OTA:2078-2080
            std::vector<const double *> T(T_prog.size());
            std::vector<double *> th(T_prog.size()), adv(T_prog.size());
            for (size_t n = 0; n < T_prog.size(); n++) {
                T[n] = level(T_prog[n].field, Time.taum1);
                th[n] = T_prog[n].th_tendency;
                adv[n] = T_prog[n].wrk1;
            }
            check(mom5adv_sweby_all(h_, (int)T_prog.size(), dtime, T.data(), th.data(), adv.data(), Adv_vel.uhrho_et,
                                    Adv_vel.vhrho_nt, Adv_vel.wrho_bt, rho_tau, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr),
                  "advect_tracer_sweby_all");
        }
    }

    // subroutine vert_advect_tracer(Time, Adv_vel, Dens, Thickness, T_prog, Tracer, ntracer, dtime)   OTA:2095-2226
    void vert_advect_tracer(const ocean_time_type &Time, const ocean_adv_vel_type &Adv_vel, ocean_prog_tracer_type &Tracer)
    {
        if (nml_.zero_tracer_advect_vert) return;                                  // OTA:2109
        if (nml_.advect_sweby_all) return;                                         // OTA:2114
        switch (Tracer.vert_advect_scheme) {
        case ADVECT_UPWIND: case ADVECT_QUICKER: case ADVECT_MDFL_SWEBY: case ADVECT_DST_LINEAR:
            case ADVECT_MDFL_SWEBY_TEST: case ADVECT_DST_LINEAR_TEST: case ADVECT_MDPPM: break;
        default: throw std::runtime_error("==>Error from ocean_tracer_advect_mod (vert_advect_tracer): invalid advection scheme chosen");
        }
        check(mom5adv_vert(h_, Tracer.vert_advect_scheme, level(Tracer.field, Time.taum1), level(Tracer.field, Time.tau),
                           Tracer.tmask_limit, Adv_vel.wrho_bt, Tracer.th_tendency, Tracer.wrk1, nullptr),
              "vert_advect_tracer");
    }

    mom5adv_handle handle() const { return h_; }

private:
    static void check(int rc, const char *where)
    {
        if (rc != 0) throw std::runtime_error(std::string("==>Error from ocean_tracer_advect_mod (") + where + "): " + mom5adv_last_error());
    }
    const double *level(const double *field, int t) const { return field + n3_ * (size_t)(t - 1); }
    ocean_tracer_advect_nml nml_;
    int num_prog_tracers_;
    size_t slab_ = 0, n3_ = 0;
    mom5adv_handle h_ = nullptr;
};

}  // namespace mom5
