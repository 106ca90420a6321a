// mirror_driver.cpp -- exercises include/ocean_tracer_advect.hpp (the C++ mirror of the reference's
// horz_advect_tracer / vert_advect_tracer interface) from plain C++, the way a compiled caller would.
//   mirror_driver <dir>      reads <dir>/meta.txt and raw float64 arrays, runs the dispatchers, writes results.
// Built and run by tests/test_cpp_mirror.py; results are compared with the reference-text golden vectors.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "ocean_tracer_advect.hpp"

static std::vector<double> rd(const std::string &p, size_t n)
{
    std::vector<double> v(n);
    std::ifstream f(p, std::ios::binary);
    if (!f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(n * 8))) { std::cerr << "cannot read " << p << "\n"; std::exit(2); }
    return v;
}
static void wr(const std::string &p, const double *v, size_t n)
{
    std::ofstream f(p, std::ios::binary);
    f.write(reinterpret_cast<const char *>(v), (std::streamsize)(n * 8));
}

int main(int argc, char **argv)
{
    if (argc < 2) { std::cerr << "usage: mirror_driver <dir> [--link-check]\n"; return 2; }
    if (std::string(argv[1]) == "--link-check") { std::printf("mom5adv %d\n", mom5adv_version()); return 0; }
    const std::string d = std::string(argv[1]) + "/";
    int ni, nj, nk, ntr, cx, cy, tri, mode, scheme, limit;
    double dtime;
    { std::ifstream m(d + "meta.txt"); m >> ni >> nj >> nk >> ntr >> cx >> cy >> tri >> dtime >> mode >> scheme >> limit; }
    const size_t n2 = (size_t)(ni + 2) * (nj + 2), n3 = n2 * nk;
    auto dat = rd(d + "dat.bin", n2), datr = rd(d + "datr.bin", n2), dxt = rd(d + "dxt.bin", n2), dyt = rd(d + "dyt.bin", n2),
         dxte = rd(d + "dxte.bin", n2), dyte = rd(d + "dyte.bin", n2), dxtn = rd(d + "dxtn.bin", n2), dytn = rd(d + "dytn.bin", n2);
    auto dzt = rd(d + "dzt.bin", nk), tmask = rd(d + "tmask.bin", n3);
    auto u = rd(d + "u.bin", n3), v = rd(d + "v.bin", n3), w = rd(d + "w.bin", n2 * (nk + 1)), rho = rd(d + "rho.bin", 3 * n3);
    mom5::ocean_grid_type G{};
    G.ni = ni; G.nj = nj; G.nk = nk; G.cyclic_x = cx; G.cyclic_y = cy; G.tripolar = tri;
    G.dat = dat.data(); G.datr = datr.data(); G.dxt = dxt.data(); G.dyt = dyt.data(); G.dxte = dxte.data(); G.dyte = dyte.data();
    G.dxtn = dxtn.data(); G.dytn = dytn.data(); G.dzt = dzt.data(); G.tmask = tmask.data();
    mom5::ocean_domain_type D{1, ni, 1, nj, 0, ni + 1, 0, nj + 1, {1, 1}};
    mom5::ocean_adv_vel_type A{u.data(), v.data(), w.data()};
    mom5::ocean_thickness_type Th{rho.data()};
    mom5::ocean_time_type Time;   // taum1 = 1, tau = 2, taup1 = 3
    std::vector<std::vector<double>> field(ntr), th(ntr), wrk1(ntr), tl(ntr);
    std::vector<mom5::ocean_prog_tracer_type> T_prog(ntr);
    for (int n = 0; n < ntr; n++) {
        field[n] = rd(d + "field" + std::to_string(n) + ".bin", 3 * n3);
        th[n] = rd(d + "th" + std::to_string(n) + ".bin", n3);
        tl[n] = rd(d + "tl" + std::to_string(n) + ".bin", n3);
        wrk1[n].assign(n3, -777.0);
        T_prog[n].name = "tr" + std::to_string(n + 1);
        T_prog[n].field = field[n].data(); T_prog[n].th_tendency = th[n].data(); T_prog[n].wrk1 = wrk1[n].data();
        T_prog[n].tmask_limit = tl[n].data();
        T_prog[n].horz_advect_scheme = T_prog[n].vert_advect_scheme = scheme;
    }
    mom5::ocean_tracer_advect_nml nml;
    nml.advect_sweby_all = (mode == 0);
    nml.limit_with_upwind = (limit != 0);
    try {
        mom5::ocean_tracer_advect adv(G, D, ntr, nml);
        // the loop of update_ocean_tracer (ocean_tracer.F90:2295-2303)
        for (int n = 1; n <= ntr; n++) {
            adv.horz_advect_tracer(Time, A, Th, T_prog, T_prog[n - 1], n, dtime);
            if (mode == 1) wr(d + "out_horz_wrk1_" + std::to_string(n - 1) + ".bin", wrk1[n - 1].data(), n3);
            adv.vert_advect_tracer(Time, A, T_prog[n - 1]);
        }
    } catch (const std::exception &e) {
        std::cerr << e.what() << "\n";
        return 1;
    }
    for (int n = 0; n < ntr; n++) {
        wr(d + "out_th" + std::to_string(n) + ".bin", th[n].data(), n3);
        wr(d + "out_wrk1_" + std::to_string(n) + ".bin", wrk1[n].data(), n3);
    }
    return 0;
}
