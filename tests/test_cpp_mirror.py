"""The C++ host-side mirror of the reference interface (include/ocean_tracer_advect.hpp): compiles and links against
libmom5adv.so on CPU; on the GPU box the compiled driver runs the dispatchers exactly as update_ocean_tracer does and
must reproduce the reference-text golden vectors bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from tests.util import assert_bit_equal, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from mom5_b200 import build
    build.build()
    exe = str(tmp_path / "mirror_driver")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "mirror_driver.cpp"),
           "-o", exe, "-L", os.path.join(ROOT, "mom5_b200"), "-lmom5adv", "-Wl,-rpath," + os.path.join(ROOT, "mom5_b200")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "--link-check"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "mom5adv 200"


def _dump(b, d, mode, scheme, limit):
    s = b.spec
    with open(d / "meta.txt", "w") as f:
        f.write(f"{b.ni} {b.nj} {b.nk} {len(b.T)} {int(s.cyclic_x)} {int(s.cyclic_y)} {int(s.tripolar)} {s.dtime!r} {mode} {scheme} {limit}\n")
    w = lambda name, a: np.ascontiguousarray(a.numpy() if hasattr(a, "numpy") else a, dtype=np.float64).tofile(d / (name + ".bin"))
    for k, v in b.grid2d.items():
        w(k, v)
    w("dzt", b.dzt); w("tmask", b.tmask); w("u", b.uhrho_et); w("v", b.vhrho_nt); w("w", b.wrho_bt)
    rho = b.rho_dzt.numpy()
    w("rho", np.stack([rho * 0.97, rho, rho * 1.01]))            # only the tau level is read
    for n in range(len(b.T)):
        w(f"field{n}", np.stack([b.T[n].numpy(), b.T_tau[n].numpy(), np.zeros_like(b.T[n].numpy())]))
        w(f"th{n}", b.th_tendency[n]); w(f"tl{n}", b.tmask_limit[n])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["g_tripolar", "g_walls"])
def test_cpp_mirror_sweby_all_matches_golden(tmp_path, name):
    exe = _build(tmp_path)
    b, gold, _ = load_golden(name)
    _dump(b, tmp_path, mode=0, scheme=9, limit=0)
    r = subprocess.run([exe, str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    shp = tuple(b.T[0].shape)
    for n in range(len(b.T)):
        assert_bit_equal(np.fromfile(tmp_path / f"out_th{n}.bin").reshape(shp), gold[f"sweby_all.th_tendency.{n + 1}"], f"th[{n}]")
        assert_bit_equal(np.fromfile(tmp_path / f"out_wrk1_{n}.bin").reshape(shp), gold[f"sweby_all.wrk1.{n + 1}"], f"wrk1[{n}]")


@pytest.mark.gpu
@pytest.mark.parametrize("tag,scheme,limit", [("quicker", 5, 0), ("quicker_lim", 5, 1), ("upwind", 1, 0), ("mdfl_sweby", 9, 0)])
def test_cpp_mirror_dispatcher_arms_match_golden(tmp_path, tag, scheme, limit):
    exe = _build(tmp_path)
    b, gold, _ = load_golden("g_tripolar")
    n = int(gold[f"{tag}.tracer"]) - 1
    # one-tracer problem: the golden arm was run on tracer n+1
    b.T, b.T_tau, b.th_tendency, b.tmask_limit = [b.T[n]], [b.T_tau[n]], [b.th_tendency[n]], [b.tmask_limit[n]]
    _dump(b, tmp_path, mode=1, scheme=scheme, limit=limit)
    r = subprocess.run([exe, str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    shp = tuple(b.T[0].shape)
    assert_bit_equal(np.fromfile(tmp_path / "out_horz_wrk1_0.bin").reshape(shp), gold[f"{tag}.horz.wrk1"], "horz wrk1")
    assert_bit_equal(np.fromfile(tmp_path / "out_wrk1_0.bin").reshape(shp), gold[f"{tag}.vert.wrk1"], "vert wrk1")
    assert_bit_equal(np.fromfile(tmp_path / "out_th0.bin").reshape(shp), gold[f"{tag}.vert.th_tendency"], "th after horz+vert")


@pytest.mark.gpu
def test_cpp_mirror_invalid_scheme_is_fatal(tmp_path):
    exe = _build(tmp_path)
    b, _, _ = load_golden("g_walls")
    _dump(b, tmp_path, mode=1, scheme=3, limit=0)      # ADVECT_4TH_ORDER: not on this path
    r = subprocess.run([exe, str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "chose invalid horz advection scheme" in r.stderr
