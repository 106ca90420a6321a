#!/usr/bin/env python
"""bench.py -- Sweby MDFL tracer cell-updates/s (FP64) on N B200s, next to the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--case NAME] [--no-e2e] [--no-cpu]

One "step" = one advect_tracer_sweby_all pass (z, x, y sweeps + halo updates) over all tracers of the block.
Workload at N = 1: the synthetic 0.1-degree ACCESS-OM2-01-shaped grid 3600 x 2700 x 75, 3 tracers
(BASELINE.json configs[4], the configuration the metric is quoted on; it fits one GPU).  For N > 1 every rank
keeps the same 3600 x 2700 x 75 block (weak scaling), ranks are laid out with mpp_define_layout's rule and
exchange width-2 halos over NCCL.  Inputs are resident in HBM (far larger than L2, so no L2 flush is needed)
for `value`; `e2e` goes through the host-pointer C-ABI entry point with pinned host buffers.

Prints ONE JSON line (rank 0).  --impl reference times the CPU oracle (the restatement of the reference's
Fortran loops; the Fortran itself cannot be compiled in this image) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Sweby MDFL tracer cell-updates/s (FP64)"
UNIT = "cell-updates/s"


def b_alg(ntr: int) -> float:
    """algorithmic bytes per cell-update, SURVEY.md section 8d: 80 + 64/ntr"""
    return 80.0 + 64.0 / ntr


# per-sweep algorithmic bytes per CELL (all ntr tracers), SURVEY.md section 8d table
def sweep_bytes(ntr: int):
    return dict(z=16.0 * ntr + 16.0, x=24.0 * ntr + 16.0, y=40.0 * ntr + 32.0)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out = dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm), reasons=sorted(reasons))
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle in multi-block mode on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------
def cpu_sample(case: str, ntr: int, target_s: float, steps: int = 1, warmup: int = 0, cores=None):
    """Time the oracle (OpenMP over px*py blocks = stand-ins for MPI ranks) on the first `rows` rows of the
    workload.  Returns (cu_per_s, descriptor dict, per-step seconds list)."""
    from mom5_b200.domain import define_layout
    from mom5_b200.synthetic import make_case
    from oracle.oracle import Oracle, split_blocks

    cores = cores or len(os.sched_getaffinity(0))
    gen = make_case(case, ntr=ntr)
    s = gen.s
    # a band of the global grid: all of i, rows 1..rows, all k; flow scale fixed analytically (no global pass)
    s.flow_scale = s.cfl / 12.0
    # calibrate the band height on a thin band first
    nblk = max(cores, 1)

    def run(rows, nsteps, nwarm):
        band = make_case(case, ntr=ntr, nj=rows, tripolar=False, flow_scale=s.flow_scale)
        gb = band.block()
        px, py = define_layout(band.s.ni, rows, nblk)
        if rows // py < 4:
            px, py = nblk, 1
        dec = band.s.decomposition(px, py)
        blocks = split_blocks(dec, gb)
        o = Oracle(dec, blocks)
        T = [[t.numpy() for t in b.T] for b in blocks]
        th = [[t.numpy().copy() for t in b.th_tendency] for b in blocks]
        times = []
        for it in range(nwarm + nsteps):
            o.sweby_all_timed(T, th, band.s.dtime, nthreads=cores)
            dt = o.last_seconds
            if it >= nwarm:
                times.append(dt)
        cu = band.s.ni * rows * band.s.nk * ntr
        return cu, times, (px, py)

    rows0 = max(8, 2 * nblk // max(1, define_layout(s.ni, 8, nblk)[1]))
    rows0 = min(max(rows0, 16), s.nj)
    cu0, t0s, _ = run(rows0, 1, 1)
    rate0 = cu0 / min(t0s)
    per_step = target_s / max(steps + warmup, 1)
    rows = int(min(s.nj, max(rows0, per_step * rate0 / (s.ni * s.nk * ntr))))
    rows = max(rows0, min(rows, int(120e6 / (s.ni * s.nk))))   # bound the sample (and its generation time): <= 120 M cells
    cu, times, lay = run(rows, steps, warmup)
    med = sorted(times)[len(times) // 2]
    desc = dict(kind="port", cores=cores,
                sample=f"rows 1..{rows} of {s.ni}x{s.nj}x{s.nk} ({case}), {ntr} tracers, {lay[0]}x{lay[1]} blocks on {cores} threads, "
                       f"gcc -O2 -ffp-contract=off C restatement of OTA:4104-4511 (the Fortran cannot be compiled here)")
    return cu / med, desc, times


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on the host cores, same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec_ntr = args.ntr
    val, desc, times = cpu_sample(args.case, spec_ntr, target_s=max(20.0, 4.0 * (args.steps + args.warmup)),
                                  steps=args.steps, warmup=args.warmup)
    ms = 1e3 * sorted(times)[len(times) // 2]
    desc["value"] = val
    desc["unit"] = UNIT
    emit(dict(metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
              ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
              data="synthetic", impl="reference",
              config=dict(workload=f"{args.case}: Sweby MDFL advect_tracer_sweby_all, {spec_ntr} tracers; "
                                   f"CPU sample = {desc['sample']}"),
              cpu_baseline=desc,
              e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0)))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def generate_banded(gen, i0, i1, j0, j1, ntr, band=96):
    """Assemble one rank's BlockInputs on the device from j-bands (bounds the generator's temporaries)."""
    import torch
    from mom5_b200.synthetic import BlockInputs
    out = None
    for ja in range(j0, j1 + 1, band):
        jb = min(ja + band - 1, j1)
        b = gen.block(i0, i1, ja, jb, ntr=ntr)
        if out is None:
            nyd = j1 - j0 + 3

            def big(t):
                return torch.empty(t.shape[:-2] + (nyd, t.shape[-1]), dtype=t.dtype, device=t.device)

            out = BlockInputs(gen.s, i0, i1, j0, j1, {k: big(v) for k, v in b.grid2d.items()}, b.dzt, big(b.tmask),
                              big(b.rho_dzt), big(b.uhrho_et), big(b.vhrho_nt), big(b.wrho_bt), [big(t) for t in b.T], [],
                              [big(t) for t in b.th_tendency], [])
        r0 = ja - j0   # band's halo row jb'-1 sits at big row (ja-1) - (j0-1)
        n = jb - ja + 3
        for k in out.grid2d:
            out.grid2d[k][r0:r0 + n] = b.grid2d[k]
        for nm in ("tmask", "rho_dzt", "uhrho_et", "vhrho_nt", "wrho_bt"):
            getattr(out, nm)[:, r0:r0 + n] = getattr(b, nm)
        for q in range(ntr):
            out.T[q][:, r0:r0 + n] = b.T[q]
            out.th_tendency[q][:, r0:r0 + n] = b.th_tendency[q]
        del b
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from mom5_b200.api import Communicator, TracerAdvect
    from mom5_b200.domain import define_layout
    from mom5_b200.synthetic import CASES, Generator
    import dataclasses

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout = the one JSON line
        dist.init_process_group("nccl", device_id=dev)
        comm = Communicator.create_from_torch_distributed()

    base = CASES[args.case]
    ntr = args.ntr
    # weak scaling: every rank owns one base-sized block; the global grid is px x py such blocks
    px, py = define_layout(base.ni, base.nj, world)
    spec = dataclasses.replace(base, ni=base.ni * px, nj=base.nj * py, ntr=ntr, flow_scale=base.cfl / 12.0)
    gen = Generator(spec, device=dev)
    dec = spec.decomposition(px, py)
    i0, i1, j0, j1 = dec.extent(rank)
    t_setup = time.time()
    b = generate_banded(gen, i0, i1, j0, j1, ntr)
    torch.cuda.synchronize()
    adv = TracerAdvect(b, dec=dec, rank=rank, ntracers_max=ntr, comm=comm)
    T, th = b.T, b.th_tendency
    out = [torch.empty_like(t) for t in T]
    u, v, w, rho = b.uhrho_et, b.vhrho_nt, b.wrho_bt, b.rho_dzt
    del b.tmask
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup

    def step():
        adv.advect_tracer_sweby_all(T, th, out, u, v, w, rho, spec.dtime)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    l0 = adv.kernel_launches()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = dict(z=0.0, x=0.0, y=0.0, halo=0.0, total=0.0)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = adv.kernel_launches() - l0
    # per-sweep device time of the last step (CUDA events recorded by the library on the launching stream);
    # average over a few extra steps outside the timed region
    nph = 3
    for _ in range(nph):
        step()
        tm = adv.last_timing_ms()
        for k in phase:
            phase[k] += tm[k] / nph
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    cells = (i1 - i0 + 1) * (j1 - j0 + 1) * spec.nk
    cu_rank = cells * ntr
    value = world * cu_rank / (ms_step * 1e-3)

    peak, peak_src = measured_peaks()
    sb = sweep_bytes(ntr)
    fused = os.environ.get("MOM5ADV_FUSE", "1") != "0"
    if fused:
        # z sweep + ONE pass doing the x and y sweeps (k_sweby_xy): the pass does the algorithmic work of both sweeps
        # (SURVEY.md section 8d counts 24*ntr+16 + 40*ntr+32 B per cell for them) while moving only the y sweep's bytes;
        # phase "x" is the stand-alone x sweep on the 4 edge rows whose halo images the pass needs.
        sb = dict(z=sb["z"], xy=sb["x"] + sb["y"])
        phase_of = dict(z="z", xy="y")
        kernels = dict(z="k_sweby_z", xy="k_sweby_xy")
    else:
        phase_of = dict(z="z", x="x", y="y")
        kernels = dict(z="k_sweby_z", x="k_sweby_x", y="k_sweby_y")
    dom = max(sb, key=lambda k: phase[phase_of[k]])
    kname = kernels[dom]
    ach = cells * sb[dom] / (phase[phase_of[dom]] * 1e-3) / 1e9
    roofline = dict(bound="hbm", kernel=kname, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=None,
                    peak_source=peak_src,
                    per_sweep={k: dict(kernel=kernels[k], ms=phase[phase_of[k]], alg_bytes_per_cell=sb[k],
                                       achieved_gbs=cells * sb[k] / (phase[phase_of[k]] * 1e-3) / 1e9)
                               for k in sb},
                    halo_ms=phase["halo"],
                    whole_call=dict(alg_bytes_per_cell_update=b_alg(ntr), achieved_gbs=value / world * b_alg(ntr) / 1e9,
                                    frac=value / world * b_alg(ntr) / 1e9 / peak))
    if fused:
        roofline["per_sweep"]["xy"]["moved_bytes_per_cell"] = sweep_bytes(ntr)["y"]
        roofline["per_sweep"]["xy"]["moved_gbs"] = cells * sweep_bytes(ntr)["y"] / (phase["y"] * 1e-3) / 1e9
        roofline["edge_rows_x_ms"] = phase["x"]
    traffic_file = os.path.join(ROOT, "profiles", "traffic_r01.json")   # ncu --set full DRAM bytes per launch, by kernel name
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(kname)
        except Exception:
            pass

    res = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
               higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
               config=dict(workload=f"{args.case}: {base.ni}x{base.nj}x{base.nk} per GPU, {ntr} tracers, Sweby MDFL advect_tracer_sweby_all "
                                    f"({'z sweep + fused x/y pass' if fused else 'z,x,y sweeps'} + halo-2 updates), cyclic x + tripolar fold",
                           global_grid=[spec.ni, spec.nj, spec.nk], layout=[px, py], tracers=ntr,
                           l2="inputs (>100 GB) far exceed the 126 MB L2; no flush needed", fmad=False,
                           setup_s=round(t_setup, 1)),
               gpu_launches=int(launches), clocks=clocks, roofline=roofline, phase_ms=phase)

    # ---- secondary: update_advection_only stepping (advection + tracer time update + halo-1 update), device-resident ----
    if fused and world == 1:
        try:
            rhor = 1.0 / rho
            Tn = [torch.empty_like(t) for t in T]
            e2, e3, e4 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            thz = [torch.zeros_like(t) for t in T]
            for rep in range(2):   # rep 0 = warm-up
                e2.record()
                for _ in range(3):
                    adv.advect_sweby_all_and_update(T, Tn, rho, rhor, u, v, w, rho, spec.dtime)
                e3.record()
                for _ in range(3):
                    for t in thz:
                        t.zero_()
                    adv.advect_tracer_sweby_all(T, thz, out, u, v, w, rho, spec.dtime)
                    adv.tracer_update(T, thz, Tn, rho, rhor, spec.dtime)
                e4.record()
                torch.cuda.synchronize()
            res["advection_only_step"] = dict(
                what="update_advection_only (ocean_tracer.F90:2618-2649) with sweby_all: th=0; advect; field(taup1); halo-1 update",
                fused_epilogue_ms=e2.elapsed_time(e3) / 3, separate_passes_ms=e3.elapsed_time(e4) / 3,
                fused_cell_updates_per_s=cu_rank / (e2.elapsed_time(e3) / 3 * 1e-3))
            del Tn, thz, rhor
            torch.cuda.empty_cache()
        except Exception as ex:
            res["advection_only_step"] = dict(error=f"{type(ex).__name__}: {ex}")

    # ---- end-to-end through the host-pointer C ABI (pinned host buffers; H2D + D2H inside the timed region) ----
    if not args.no_e2e:
        try:
            res["e2e"] = run_e2e(args, adv, spec, b, T, th, out, u, v, w, rho, world, rank, dev, cells)
        except Exception as ex:  # report, never fake
            res["e2e"] = dict(value=None, unit=UNIT, error=f"{type(ex).__name__}: {ex}")
    adv.close()
    if rank == 0 and world == 1 and not args.no_cpu:
        val, desc, _ = cpu_sample(args.case, ntr, target_s=20.0, steps=3, warmup=1)
        desc["value"], desc["unit"] = val, UNIT
        res["cpu_baseline"] = desc
    if rank == 0:
        # flush NOW: with an eagerly initialised NCCL process group the interpreter can leave through the communicator teardown
        # without draining Python's stdout buffer (seen on the GPU box: exit status 0 and an empty JSON file)
        emit(res)
    if world > 1:
        dist.barrier()
        if comm:
            comm.destroy()
        dist.destroy_process_group()


def run_e2e(args, adv, spec, b, T, th, out, u, v, w, rho, world, rank, dev, cells):
    """Same metric through mom5adv_sweby_all (host pointers).  Per step: H2D of T, th_tendency (ntr each) and
    uhrho_et, vhrho_nt, wrho_bt, rho_dzt; D2H of th_tendency and adv_tendency (ntr each)."""
    import psutil
    import torch
    import torch.distributed as dist
    ntr = len(T)
    n3 = T[0].numel()
    nw = w.numel()
    need = (2 * ntr + 3) * n3 * 8 + nw * 8 + ntr * n3 * 8
    avail = psutil.virtual_memory().available
    if world > 1 or need * 1.3 * max(world, 1) > avail:
        raise RuntimeError(f"host buffers need {need / 1e9:.0f} GB/rank, {avail / 1e9:.0f} GB available"
                           if world == 1 else "e2e is measured at N=1 only (one PCIe root per rank)")
    pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
    hT = [pin(t) for t in T]
    hth = [pin(t) for t in th]
    hu, hv, hw, hr = pin(u), pin(v), pin(w), pin(rho)
    hout = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in T]
    # free the device-resident copies: the host-pointer entry point owns its own device mirrors
    for lst in (T, th, out):
        lst.clear()
    del u, v, w, rho
    b.T.clear(); b.th_tendency.clear()
    b.uhrho_et = b.vhrho_nt = b.wrho_bt = b.rho_dzt = None
    torch.cuda.empty_cache()
    npT = [t.numpy() for t in hT]; npth = [t.numpy() for t in hth]; npo = [t.numpy() for t in hout]
    nu, nv, nw_, nr = hu.numpy(), hv.numpy(), hw.numpy(), hr.numpy()
    steps = max(2, min(args.steps, 3))
    adv.advect_tracer_sweby_all(npT, npth, npo, nu, nv, nw_, nr, spec.dtime)   # warm-up (allocates the mirrors)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        adv.advect_tracer_sweby_all(npT, npth, npo, nu, nv, nw_, nr, spec.dtime)   # synchronous on return
    dt = (time.perf_counter() - t0) / steps
    h2d = (2 * ntr + 3) * n3 * 8 + nw * 8
    d2h = 2 * ntr * n3 * 8
    return dict(value=cells * ntr / dt, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h), ms_per_step=dt * 1e3,
                steps=steps, api="mom5adv_sweby_all (host pointers, pinned; copy pipeline over "
                          + ("j-bands" if os.environ.get("MOM5ADV_BANDED", "1") != "0" and os.environ.get("MOM5ADV_FUSE", "1") != "0" else "tracers") + ")", pcie_gbs=(h2d + d2h) / dt / 1e9)


_REAL_STDOUT = None


def emit(obj):
    """the ONE JSON line, on the process's original stdout"""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    # stdout carries exactly one JSON line.  Native libraries write there too (NCCL prints its version banner on stdout when
    # NCCL_DEBUG is set in the environment, whatever NCCL_DEBUG_FILE says): keep a private handle on the real stdout and point
    # fd 1 at stderr for everything else.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--case", default="global_01deg")
    ap.add_argument("--ntr", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
