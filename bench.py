#!/usr/bin/env python
"""bench.py -- Sweby MDFL tracer cell-updates/s (FP64) on N B200s, next to the CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--case NAME] [--ntr T]
                    [--scaling strong|weak|replicate] [--no-e2e] [--no-cpu] [--no-extras]

One "step" = one advect_tracer_sweby_all pass (z sweep, fused x/y pass, halo-2 updates) over all tracers of the block.

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): the synthetic 0.1-degree ACCESS-OM2-01-shaped
grid 3600 x 2700 x 75, cyclic in x, tripolar fold in the north, 3 tracers.
  N = 1 : the whole grid on one GPU (it fits).
  N > 1 : default --scaling strong: the SAME global grid split over the N GPUs with mpp_define_layout2D's rule
          ((2,1), (2,2), (2,4) blocks; 1800 x 675 x 75 per GPU at N = 8) -- "0.1-degree ... 2/4/8 B200 scaling".  The line also
          carries `weak`: SURVEY section 8e's weak-scaling test, 1800 x 675 x 75 per GPU replicated (the N = 1 line carries the
          same block on one GPU, so the weak efficiency can be read off the lines).  --scaling replicate is round 1's mode
          (every rank keeps a whole 3600 x 2700 x 75 block).
Inputs are resident in HBM (far larger than L2: no flush needed) for `value`; `e2e` goes through the host-pointer C-ABI entry
point with pinned host buffers.  Before anything is timed the run VERIFIES itself (`parity_check`): a 1-degree tripolar case is
advected through the same communicator / layout rule and the all-reduced mpp_chksum of every tendency is compared with the
checksums committed in tests/golden/parity_chksums.json (produced by the CPU oracle, pinned by a CPU test); a mismatch is fatal.

Prints ONE JSON line (rank 0).  --impl reference times the reference's CPU path -- the C restatement of the Fortran loops in
multi-block mode (the Fortran itself cannot be compiled in this image) -- on the SAME grid, fold and tracer count, on all host cores.
"""
from __future__ import annotations

import argparse
import dataclasses
import hashlib
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
ROUND = "r02"


def b_alg(ntr: int) -> float:
    """algorithmic bytes per cell-update, SURVEY.md section 8d: 80 + 64/ntr"""
    return 80.0 + 64.0 / ntr


def sweep_bytes(ntr: int):
    """per-sweep algorithmic bytes per CELL (all ntr tracers), SURVEY.md section 8d table"""
    return dict(z=16.0 * ntr + 16.0, x=24.0 * ntr + 16.0, y=40.0 * ntr + 32.0)


def workload_name(case: str, spec, ntr: int) -> str:
    """the same string in both arms (the driver compares the arms' configs)"""
    geo = ("cyclic x + tripolar fold" if spec.tripolar else
           "cyclic x + cyclic y" if (spec.cyclic_x and spec.cyclic_y) else "cyclic x" if spec.cyclic_x else "solid walls")
    return f"{case}: {spec.ni}x{spec.nj}x{spec.nk} global grid, {ntr} tracers, Sweby MDFL advect_tracer_sweby_all, {geo}"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash() -> str:
    """hash of the kernel sources (csrc/*.cuh: every kernel of the hot path lives in a header; capi.cu is the host driver):
    a traffic file measured on other kernels is refused"""
    h = hashlib.sha1()
    d = os.path.join(ROOT, "mom5_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith(".cuh"):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


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
# CPU arm: the oracle's multi-block driver (blocks <-> MPI ranks, OpenMP threads <-> ranks, strip copies <-> mpp_update_domains)
# ------------------------------------------------------------------------------------------------
def _gen_device():
    """inputs are generated on the GPU when there is one (seconds instead of minutes) -- generation is not the measured path"""
    try:
        import torch
        if torch.cuda.is_available():
            return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    except Exception:
        pass
    return "cpu"


def cpu_blocks(spec, ntr, cores):
    """every rank's BlockInputs of `spec` on the host, one block per host thread, generated block by block"""
    import torch
    from mom5_b200.domain import define_layout
    from mom5_b200.synthetic import BlockInputs, Generator
    px, py = define_layout(spec.ni, spec.nj, cores)
    if spec.nj // py < 4 or spec.ni // px < 4:
        px, py = (cores, 1) if spec.ni >= spec.nj else (1, cores)
    dec = spec.decomposition(px, py)
    dev = _gen_device()
    gen = Generator(spec, device=dev)
    blocks = []
    for r in range(px * py):
        i0, i1, j0, j1 = dec.extent(r)
        b = gen.block(i0, i1, j0, j1, ntr=ntr)
        c = lambda t: t.cpu()
        blocks.append(BlockInputs(spec, i0, i1, j0, j1, {k: c(v) for k, v in b.grid2d.items()}, c(b.dzt), c(b.tmask), c(b.rho_dzt),
                                  c(b.uhrho_et), c(b.vhrho_nt), c(b.wrho_bt), [c(t) for t in b.T], [], [c(t) for t in b.th_tendency], []))
        del b
    if dev != "cpu":
        torch.cuda.empty_cache()
    return dec, blocks


def cpu_verify_multiblock(cores) -> bool:
    """multi-block == single-block, bitwise, on a small tripolar case (the decomposition invariance the timed driver relies on)"""
    import numpy as np
    from mom5_b200.domain import define_layout
    from mom5_b200.synthetic import make_case
    from oracle.oracle import Oracle, gather, split_blocks
    g = make_case("mini_tripolar")
    gb = g.block()
    o1 = Oracle(g.s.decomposition(1, 1), [gb])
    th1 = [[t.numpy().copy() for t in gb.th_tendency]]
    r1 = o1.sweby_all_timed([[t.numpy() for t in gb.T]], th1, g.s.dtime, nthreads=1)
    px, py = define_layout(g.s.ni, g.s.nj, min(cores, 8))
    dec = g.s.decomposition(px, py)
    bl = split_blocks(dec, gb)
    o = Oracle(dec, bl)
    th = [[t.numpy().copy() for t in b.th_tendency] for b in bl]
    r = o.sweby_all_timed([[t.numpy() for t in b.T] for b in bl], th, g.s.dtime, nthreads=cores)
    for n in range(len(gb.T)):
        a = gather(dec, [r["adv"][q][n] for q in range(len(bl))])
        if not np.array_equal(a.view(np.int64), r1["adv"][0][n][:, 1:-1, 1:-1].view(np.int64)):
            return False
    return True


def cpu_time(spec, ntr, steps, warmup, cores):
    """(cell-updates/s, description, per-step seconds) of the oracle on the whole grid of `spec`"""
    from oracle.oracle import Oracle
    t0 = time.time()
    dec, blocks = cpu_blocks(spec, ntr, cores)
    o = Oracle(dec, blocks)
    T = [[t.numpy() for t in b.T] for b in blocks]
    th = [[t.numpy() for t in b.th_tendency] for b in blocks]
    t_setup = time.time() - t0
    times = []
    for it in range(warmup + steps):
        o.sweby_all_timed(T, th, spec.dtime, nthreads=cores)
        if it >= warmup:
            times.append(o.last_seconds)
    med = sorted(times)[len(times) // 2]
    cu = spec.ni * spec.nj * spec.nk * ntr
    desc = dict(kind="port", cores=cores,
                sample=f"whole {spec.ni}x{spec.nj}x{spec.nk} grid ({'tripolar fold on' if spec.tripolar else 'no fold'}), {ntr} tracers, "
                       f"{dec.px}x{dec.py} blocks on {cores} threads (blocks stand in for MPI ranks, strip copies for mpp_update_domains); "
                       f"gcc -O2 -ffp-contract=off C restatement of OTA:4104-4511 (the Fortran cannot be compiled in this image)",
                blocks=[dec.px, dec.py], setup_s=round(t_setup, 1))
    return cu / med, desc, times


def host_mem_needed(spec, ntr) -> float:
    """bytes the CPU arm keeps resident: inputs (T, th per tracer; u, v, w, rho, tmask) + oracle scratch (tm h2, adv per tracer; mask h2)"""
    cells = spec.ni * spec.nj * spec.nk
    return cells * 8.0 * (2 * ntr + 5 + 2 * ntr + 1) * 1.12


def reference_spec(case, ntr):
    """the spec the CPU arm runs: the whole grid when the host has the memory for it, else the tallest tripolar band that fits"""
    import psutil
    from mom5_b200.synthetic import CASES
    base = dataclasses.replace(CASES[case], ntr=ntr)
    base = dataclasses.replace(base, flow_scale=base.cfl / 12.0)
    avail = psutil.virtual_memory().available
    need = host_mem_needed(base, ntr)
    if need <= 0.9 * avail:
        return base, True, need, avail
    rows = int(base.nj * 0.9 * avail / need) // 4 * 4
    rows = max(rows, 16)
    return dataclasses.replace(base, nj=rows), False, need, avail


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on the host cores, same grid / fold / tracers / metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mom5_b200.synthetic import CASES
    cores = len(os.sched_getaffinity(0))
    full = dataclasses.replace(CASES[args.case], ntr=args.ntr)
    spec, same, need, avail = reference_spec(args.case, args.ntr)
    ok = cpu_verify_multiblock(cores)
    val, desc, times = cpu_time(spec, args.ntr, args.steps, args.warmup, cores)
    ms = 1e3 * sorted(times)[len(times) // 2]
    desc["value"], desc["unit"] = val, UNIT
    desc["multiblock_equals_singleblock"] = ok
    if not same:
        desc["sample"] = (f"host memory ({avail / 1e9:.0f} GB available, {need / 1e9:.0f} GB needed for the whole grid) bounds the sample: " + desc["sample"])
    emit(dict(metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
              ms_per_step=ms, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
              data="synthetic", impl="reference",
              config=dict(workload=workload_name(args.case, full, args.ntr), tracers=args.ntr, same_grid_as_gpu_arm=same,
                          global_grid=[spec.ni, spec.nj, spec.nk]),
              cpu_baseline=desc,
              e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0)))
    if not ok:
        sys.exit(3)


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


class Env:
    """process-group plumbing of one bench process"""

    def __init__(self):
        import torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.comm = None
        if self.world > 1:
            import torch.distributed as dist
            from mom5_b200.api import Communicator
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout = the one JSON line
            dist.init_process_group("nccl", device_id=self.dev)
            self.comm = Communicator.create_from_torch_distributed()

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        import torch
        if self.world == 1:
            return x
        import torch.distributed as dist
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_int64(self, vals):
        import torch
        t = torch.tensor(vals, device=self.dev, dtype=torch.int64)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM)     # wrap-around int64 sum == mpp_chksum over all PEs
        return [int(v) for v in t.cpu()]


def parity_check(env):
    """Decomposition-invariance self-check on THIS communicator, before anything is timed: the 1-degree tripolar case
    (360 x 300 x 50, 3 tracers) goes through the default driver with the layout mpp_define_layout2D gives for this rank count; the
    mpp_chksum (mpp_chksum_int.h:20-38) of every th_tendency / adv_tendency, summed over the ranks, must equal the CPU oracle's on
    the same inputs (rank 0 generates the global block, copies it to the host and runs the oracle there as the checker).
    tests/golden/parity_chksums.json pins the same case for CPU-generated inputs (device and host libm differ in the last bit of
    sin / cos / exp, so generated inputs are only comparable within one device class)."""
    import torch
    from mom5_b200.api import TracerAdvect
    from mom5_b200.domain import define_layout
    from mom5_b200.synthetic import CASES, BlockInputs, Generator
    base = CASES["global_1deg"]
    spec = dataclasses.replace(base, ntr=3, flow_scale=base.cfl / 12.0)
    px, py = define_layout(spec.ni, spec.nj, env.world)
    dec = spec.decomposition(px, py)
    i0, i1, j0, j1 = dec.extent(env.rank)
    gen = Generator(spec, device=env.dev)
    b = gen.block(i0, i1, j0, j1, ntr=3)
    adv = TracerAdvect(b, dec=dec, rank=env.rank, ntracers_max=3, comm=env.comm)
    th = [t.clone() for t in b.th_tendency]
    out = [torch.empty_like(t) for t in b.T]
    adv.advect_tracer_sweby_all(b.T, th, out, b.uhrho_et, b.vhrho_nt, b.wrho_bt, b.rho_dzt, spec.dtime)
    torch.cuda.synchronize()
    got = env.sum_int64([adv.chksum(t) for t in th] + [adv.chksum(t) for t in out])
    adv.close()
    res = dict(case="global_1deg 360x300x50, 3 tracers, default driver", layout=[px, py], chksum=dict(th=got[:3], adv=got[3:]))
    if env.rank == 0:
        from oracle.oracle import Oracle
        gb = gen.block(1, spec.ni, 1, spec.nj, ntr=3)
        c = lambda t: t.cpu()
        hb = BlockInputs(spec, 1, spec.ni, 1, spec.nj, {k: c(v) for k, v in gb.grid2d.items()}, c(gb.dzt), c(gb.tmask), c(gb.rho_dzt),
                         c(gb.uhrho_et), c(gb.vhrho_nt), c(gb.wrho_bt), [c(t) for t in gb.T], [], [c(t) for t in gb.th_tendency], [])
        del gb
        o = Oracle(spec.decomposition(1, 1), [hb])
        tho = [[t.numpy().copy() for t in hb.th_tendency]]
        ref = o.sweby_all_timed([[t.numpy() for t in hb.T]], tho, spec.dtime, nthreads=1)
        want = [o.chksum([tho[0][n]]) for n in range(3)] + [o.chksum([ref["adv"][0][n]]) for n in range(3)]
        res["oracle_chksum"] = dict(th=want[:3], adv=want[3:])
        res["ok"] = (got == want)
    ok = env.sum_int64([1 if (env.rank != 0 or res["ok"]) else 0])[0] == env.world
    res["ok"] = ok
    return res


def timed_steps(env, step, steps, warmup):
    """W warm-up steps, then EXACTLY K steps between barrier + synchronize; CUDA events; max over ranks -> ms per step"""
    import torch
    for _ in range(warmup):
        step()
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    env.barrier()
    return env.max_over_ranks(e0.elapsed_time(e1)) / steps


def make_rank_block(env, spec, px, py, ntr):
    from mom5_b200.synthetic import Generator
    gen = Generator(spec, device=env.dev)
    dec = spec.decomposition(px, py)
    i0, i1, j0, j1 = dec.extent(env.rank)
    return dec, generate_banded(gen, i0, i1, j0, j1, ntr)


def run_block(env, spec, px, py, ntr, steps, warmup, want_phases=False, sampler=None):
    """device-resident sweby_all on this rank's block of `spec` split px x py; returns dict(ms, value, cells_rank, phase, launches)"""
    import torch
    from mom5_b200.api import TracerAdvect
    t0 = time.time()
    dec, b = make_rank_block(env, spec, px, py, ntr)
    torch.cuda.synchronize()
    adv = TracerAdvect(b, dec=dec, rank=env.rank, ntracers_max=ntr, comm=env.comm)
    T, th = b.T, b.th_tendency
    out = [torch.empty_like(t) for t in T]
    u, v, w, rho = b.uhrho_et, b.vhrho_nt, b.wrho_bt, b.rho_dzt
    b.tmask = None
    torch.cuda.synchronize()
    setup = time.time() - t0

    def step():
        adv.advect_tracer_sweby_all(T, th, out, u, v, w, rho, spec.dtime)

    for _ in range(warmup):
        step()
    env.barrier()
    l0 = adv.kernel_launches()
    smp = sampler() if sampler else None
    ms = timed_steps(env, step, steps, 0)
    launches = adv.kernel_launches() - l0
    phase = dict(z=0.0, x=0.0, y=0.0, halo=0.0, total=0.0)
    if want_phases:
        nph = 3
        for _ in range(nph):
            step()
            tm = adv.last_timing_ms()
            for k in phase:
                phase[k] += tm[k] / nph
    clocks = smp.stop() if smp else None
    cells = b.ni * b.nj * spec.nk
    cells_all = spec.ni * spec.nj * spec.nk
    res = dict(ms=ms, value=cells_all * ntr / (ms * 1e-3), cells_rank=cells, phase=phase, launches=int(launches), setup_s=round(setup, 1),
               clocks=clocks, block=[b.ni, b.nj, spec.nk])
    keep = dict(adv=adv, b=b, T=T, th=th, out=out, u=u, v=v, w=w, rho=rho)
    return res, keep


def release(keep):
    import torch
    keep["adv"].close()
    keep.clear()
    torch.cuda.empty_cache()


def traffic_record(kname, cells=None, ntr=None):
    """ncu --set full DRAM bytes per launch of `kname` at the bench workload, from this round's profile file -- refused when the
    file was measured on different kernel sources"""
    p = os.path.join(ROOT, "profiles", f"traffic_{ROUND}.json")
    if not os.path.exists(p):
        return None, f"profiles/traffic_{ROUND}.json absent"
    try:
        d = json.load(open(p))
    except Exception as ex:
        return None, f"unreadable: {ex}"
    if d.get("kernel_source_sha1") != kernel_source_hash():
        return None, f"stale: measured on sources {d.get('kernel_source_sha1')}, current {kernel_source_hash()}"
    t = d.get("kernels", {}).get(kname)
    src = "ncu --set full, " + d.get("source", "")
    if t is not None and cells and ntr and d.get("cells_per_launch"):
        if ntr != d.get("tracers", ntr):
            return None, f"measured with {d.get('tracers')} tracers, this run has {ntr}"
        if cells != d["cells_per_launch"]:      # a smaller block per GPU: the streaming traffic of a launch scales with its cells
            t = t * cells / d["cells_per_launch"]
            src = f"scaled by cells per launch ({cells} / {d['cells_per_launch']}) from: " + src
    return t, src


def config2_extras(env):
    """BASELINE config 2 (gyre1 / torus1-shaped 256 x 256 x 50, quicker + MDFL, one B200): device time per tracer of the dispatcher
    arms, CUDA events over 20 repetitions; algorithmic bytes = every array the arm must touch once per call."""
    import torch
    from mom5_b200.api import ADVECT_MDFL_SWEBY, ADVECT_QUICKER, ADVECT_UPWIND, TracerAdvect
    from mom5_b200.synthetic import CASES, Generator
    peak, _ = measured_peaks()
    out = {}
    for case in ("torus", "gyre"):
        base = CASES[case]
        spec = dataclasses.replace(base, ntr=2, flow_scale=base.cfl / 12.0)
        b = Generator(spec, device=env.dev).block(1, spec.ni, 1, spec.nj, ntr=2, with_tau=True)
        adv = TracerAdvect(b, ntracers_max=2, limit_with_upwind=True)
        cells = spec.ni * spec.nj * spec.nk
        T, Tt, tl = b.T[0], b.T_tau[0], b.tmask_limit[0]
        th = b.th_tendency[0].clone()
        wrk = torch.empty_like(th)
        u, v, w, rho = b.uhrho_et, b.vhrho_nt, b.wrho_bt, b.rho_dzt
        arms = dict(
            quicker_horz_plus_vert=(lambda: (adv.horz_advect_tracer(ADVECT_QUICKER, T, th, wrk, u, v, spec.dtime, T_tau=Tt, tmask_limit=tl),
                                             adv.vert_advect_tracer(ADVECT_QUICKER, T, th, wrk, w, T_tau=Tt, tmask_limit=tl)),
                                    # horz: R T(taum1), T(tau), u, v, tmask_limit, th; W th, wrk1.  vert: R T(taum1), T(tau), w, tmask_limit, th; W th, wrk1
                                    8.0 * (8 + 7)),
            upwind_horz_plus_vert=(lambda: (adv.horz_advect_tracer(ADVECT_UPWIND, T, th, wrk, u, v),
                                            adv.vert_advect_tracer(ADVECT_UPWIND, T, th, wrk, w)),
                                   8.0 * (6 + 5)),
            mdfl_sweby_one_tracer=(lambda: adv.horz_advect_tracer(ADVECT_MDFL_SWEBY, T, th, wrk, u, v, spec.dtime, wrho_bt=w, rho_dzt=rho),
                                   b_alg(1)))
        res = {}
        for nm, (fn, bytes_cu) in arms.items():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            res[nm] = dict(ms=ms, cell_updates_per_s=cells / (ms * 1e-3), alg_bytes_per_cell_update=bytes_cu,
                           alg_gbs=cells * bytes_cu / (ms * 1e-3) / 1e9, frac_of_hbm=cells * bytes_cu / (ms * 1e-3) / 1e9 / peak)
        adv.close()
        out[case] = dict(grid=[spec.ni, spec.nj, spec.nk], **res)
    out["note"] = ("5.2 MB per array: the whole working set sits in the 126 MB L2 and a call is a handful of ~10 us kernels, so these are "
                   "launch / latency numbers, not HBM numbers")
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from mom5_b200.domain import define_layout
    from mom5_b200.synthetic import CASES

    env = Env()
    world, rank = env.world, env.rank
    base = CASES[args.case]
    ntr = args.ntr
    base = dataclasses.replace(base, ntr=ntr, flow_scale=base.cfl / 12.0)
    mode = args.scaling if world > 1 else "single"

    # ---- self-verification before anything is timed ----
    pc = parity_check(env)
    if not pc["ok"]:
        if rank == 0:
            emit(dict(metric=METRIC, value=None, unit=UNIT, n_gpus=world, error="parity_check failed", parity_check=pc))
        sys.exit(4)

    # ---- headline ----
    if mode == "replicate":          # round 1's mode: every rank owns a whole base-sized block
        px, py = define_layout(base.ni, base.nj, world)
        spec = dataclasses.replace(base, ni=base.ni * px, nj=base.nj * py)
    else:                            # single GPU, or strong scaling: the base grid split over the ranks
        px, py = define_layout(base.ni, base.nj, world)
        spec = base
    sampler = (lambda: ClockSampler(env.local)) if rank == 0 else None
    head, keep = run_block(env, spec, px, py, ntr, args.steps, args.warmup, want_phases=True, sampler=sampler)
    ms_step, value, cells, phase = head["ms"], head["value"], head["cells_rank"], head["phase"]

    peak, peak_src = measured_peaks()
    sb = sweep_bytes(ntr)
    fused = os.environ.get("MOM5ADV_FUSE", "1") != "0"
    tma_bits = int(os.environ.get("MOM5ADV_TMA", "3"))
    tma, tma_z = bool(tma_bits & 2), bool(tma_bits & 1)
    if fused:
        # z sweep + ONE pass doing the x and y sweeps: the pass does the algorithmic work of both sweeps (SURVEY.md section 8d counts
        # 24*ntr+16 + 40*ntr+32 B per cell for them) while moving only the y sweep's bytes; phase "x" is the stand-alone x sweep on the
        # 4 edge rows whose halo images the pass needs.
        sb = dict(z=sb["z"], xy=sb["x"] + sb["y"])
        phase_of = dict(z="z", xy="y")
        kernels = dict(z="k_sweby_z_tma" if tma_z else "k_sweby_z", xy="k_sweby_xy_tma" if tma else "k_sweby_xy")
    else:
        phase_of = dict(z="z", x="x", y="y")
        kernels = dict(z="k_sweby_z_tma" if tma_z else "k_sweby_z", x="k_sweby_x", y="k_sweby_y")
    dom = max(sb, key=lambda k: phase[phase_of[k]])
    kname = kernels[dom]
    dom_ms = max(phase[phase_of[dom]], 1e-9)
    ach = cells * sb[dom] / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = traffic_record(kname, cells, ntr)
    roofline = dict(bound="hbm", kernel=kname, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=traffic,
                    traffic_source=traffic_src, peak_source=peak_src,
                    per_sweep={k: dict(kernel=kernels[k], ms=phase[phase_of[k]], alg_bytes_per_cell=sb[k],
                                       achieved_gbs=cells * sb[k] / (max(phase[phase_of[k]], 1e-9) * 1e-3) / 1e9)
                               for k in sb},
                    halo_ms=phase["halo"],
                    whole_call=dict(alg_bytes_per_cell_update=b_alg(ntr), achieved_gbs=value / world * b_alg(ntr) / 1e9,
                                    frac=value / world * b_alg(ntr) / 1e9 / peak))
    if traffic:
        # the same kernel on REAL bytes: measured DRAM traffic per launch / live kernel time
        roofline["real_dram_gbs"] = traffic / (dom_ms * 1e-3) / 1e9
        roofline["real_dram_frac"] = traffic / (dom_ms * 1e-3) / 1e9 / peak
    if fused:
        roofline["per_sweep"]["xy"]["moved_bytes_per_cell"] = sweep_bytes(ntr)["y"]
        roofline["per_sweep"]["xy"]["moved_gbs"] = cells * sweep_bytes(ntr)["y"] / (max(phase["y"], 1e-9) * 1e-3) / 1e9
        roofline["edge_rows_x_ms"] = phase["x"]

    res = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
               higher_is_better=True, scaling=("weak" if mode == "replicate" else "strong"), vs_baseline=None, dtype="f64", data="synthetic",
               config=dict(workload=workload_name(args.case, base, ntr), mode=mode,
                           driver=("z sweep + fused x/y pass" if fused else "z, x, y sweeps") + (", TMA staging" if (tma and tma_z) else ", LDGSTS staging" if not (tma or tma_z) else f", TMA staging bits {tma_bits}"),
                           global_grid=[spec.ni, spec.nj, spec.nk], layout=[px, py], block_per_gpu=head["block"], tracers=ntr,
                           l2="inputs per GPU far exceed the 126 MB L2; no flush needed", fmad=False, setup_s=head["setup_s"]),
               gpu_launches=head["launches"], clocks=head["clocks"], roofline=roofline, phase_ms=phase, parity_check=pc)

    extras = not args.no_extras
    # ---- advection-only stepping (time update fused into the pass), N = 1 ----
    if extras and fused and world == 1:
        try:
            adv, T, out, u, v, w, rho = (keep[k] for k in ("adv", "T", "out", "u", "v", "w", "rho"))
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
                fused_cell_updates_per_s=cells * ntr / (e2.elapsed_time(e3) / 3 * 1e-3))
            del Tn, thz, rhor
            torch.cuda.empty_cache()
        except Exception as ex:
            res["advection_only_step"] = dict(error=f"{type(ex).__name__}: {ex}")

    # ---- end-to-end through the host-pointer C ABI (pinned host buffers; H2D + D2H inside the timed region), N = 1 ----
    if not args.no_e2e:
        try:
            res["e2e"] = run_e2e(args, env, spec, keep, cells)
        except Exception as ex:  # report, never fake
            res["e2e"] = dict(value=None, unit=UNIT, error=f"{type(ex).__name__}: {ex}")
    release(keep)

    # ---- the same entry point with PAGEABLE caller arrays (what the Fortran model hands over), 0.25-degree grid, N = 1 ----
    if extras and world == 1 and not args.no_e2e:
        try:
            res["e2e_pageable"] = run_e2e_pageable(env)
        except Exception as ex:
            res["e2e_pageable"] = dict(error=f"{type(ex).__name__}: {ex}")

    # ---- SURVEY section 8e weak-scaling test: 1800 x 675 x 75 per GPU (the 0.1-degree grid / 8), replicated ----
    if extras and args.case == "global_01deg":
        try:
            wpx, wpy = define_layout(base.ni, base.nj, world)
            bni, bnj = base.ni // 2, base.nj // 4
            wspec = dataclasses.replace(base, ni=bni * wpx, nj=bnj * wpy)
            wk, keep2 = run_block(env, wspec, wpx, wpy, ntr, args.steps, args.warmup, want_phases=True)
            release(keep2)
            res["weak"] = dict(what="SURVEY 8e weak-scaling test: fixed 1800x675x75 block per GPU, global grid = layout x block",
                               block_per_gpu=wk["block"], global_grid=[wspec.ni, wspec.nj, wspec.nk], layout=[wpx, wpy],
                               ms_per_step=wk["ms"], value=wk["value"], per_gpu_value=wk["value"] / world, phase_ms=wk["phase"])
        except Exception as ex:
            res["weak"] = dict(error=f"{type(ex).__name__}: {ex}")

    # ---- BASELINE config 2: quicker / upwind / per-tracer MDFL arms at 256 x 256 x 50, N = 1 ----
    if extras and world == 1:
        try:
            res["config2"] = config2_extras(env)
        except Exception as ex:
            res["config2"] = dict(error=f"{type(ex).__name__}: {ex}")

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, N = 1) ----
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cores = len(os.sched_getaffinity(0))
            rows = max(64, int(120e6 / (base.ni * base.nk)) // 4 * 4)        # <= 120 M cells: ~1 s per step, ~20 GB of host memory
            band = dataclasses.replace(base, nj=min(rows, base.nj))
            val, desc, _ = cpu_time(band, ntr, steps=3, warmup=1, cores=cores)
            desc["value"], desc["unit"] = val, UNIT
            desc["sample"] = f"bounded sample: a {band.nj}-row tripolar band of the workload -- " + desc["sample"]
            res["cpu_baseline"] = desc
        except Exception as ex:
            res["cpu_baseline"] = dict(error=f"{type(ex).__name__}: {ex}")
    if rank == 0:
        # flush NOW: with an eagerly initialised NCCL process group the interpreter can leave through the communicator teardown
        # without draining Python's stdout buffer (seen on the GPU box: exit status 0 and an empty JSON file)
        emit(res)
    if world > 1:
        dist.barrier()
        if env.comm:
            env.comm.destroy()
        dist.destroy_process_group()


def run_e2e(args, env, spec, keep, cells):
    """Same metric through mom5adv_sweby_all (host pointers).  Per step: H2D of T (ntr) and uhrho_et, vhrho_nt, wrho_bt, rho_dzt;
    D2H of adv_tendency (ntr); th_tendency += adv_tendency is formed on the HOST by the library (it never crosses the link)."""
    import psutil
    import torch
    world = env.world
    T, th, out, u, v, w, rho = (keep[k] for k in ("T", "th", "out", "u", "v", "w", "rho"))
    b = keep["b"]
    adv = keep["adv"]
    ntr = len(T)
    n3 = T[0].numel()
    nw = w.numel()
    need = (2 * ntr + 3) * n3 * 8 + nw * 8 + ntr * n3 * 8
    avail = psutil.virtual_memory().available
    if world > 1 or need * 1.3 > avail:
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
    for k in ("u", "v", "w", "rho"):
        keep[k] = None
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
    tr = adv.last_transfer_bytes()
    return dict(value=cells * ntr / dt, unit=UNIT, h2d_bytes_per_step=int(tr[0]), d2h_bytes_per_step=int(tr[1]), ms_per_step=dt * 1e3,
                steps=steps, api="mom5adv_sweby_all (host pointers, pinned; copy pipeline over j-bands; bytes counted by the library)",
                pcie_gbs=(tr[0] + tr[1]) / dt / 1e9)


def run_e2e_pageable(env):
    """mom5adv_sweby_all on plain (pageable) numpy arrays, 1440 x 1080 x 50, 3 tracers: the first call page-locks the caller's arrays
    (cudaHostRegister, cached by address -- the model's arrays live for the whole run), later calls run the same pipeline as with
    pinned buffers.  MOM5ADV_PIN=0 would leave every copy staged through the driver's bounce buffer."""
    import numpy as np
    import torch
    from mom5_b200.api import TracerAdvect
    from mom5_b200.synthetic import CASES, Generator
    base = CASES["global_025deg"]
    spec = dataclasses.replace(base, ntr=3, flow_scale=base.cfl / 12.0)
    b = Generator(spec, device=env.dev).block(1, spec.ni, 1, spec.nj, ntr=3)
    host = lambda t: np.array(t.cpu().numpy(), copy=True)      # fresh pageable allocations
    T = [host(t) for t in b.T]; th = [host(t) for t in b.th_tendency]; out = [np.empty_like(t) for t in T]
    u, v, w, rho = host(b.uhrho_et), host(b.vhrho_nt), host(b.wrho_bt), host(b.rho_dzt)
    adv = TracerAdvect(b, ntracers_max=3)
    del b
    torch.cuda.empty_cache()
    cu = spec.ni * spec.nj * spec.nk * 3
    t0 = time.perf_counter()
    adv.advect_tracer_sweby_all(T, th, out, u, v, w, rho, spec.dtime)
    first = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(3):
        adv.advect_tracer_sweby_all(T, th, out, u, v, w, rho, spec.dtime)
    steady = (time.perf_counter() - t0) / 3
    tr = adv.last_transfer_bytes()
    adv.close()
    return dict(grid=[spec.ni, spec.nj, spec.nk], first_call_ms=first * 1e3, steady_ms=steady * 1e3, value=cu / steady, unit=UNIT,
                h2d_bytes_per_step=int(tr[0]), d2h_bytes_per_step=int(tr[1]), pcie_gbs=(tr[0] + tr[1]) / steady / 1e9,
                note="first call includes cudaHostRegister of the caller's 14 arrays and the allocation of the device mirrors")


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
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak", "replicate"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.scaling == "weak":
        args.scaling = "replicate"
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
