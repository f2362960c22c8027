#!/usr/bin/env python
"""bench.py -- particle tetra-crossings per second of the orbit-pusher hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun by the driver)
    python bench.py --impl reference --steps K --warmup W    (reference arm: the CPU restatement, all host cores)

A "step" is one batched orbit_timestep_gorilla call (t_step of physical time) over all particles of the
rank.  Particles shard across ranks with the mesh replicated; the only collective is the final reduction of
the counters (NCCL all-reduce through torch.distributed).  `value` = all pushes of all ranks in the K timed
steps / max-over-ranks device time (CUDA events), particle state resident in HBM.  `e2e` = the same metric
through the host-buffer C ABI (pinned host arrays, H2D + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np


def _claim_stdout():
    """stdout carries exactly ONE line, the JSON result.  Libraries write there too (NCCL prints its version banner to
    stdout at NCCL_DEBUG=VERSION and ignores NCCL_DEBUG_FILE at that level), so the process keeps a private copy of the
    original stdout for the result and points file descriptor 1 at stderr for everything else."""
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return out


ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "particle_tetra_crossings_per_second"
UNIT = "crossings/s"
# the shipped library is the strict build; GORILLA_B200_LIB=.../libgorilla_b200_fma.so (GORILLA_VARIANT=fma
# GORILLA_NVCC_EXTRA=--fmad=true python -m gorilla_b200.build) is a measurement-only variant that is NOT bit-exact
FP_MODE = ("fma (--fmad=true, measurement-only variant, not bit-exact)" if "_fma" in os.environ.get("GORILLA_B200_LIB", "")
           else "strict (--fmad=false, bit-exact vs oracle)")
BYTES_PER_CROSSING = {False: 344.0, True: 488.0}  # hot record (+8 B topology), without/with Phi part
BYTES_STRONG_E = 192.0  # + 24 doubles of the strong-electric-field group
# FP64 thread-instructions (DADD+DMUL+DFMA) and DRAM bytes per crossing of the strict build, from one ncu capture per
# kernel (profiles/r01_ncu_per_crossing.json; order 1 not captured: order-2 figure)
FP64_INST_PER_CROSSING = {1: 463.0, 2: 463.0, 3: 2026.0, 4: 3349.0, "rk4": 767.0}
try:
    _NCU = json.loads((ROOT / "profiles" / "r01_ncu_per_crossing.json").read_text())
except Exception:  # the table is documentation of a capture; the bench runs without it (traffic: null)
    _NCU = {}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "vmec_qi", "analytic", "west_soledge3x", "efit_rect", "efit_flux"])
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (0 = workload default)")
    ap.add_argument("--poly-order", type=int, default=0, help="0 = workload default")
    ap.add_argument("--ipusher", type=int, default=0, help="1 = RK4 pusher, 2 = polynomial pusher (0 = workload default)")
    ap.add_argument("--time-tracing", type=int, default=0, choices=[0, 1, 2],
                    help="i_time_tracing_option: 1 = dt/dtau constant per cell, 2 = Hamiltonian time (0 = workload default)")
    ap.add_argument("--optional-quantities", action="store_true",
                    help="also form pusher_tetra_poly's optional quantities (t_hamiltonian, gyrophase, vpar_int, vpar2_int)")
    ap.add_argument("--adaptive", type=float, default=0.0,
                    help="boole_adaptive_time_steps with this desired_delta_energy (max_n_intermediate_steps = 10000)")
    ap.add_argument("--t-step", type=float, default=0.0, help="physical time per step [s] (0 = workload default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sort", type=int, default=1, help="re-sort particles by tetra index before the timed region")
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-group", action="store_true", help="orders 3/4: 4-warp CTAs instead of the lock-step solver kernel")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workloads
def make_workload(name: str):
    """Returns dict(name, grid, settings, particles(n, seed) -> (x, vpar, vperp), n_default, t_step, desc)."""
    import workloads
    vmec_file = ROOT / "data" / "equilibria" / "netcdf_file_for_test.nc"
    if name == "auto":
        name = "vmec_qi" if (vmec_file.exists() and hasattr(workloads, "vmec_qi")) else "analytic"
    if name == "vmec_qi":
        grid, settings = workloads.vmec_qi(str(vmec_file))
        return dict(name="vmec_qi_alpha_3.5MeV_100x40x40", grid=grid, settings=settings,
                    particles=workloads.particles_vmec_alpha, n_default=1_000_000, t_step=1.0e-4,
                    desc="QI stellarator netcdf_file_for_test.nc (VMEC), grid_kind=3 100x40x40, 3.5 MeV alphas, "
                         "s0=0.5, pitch U[-1,1], time step 1e-4 s (BASELINE config 3: 100 steps of 1e-4 s)")
    data = ROOT / "data" / "equilibria"
    if name == "west_soledge3x":
        grid, settings = workloads.west_soledge3x(data, n2=60)
        return dict(name="west_soledge3x_W74_600keV_strongE_rk4", grid=grid, settings=settings,
                    particles=lambda n, seed: workloads.particles_on_triangles(data, n, seed), n_default=1_000_000,
                    t_step=1.0e-4,
                    desc="BASELINE config 4: WEST equilibrium + SOLEDGE3X-EIRENE mesh, grid_kind=4, n2=60 (4 242 060 "
                         "tetrahedra), strong-electric-field mode eps_Phi=-1.5e-5, 600 keV W74+ uniform over the "
                         "poloidal mesh (scrape-off-layer starts are lost in the first steps), RK4 pusher, steps of 1e-4 s")
    if name == "efit_flux":
        grid, settings = workloads.efit_flux(data)
        return dict(name="efit_aug_flux_D_3keV_100x40x40", grid=grid, settings=settings,
                    particles=lambda n, seed: workloads.particles_flux(n, seed), n_default=1_000_000, t_step=1.0e-4,
                    desc="BASELINE configs 1/2: ASDEX Upgrade g_file_for_test, grid_kind=2 coord_system=2 field-aligned "
                         "100x40x40 (960 000 tetrahedra), 3 keV deuterons, s in U[0.2,0.9], steps of 1e-4 s")
    if name == "efit_rect":
        from gorilla_b200 import GorillaSettings, TetraGridSettings
        grid = TetraGridSettings(grid_kind=1, n1=100, n2=40, n3=160, boole_n_field_periods=True,
                                 g_file_filename=str(data / "g_file_for_test"),
                                 convex_wall_filename=str(data / "convex_wall_for_test.dat"))
        settings = GorillaSettings(eps_Phi=0.0, coord_system=1, ispecies=2, boole_periodic_relocation=True, ipusher=2,
                                   poly_order=2, boole_guess=True)
        return dict(name="efit_aug_rect_D_3keV_100x40x160", grid=grid, settings=settings,
                    particles=lambda n, seed: workloads.particles_cyl(n, seed, R0=165.0, a=45.0), n_default=1_000_000,
                    t_step=2.0e-5,
                    desc="BASELINE config 1/2 geometry in cylindrical coordinates: ASDEX Upgrade g_file_for_test, grid_kind=1 "
                         "100x40x160 (3 840 000 tetrahedra), 3 keV deuterons")
    grid, settings = workloads.analytic_tokamak(40, 80, 40)
    settings.poly_order = 2
    return dict(name="analytic_tokamak_D_3keV_40x80x40", grid=grid, settings=settings,
                particles=lambda n, seed: workloads.particles_cyl(n, seed), n_default=1_000_000, t_step=2.0e-5,
                desc="EXAMPLES/example_8 analytic circular tokamak, grid_kind=5 40x80x40 (768000 tetrahedra), "
                     "3 keV deuterons")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self._stop = [], set(), threading.Event()
        self.index = index
        self.max_mhz = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [s.strip() for s in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self.th.start()

    def stop(self):
        self._stop.set()
        self.th.join(timeout=6)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_run(wl, mesh, settings, n_sample, t_step, steps, warmup, nthreads):
    """The CPU restatement (oracle) with `nthreads` OpenMP threads on a bounded sample of the workload."""
    from oracle_binding import OracleMesh
    om = OracleMesh(mesh, settings)
    x, vpar, vperp = wl["particles"](n_sample, 12345)
    import workloads
    st = workloads.fresh_state(n_sample)
    om.orbit_timestep_batch(x, vpar, vperp, 0.0, *st, nthreads=nthreads)  # localise
    for _ in range(warmup):
        om.orbit_timestep_batch(x, vpar, vperp, t_step, *st, nthreads=nthreads)
    t0 = time.perf_counter()
    pushes = 0
    for _ in range(steps):
        pushes += om.orbit_timestep_batch(x, vpar, vperp, t_step, *st, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return pushes / dt, dt, pushes


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gorilla_b200 import build_mesh
    wl = make_workload(args.workload)
    settings = wl["settings"]
    if args.poly_order:
        settings.poly_order = args.poly_order
    if args.ipusher:
        settings.ipusher = args.ipusher
    if args.time_tracing:
        settings.i_time_tracing_option = args.time_tracing
    if args.adaptive > 0.0:
        settings.boole_adaptive_time_steps = True
        settings.desired_delta_energy = args.adaptive
    if getattr(args, "optional_quantities", False):
        settings.boole_time_Hamiltonian = settings.boole_gyrophase = settings.boole_vpar_int = settings.boole_vpar2_int = True
    t_step = args.t_step or wl["t_step"]
    mesh = build_mesh(wl["grid"], settings)
    cores = os.cpu_count() or 1
    n_sample = 2000 * cores   # ~1.5 s of CPU work per step on 16 cores
    value, dt, pushes = cpu_run(wl, mesh, settings, n_sample, t_step, args.steps, args.warmup, cores)
    sample = f"{n_sample} particles x {args.steps} steps of {t_step:g} s ({pushes} pushes, {dt:.1f} s wall)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "poly_order": settings.poly_order, "t_step_s": t_step, "desc": wl["desc"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference (oracle/), OpenMP over particles; the Fortran "
                                 "reference cannot be compiled in this image (no gfortran)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from gorilla_b200 import Gorilla, build_mesh, launch_count
    from gorilla_b200.api import fp64_peak
    import workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = make_workload(args.workload)
    settings = wl["settings"]
    if args.poly_order:
        settings.poly_order = args.poly_order
    if args.ipusher:
        settings.ipusher = args.ipusher
    if args.time_tracing:
        settings.i_time_tracing_option = args.time_tracing
    if args.adaptive > 0.0:
        settings.boole_adaptive_time_steps = True
        settings.desired_delta_energy = args.adaptive
    if getattr(args, "optional_quantities", False):
        settings.boole_time_Hamiltonian = settings.boole_gyrophase = settings.boole_vpar_int = settings.boole_vpar2_int = True
    t_step = args.t_step or wl["t_step"]
    n = args.particles or wl["n_default"]

    t0 = time.perf_counter()
    mesh = build_mesh(wl["grid"], settings)        # host, once, replicated on every rank
    t_mesh = time.perf_counter() - t0
    g = Gorilla(mesh, settings)
    if args.ctas_per_sm or args.threads:
        g.set_launch_config(args.ctas_per_sm, args.threads)
    if args.no_group:
        g._debug_use_group(False)
    has_phi = bool(np.any(mesh.tetra_physics[:, 116:125] != 0.0))
    strong = bool(settings.boole_strong_electric_field)
    bytes_per_crossing = BYTES_PER_CROSSING[has_phi or strong] + (BYTES_STRONG_E if strong else 0.0)
    ext = settings.ipusher == 2 and settings.i_time_tracing_option == 2
    if ext or args.optional_quantities:
        bytes_per_crossing += 64.0   # hamiltonian_time record (8 doubles) read at the end of every push

    # particles of this rank (weak scaling: n per GPU fixed); independent streams per rank
    x, vpar, vperp = wl["particles"](n, 1000 + rank)
    binit, ind, ifc = workloads.fresh_state(n)
    tt = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    xd, vd, wd, bd, it, fd = tt(x), tt(vpar), tt(vperp), tt(binit), tt(ind), tt(ifc)
    npd = torch.zeros(n, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    g.orbit_timestep_gorilla_dev(xd, vd, wd, 0.0, bd, it, fd, stream=stream)   # localisation (find_tetra), untimed
    find_ms = g.counters().find_ms
    n_located = int((it > 0).sum())

    oqd = torch.zeros((n, 4), dtype=torch.float64, device=dev) if args.optional_quantities else None

    def step_dev():
        if oqd is not None:
            g.orbit_timestep_gorilla_optional_dev(xd, vd, wd, t_step, bd, it, fd, oqd, n_pushes=npd, stream=stream)
        else:
            g.orbit_timestep_gorilla_dev(xd, vd, wd, t_step, bd, it, fd, n_pushes=npd, stream=stream)

    def resort():
        perm = torch.empty(n, dtype=torch.int64, device=dev)
        g.sort_permutation_dev(it, perm, stream=stream)
        return [a[perm].contiguous() for a in (xd, vd, wd, bd, it, fd)]

    for _ in range(args.warmup):
        if args.sort:
            xd, vd, wd, bd, it, fd = resort()
        step_dev()
    torch.cuda.synchronize()

    # ---- timed region: K steps, device time via CUDA events on the launch stream, max over ranks
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pushes = 0
    kernel_ms = 0.0
    fallback = np.zeros(4, np.int64)
    n_adaptive = 0
    ev0.record()
    for _ in range(args.steps):
        if args.sort:
            xd, vd, wd, bd, it, fd = resort()
        step_dev()
        c = g.counters()           # synchronises the stream; reads the device counters of this step
        pushes += c.n_pushes
        kernel_ms += c.kernel_ms
        fallback += np.array(c.n_fallback)
        n_adaptive += c.n_adaptive
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    launches = launch_count() - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    n_lost = int((it < 1).sum())

    red = torch.tensor([float(pushes), float(n_lost), float(n)], dtype=torch.float64, device=dev)
    tmax = torch.tensor([elapsed_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.SUM)     # the path's only collective: counters
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tot_pushes, tot_lost, tot_n = (float(v) for v in red.tolist())
    elapsed_ms_max, kernel_ms_max = (float(v) for v in tmax.tolist())
    value = tot_pushes / (elapsed_ms_max * 1e-3)

    # ---- e2e through the host-buffer C ABI (pinned host memory, copies in the timed region)
    e2e = None
    if not args.no_e2e:
        hx = torch.empty((n, 3), dtype=torch.float64).pin_memory()
        hv, hw = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
        hb, hi, hf = (torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(3))
        hnp = torch.empty(n, dtype=torch.int64).pin_memory()
        for h, d in ((hx, xd), (hv, vd), (hw, wd), (hb, bd), (hi, it), (hf, fd)):
            h.copy_(d)
        torch.cuda.synchronize()
        nx, nv, nw, nb, ni, nf, nnp = (t.numpy() for t in (hx, hv, hw, hb, hi, hf, hnp))
        e_steps = max(1, min(args.steps, 3))
        g.orbit_timestep_gorilla(nx, nv, nw, t_step, nb, ni, nf, n_pushes=nnp)  # warm
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e_push = 0
        for _ in range(e_steps):
            g.orbit_timestep_gorilla(nx, nv, nw, t_step, nb, ni, nf, n_pushes=nnp)
            e_push += int(nnp.sum())
        e_dt = time.perf_counter() - t0
        er = torch.tensor([float(e_push)], dtype=torch.float64, device=dev)
        et = torch.tensor([e_dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(er, op=dist.ReduceOp.SUM)
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        h2d = n * (3 * 8 + 8 + 8 + 4 + 4 + 4)
        d2h = h2d + n * 8 + n * 8
        e2e = {"value": float(er.item()) / float(et.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e_steps}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # dominant kernel = orbit_kernel<K,PHI>: one launch per step; algorithmic bytes = bytes/crossing x crossings
        per_rank_pushes = pushes / max(1, args.steps)
        launch_ms = kernel_ms / max(1, args.steps)
        achieved = bytes_per_crossing * per_rank_pushes / (launch_ms * 1e-3) / 1e9
        # the slower of the two per-push limits decides the bound (north_star): HBM gather vs FP64 issue
        dfma_peak, muladd_peak = fp64_peak()
        kkey = "rk4" if settings.ipusher == 1 else settings.poly_order
        fp64_per = FP64_INST_PER_CROSSING[kkey]
        # traffic: DRAM bytes per launch = per-crossing figure of the ncu capture of this kernel x crossings per launch
        ncu = _NCU.get(str(kkey))
        traffic = ncu["dram_bytes_per_crossing"] * per_rank_pushes if ncu and not has_phi and not strong and not ext else None
        traffic_src = (f"ncu capture {ncu['capture']}: {ncu['dram_bytes_per_crossing']:.2f} B/crossing x crossings per launch"
                       if traffic is not None else None)
        t_hbm, t_fp64 = bytes_per_crossing / (hbm_peak * 1e9), fp64_per / muladd_peak
        fp64_ach = fp64_per * per_rank_pushes / (launch_ms * 1e-3)
        fp64 = {"achieved": fp64_ach / 1e12, "peak": muladd_peak / 1e12, "unit": "Tinst/s (thread-level DMUL/DADD)",
                "frac": fp64_ach / muladd_peak, "inst_per_crossing": fp64_per, "dfma_peak": dfma_peak / 1e12,
                "peak_source": "measured in this run (gorilla_b200_fp64_peak)"}
        kern = f"orbit_kernel<{0 if settings.ipusher == 1 else settings.poly_order},{2 if strong else 1 if has_phi else 0}{',EXT=2' if args.optional_quantities else ',EXT=1' if ext else ',EXT=3' if settings.boole_adaptive_time_steps else ''}>"
        if t_hbm >= t_fp64:
            roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "kernel": kern,
                        "algorithmic_bytes_per_crossing": bytes_per_crossing, "launch_ms": launch_ms,
                        "kernel_share_of_step": kernel_ms / elapsed_ms, "peak_source": peak_src, "fp64": fp64}
        else:
            roofline = {"bound": "fp64", "achieved": fp64["achieved"], "peak": fp64["peak"], "unit": fp64["unit"],
                        "frac": fp64["frac"], "traffic": traffic, "traffic_source": traffic_src, "kernel": kern,
                        "launch_ms": launch_ms,
                        "kernel_share_of_step": kernel_ms / elapsed_ms, "peak_source": fp64["peak_source"],
                        "hbm": {"achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                                "algorithmic_bytes_per_crossing": bytes_per_crossing}}
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_s = 2000 * cores
            v, dt, p = cpu_run(wl, mesh, settings, n_s, t_step, max(1, min(args.steps, 3)), 1, cores)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{n_s} particles x {max(1, min(args.steps, 3))} steps of {t_step:g} s "
                                      f"({p} pushes, {dt:.1f} s wall)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms_max / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "desc": wl["desc"], "ipusher": settings.ipusher,
                       "poly_order": settings.poly_order, "i_time_tracing_option": settings.i_time_tracing_option,
                       "boole_adaptive_time_steps": bool(settings.boole_adaptive_time_steps),
                       "optional_quantities": bool(args.optional_quantities),
                       "desired_delta_energy": settings.desired_delta_energy if settings.boole_adaptive_time_steps else None,
                       "particles_per_gpu": n, "t_step_s": t_step, "ntetr": mesh.ntetr,
                       "mesh_hot_bytes": int(mesh.ntetr * (352 + (160 if has_phi else 0))),
                       "l2_policy": "inputs_larger_than_l2 (mesh hot records > 126 MB, gathered at random)",
                       "parallelism": f"particles sharded over {world} GPU(s), mesh replicated",
                       "sort_by_tetra_each_step": bool(args.sort), "fp_mode": FP_MODE},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "counters": {"pushes": tot_pushes, "lost": tot_lost, "particles": tot_n,
                         "fallback_rank0": [int(v) for v in fallback], "adaptive_pushes_rank0": int(n_adaptive), "located_rank0": n_located,
                         "find_tetra_ms_rank0": find_ms, "mesh_build_s_rank0": t_mesh},
        }
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    g.close()
    if world > 1:
        dist.destroy_process_group()


RESULT_OUT = sys.stdout

if __name__ == "__main__":
    RESULT_OUT = _claim_stdout()
    main()
