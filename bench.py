#!/usr/bin/env python
"""bench.py -- particle tetra-crossings per second of the orbit-pusher hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun by the driver)
    python bench.py --impl reference --steps K --warmup W    (reference arm: the CPU restatement, all host cores)

A "step" is one batched orbit_timestep_gorilla call (t_step of physical time) over all particles of the
rank, preceded by the library's re-sort of the batch by tetrahedron.  Particles shard across ranks with the mesh
replicated; the only exchange is the reduction of counters / conservation diagnostics / timings at the end, done by
the library itself over its NCCL communicator (gorilla_b200_diag_reduce_dev, gorilla_b200_comm_allreduce_f64);
torch.distributed only carries the NCCL id to the ranks and provides the barrier.  `value` = all pushes of all
ranks in the K timed steps / max-over-ranks device time (CUDA events), particle state resident in HBM.  `e2e` = the
same metric through the host-buffer C ABI (pinned host arrays, H2D + D2H inside the timed region).
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
# kernel and workload (profiles/r02_ncu_per_crossing.json, falling back to round 1's; order 1 not captured: order-2 figure)
FP64_INST_PER_CROSSING = {1: 463.0, 2: 463.0, 3: 2026.0, 4: 3349.0, "rk4": 767.0}
_NCU = {}
for _f in ("r01_ncu_per_crossing.json", "r02_ncu_per_crossing.json"):
    try:
        _NCU.update(json.loads((ROOT / "profiles" / _f).read_text()))
    except Exception:  # the tables document captures; the bench runs without them (traffic: null)
        pass


def ncu_entry(workload: str, kkey):
    """per-crossing figures of the ncu capture of this kernel ON THIS WORKLOAD ('<workload>:<kernel>'), else the kernel's
    capture on the default workload when that is the workload being run"""
    e = _NCU.get(f"{workload}:{kkey}")
    if e is None and workload.startswith("vmec_qi_alpha") and "spread" not in workload:
        e = _NCU.get(str(kkey))
    return e


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
    ap.add_argument("--i-precomp", type=int, default=0, choices=[0, 1, 2], help="polynomial pusher: coefficients from the precomputed poly4 record")
    ap.add_argument("--newton-precalc", action="store_true", help="RK pusher: boole_newton_precalc")
    ap.add_argument("--ode45", action="store_true", help="RK pusher: boole_pusher_ode45 (RKF45 instead of RK4)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--eps-phi", type=float, default=None, help="electrostatic potential strength eps_Phi of the mesh (PHI = 1 kernels); default: the workload's")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sort", type=int, default=1, help="re-sort particles by tetra index before every step")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --particles per GPU; strong: --total-particles sharded contiguously [rN/G,(r+1)N/G) (BASELINE config 5)")
    ap.add_argument("--total-particles", type=int, default=10_000_000, help="strong scaling: particles of the whole job")
    ap.add_argument("--prefetch", type=int, default=-1, choices=[-1, 0, 1], help="neighbour-record L2 prefetch: -1 library default (off), 0 off, 1 on")
    ap.add_argument("--gather", type=int, default=-1, choices=[-1, 0, 1, 2], help="record gather: -1 library default, 0 vector loads, 1 per-lane bulk copies (TMA), 2 warp-cooperative cp.async copies")
    ap.add_argument("--start", default="default", choices=["default", "spread"],
                    help="vmec_qi: 'spread' starts s in U[0.15, 0.95] instead of on s = 0.5 (records touched exceed the L2)")
    ap.add_argument("--no-variants", action="store_true", help="skip the short K=4 / RK4 / spread-start variant runs")
    ap.add_argument("--no-l2-flush", action="store_true", help="do not overwrite the L2 between timed steps")
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-group", action="store_true", help="orders 3/4: 4-warp CTAs instead of the lock-step solver kernel")
    ap.add_argument("--rebin", type=int, default=-1, choices=[-1, 0, 1], help="orders 3/4: re-bin the root solves by solver mode (-1 = library default)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workloads
def make_workload(name: str, start: str = "default"):
    """Returns dict(name, grid, settings, particles(n, seed) -> (x, vpar, vperp), n_default, t_step, desc)."""
    import workloads
    vmec_file = ROOT / "data" / "equilibria" / "netcdf_file_for_test.nc"
    if name == "auto":
        name = "vmec_qi" if (vmec_file.exists() and hasattr(workloads, "vmec_qi")) else "analytic"
    if name == "vmec_qi":
        grid, settings = workloads.vmec_qi(str(vmec_file))
        if start == "spread":
            return dict(name="vmec_qi_alpha_3.5MeV_100x40x40_spread", grid=grid, settings=settings,
                        particles=workloads.particles_vmec_alpha_spread, n_default=1_000_000, t_step=1.0e-4,
                        desc="QI stellarator netcdf_file_for_test.nc (VMEC), grid_kind=3 100x40x40, 3.5 MeV alphas, "
                             "s0 in U[0.15,0.95] (the batch touches the whole mesh), pitch U[-1,1], time step 1e-4 s")
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


def apply_args(settings, args):
    if args.poly_order:
        settings.poly_order = args.poly_order
    if args.ipusher:
        settings.ipusher = args.ipusher
    if args.time_tracing:
        settings.i_time_tracing_option = args.time_tracing
    if args.adaptive > 0.0:
        settings.boole_adaptive_time_steps = True
        settings.desired_delta_energy = args.adaptive
    if getattr(args, "i_precomp", 0):
        settings.i_precomp = args.i_precomp
    if getattr(args, "newton_precalc", False):
        settings.boole_newton_precalc = True
    if getattr(args, "ode45", False):
        settings.boole_pusher_ode45 = True
    if getattr(args, "eps_phi", None) is not None:
        settings.eps_Phi = args.eps_phi
    if getattr(args, "optional_quantities", False):
        settings.boole_time_Hamiltonian = settings.boole_gyrophase = settings.boole_vpar_int = settings.boole_vpar2_int = True
    return settings


def make_config(wl, settings, args, world, n, t_step, mesh):
    """The `config` object: identical for the GPU arm and the reference arm (the driver compares them)."""
    strong = args.scaling == "strong"
    has_phi = bool(np.any(mesh.tetra_physics[:, 116:125] != 0.0))
    strong_e = bool(settings.boole_strong_electric_field)
    hot_rec = 352 + (160 if has_phi or strong_e else 0) + (256 if strong_e else 0)
    return {"workload": wl["name"], "desc": wl["desc"], "ipusher": settings.ipusher,
            "poly_order": settings.poly_order, "eps_Phi": float(settings.eps_Phi),
            "i_time_tracing_option": settings.i_time_tracing_option,
            "boole_adaptive_time_steps": bool(settings.boole_adaptive_time_steps),
            "optional_quantities": bool(args.optional_quantities), "i_precomp": int(settings.i_precomp),
            "boole_newton_precalc": bool(settings.boole_newton_precalc), "boole_pusher_ode45": bool(settings.boole_pusher_ode45),
            "desired_delta_energy": settings.desired_delta_energy if settings.boole_adaptive_time_steps else None,
            "particles_per_gpu": None if strong else n, "total_particles": args.total_particles if strong else n * world,
            "t_step_s": t_step, "ntetr": mesh.ntetr, "mesh_hot_bytes": int(mesh.ntetr * hot_rec),
            "l2_policy": "no flush" if args.no_l2_flush else
                         "L2 overwritten between timed steps (256 MB device memset before every step); what the gather then "
                         "finds in the L2 is what the step itself brought in (see `footprint` and roofline.dram_frac)",
            "prefetch": int(args.prefetch), "gather": int(args.gather),
            "parallelism": f"particles sharded over {world} GPU(s) ({'contiguous shards of a fixed total' if strong else 'fixed count per GPU'}), mesh replicated",
            "sort_by_tetra_each_step": bool(args.sort), "fp_mode": FP_MODE}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gorilla_b200 import build_mesh
    wl = make_workload(args.workload, args.start)
    settings = apply_args(wl["settings"], args)
    t_step = args.t_step or wl["t_step"]
    mesh = build_mesh(wl["grid"], settings)
    cores = os.cpu_count() or 1
    n_sample = 2000 * cores   # ~1.5 s of CPU work per step on 16 cores
    value, dt, pushes = cpu_run(wl, mesh, settings, n_sample, t_step, args.steps, args.warmup, cores)
    sample = f"{n_sample} particles x {args.steps} steps of {t_step:g} s ({pushes} pushes, {dt:.1f} s wall)"
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    n = args.particles or wl["n_default"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the workload the metric is quoted on (same object as the GPU arm's); the CPU arm times a bounded SAMPLE of it
        # (crossings/s is intensive): see cpu_baseline.sample
        "config": make_config(wl, settings, args, world, n, t_step, mesh),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "sample_particles": n_sample,
                         "note": "C restatement of the reference (oracle/), OpenMP over particles; the Fortran "
                                 "reference cannot be compiled in this image (no gfortran)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class Resident:
    """A batch resident in HBM on one handle: state tensors, reference invariants, the step and re-sort calls."""

    def __init__(self, g, x, vpar, vperp, dev, optional_quantities=False):
        import torch
        import workloads
        self.g, self.n, self.dev = g, x.shape[0], dev
        binit, ind, ifc = workloads.fresh_state(self.n)
        tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        self.x, self.vpar, self.vperp, self.init, self.ind, self.iface = tt(x), tt(vpar), tt(vperp), tt(binit), tt(ind), tt(ifc)
        self.npush = torch.zeros(self.n, dtype=torch.int64, device=dev)
        self.oq = torch.zeros((self.n, 4), dtype=torch.float64, device=dev) if optional_quantities else None
        self.stream = torch.cuda.current_stream().cuda_stream
        g.orbit_timestep_gorilla_dev(self.x, self.vpar, self.vperp, 0.0, self.init, self.ind, self.iface, stream=self.stream)
        self.find_ms = g.counters().find_ms
        self.e0, self.p0, self.m0 = (torch.empty(self.n, dtype=torch.float64, device=dev) for _ in range(3))
        g.invariants_dev(self.x, self.vpar, self.vperp, self.ind, self.e0, self.p0, self.m0, stream=self.stream)
        g.diag_reset(stream=self.stream)

    def resort(self):
        # one library call: radix sort by tetrahedron + in-place permutation of the state and of the reference invariants
        self.g.resort_dev(self.x, self.vpar, self.vperp, self.init, self.ind, self.iface, extra=(self.e0, self.p0, self.m0),
                          stream=self.stream)

    def step(self, t_step):
        if self.oq is not None:
            self.g.orbit_timestep_gorilla_optional_dev(self.x, self.vpar, self.vperp, t_step, self.init, self.ind, self.iface,
                                                       self.oq, n_pushes=self.npush, stream=self.stream)
        else:
            self.g.orbit_timestep_gorilla_dev(self.x, self.vpar, self.vperp, t_step, self.init, self.ind, self.iface,
                                              n_pushes=self.npush, stream=self.stream)

    def diag(self):
        return self.g.diag_reduce_dev(self.x, self.vpar, self.vperp, self.ind, self.e0, self.p0, self.m0, stream=self.stream)


def timed_steps(res, t_step, steps, warmup, sort, flush_buf, barrier=None, sampler=None):
    """W warm-up + K timed steps; device time by CUDA events on the launch stream.  Returns per-rank numbers."""
    import torch
    from gorilla_b200 import launch_count
    for _ in range(warmup):
        if sort:
            res.resort()
        res.step(t_step)
    torch.cuda.synchronize()
    res.g.diag_reset(stream=res.stream)
    if barrier:
        barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    launches0 = launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pushes, kernel_ms, fallback, n_adaptive = 0, 0.0, np.zeros(4, np.int64), 0
    occupied = []
    ev0.record()
    for _ in range(steps):
        if flush_buf is not None:
            flush_buf.zero_()          # 256 MB written through the L2 (2 x its capacity): nothing of the last step survives
        if sort:
            res.resort()
        res.step(t_step)
        c = res.g.counters()           # waits for the call; reads the device counters of this step
        pushes += c.n_pushes
        kernel_ms += c.kernel_ms
        fallback += np.array(c.n_fallback)
        n_adaptive += c.n_adaptive
    ev1.record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    clocks = sampler.stop() if sampler else None
    occupied = int(torch.unique(res.ind[res.ind > 0]).numel())
    return dict(pushes=pushes, kernel_ms=kernel_ms, elapsed_ms=ev0.elapsed_time(ev1), fallback=fallback,
                n_adaptive=n_adaptive, launches=launch_count() - launches0, clocks=clocks, occupied_tetra=occupied)


def roofline_of(settings, wl_name, bytes_per_crossing, per_rank_pushes, launch_ms, hbm_peak, peak_src, muladd_peak, dfma_peak,
                kernel_share, has_phi, strong_e, ext, args, gather_mode=0):
    """north_star: the slower of the two per-push limits decides the bound -- HBM gather vs FP64 issue."""
    achieved = bytes_per_crossing * per_rank_pushes / (launch_ms * 1e-3) / 1e9
    kkey = "rk4" if settings.ipusher == 1 else settings.poly_order
    fp64_per = FP64_INST_PER_CROSSING[kkey]
    ncu = ncu_entry(wl_name, kkey) if not has_phi and not strong_e and not ext else _NCU.get(f"{wl_name}:{kkey}")
    traffic = ncu["dram_bytes_per_crossing"] * per_rank_pushes if ncu else None
    traffic_src = (f"ncu capture {ncu['capture']}: {ncu['dram_bytes_per_crossing']:.2f} B/crossing x crossings per launch"
                   if ncu else None)
    # what the DRAM actually delivers: the capture's bytes per crossing at this run's crossing rate, against the copy peak
    dram_frac = (traffic / (launch_ms * 1e-3) / 1e9 / hbm_peak) if traffic is not None else None
    t_hbm, t_fp64 = bytes_per_crossing / (hbm_peak * 1e9), fp64_per / muladd_peak
    fp64_ach = fp64_per * per_rank_pushes / (launch_ms * 1e-3)
    fp64 = {"achieved": fp64_ach / 1e12, "peak": muladd_peak / 1e12, "unit": "Tinst/s (thread-level DMUL/DADD)",
            "frac": fp64_ach / muladd_peak, "inst_per_crossing": fp64_per, "dfma_peak": dfma_peak / 1e12,
            "peak_source": "measured in this run (gorilla_b200_fp64_peak)"}
    phi = 2 if strong_e else 1 if has_phi else 0
    tag = (",EXT=5" if settings.boole_adaptive_time_steps and (args.optional_quantities or ext) else
           ",EXT=2" if args.optional_quantities else ",EXT=1" if ext else ",EXT=3" if settings.boole_adaptive_time_steps else
           ",EXT=4" if settings.ipusher == 2 and settings.i_precomp else
           ",EXT=2" if settings.ipusher == 1 and (settings.boole_newton_precalc or settings.boole_pusher_ode45) else "")
    kern = f"orbit_kernel{'_g' if settings.ipusher == 2 and settings.poly_order >= 3 and 'EXT=5' not in tag else ''}<{0 if settings.ipusher == 1 else settings.poly_order},{phi}{tag}>"
    if gather_mode and not tag and (settings.ipusher == 1 or settings.poly_order == 2):   # the library's launch rule
        kern = kern[:-1] + f",0,GATHER={gather_mode}>"
    common = {"traffic": traffic, "traffic_source": traffic_src, "dram_frac": dram_frac, "kernel": kern,
              "gather": {0: "per-lane vector loads", 1: "per-lane bulk copies (TMA) one push ahead",
                         2: "warp-cooperative cp.async copies one push ahead"}[int(gather_mode)],
              "launch_ms": launch_ms, "kernel_share_of_step": kernel_share,
              "note": "frac = ALGORITHMIC bytes (or FP64 instructions) per second against the peak; dram_frac = DRAM bytes "
                      "actually moved (ncu capture of this kernel on this workload) against the same peak -- a dram_frac far "
                      "below frac means the records the batch touches are L2 resident and the gather is latency / issue bound"}
    if t_hbm >= t_fp64:
        return {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "algorithmic_bytes_per_crossing": bytes_per_crossing, "peak_source": peak_src, "fp64": fp64, **common}
    return {"bound": "fp64", "achieved": fp64["achieved"], "peak": fp64["peak"], "unit": fp64["unit"], "frac": fp64["frac"],
            "peak_source": fp64["peak_source"],
            "hbm": {"achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "algorithmic_bytes_per_crossing": bytes_per_crossing}, **common}


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from gorilla_b200 import Gorilla, build_mesh
    from gorilla_b200.api import comm_unique_id, fp64_peak, shard_range, COMM_ID_BYTES
    import workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = make_workload(args.workload, args.start)
    settings = apply_args(wl["settings"], args)
    t_step = args.t_step or wl["t_step"]

    t0 = time.perf_counter()
    mesh = build_mesh(wl["grid"], settings)        # host, once, replicated on every rank
    t_mesh = time.perf_counter() - t0
    g = Gorilla(mesh, settings)
    g.set_prefetch(args.prefetch)
    g.set_gather(args.gather)
    if args.ctas_per_sm or args.threads:
        g.set_launch_config(args.ctas_per_sm, args.threads)
    if args.no_group:
        g._debug_use_group(0)
    elif args.rebin >= 0:
        g._debug_use_group(2 if args.rebin else 1)
    if world > 1:
        # the library's own communicator: rank 0 creates the NCCL id, torch.distributed only carries the 128 bytes
        idt = torch.zeros(COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        g.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    has_phi = bool(np.any(mesh.tetra_physics[:, 116:125] != 0.0))
    strong_e = bool(settings.boole_strong_electric_field)
    bytes_per_crossing = BYTES_PER_CROSSING[has_phi or strong_e] + (BYTES_STRONG_E if strong_e else 0.0)
    ext = settings.ipusher == 2 and settings.i_time_tracing_option == 2
    if ext or args.optional_quantities:
        bytes_per_crossing += 64.0   # hamiltonian_time record (8 doubles) read at the end of every push

    # particles of this rank.  weak: n per GPU fixed, independent streams per rank; strong (BASELINE config 5): one
    # population of --total-particles, rank r owns the contiguous shard [r N/G, (r+1) N/G)
    if args.scaling == "strong":
        first, n = shard_range(args.total_particles, rank, world)
        # the same population whatever G is: generated in fixed blocks of 1e6 with seeds that depend on the block only
        blk = 1_000_000
        parts = []
        for b in range(first // blk, (first + n + blk - 1) // blk if n else 0):
            xb, vb, wb = wl["particles"](blk, 5000 + b)
            lo, hi = max(first, b * blk) - b * blk, min(first + n, (b + 1) * blk) - b * blk
            parts.append((xb[lo:hi], vb[lo:hi], wb[lo:hi]))
        x, vpar, vperp = (np.concatenate([p[k] for p in parts]) for k in range(3))
    else:
        n = args.particles or wl["n_default"]
        x, vpar, vperp = wl["particles"](n, 1000 + rank)
    res = Resident(g, x, vpar, vperp, dev, args.optional_quantities)
    n_located = int((res.ind > 0).sum())
    flush_buf = None if args.no_l2_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    barrier = dist.barrier if world > 1 else None
    m = timed_steps(res, t_step, args.steps, args.warmup, args.sort, flush_buf, barrier, sampler)

    # ---- the path's only exchange, in the library: counters + conservation diagnostics (one grouped NCCL all-reduce),
    # then the timings (max over ranks) and what the e2e leg adds
    d = res.diag()
    tbuf = torch.tensor([m["elapsed_ms"], m["kernel_ms"], -m["elapsed_ms"]], dtype=torch.float64, device=dev)
    g.comm_allreduce_f64(tbuf, "max", stream=res.stream)
    torch.cuda.synchronize()
    elapsed_ms_max, kernel_ms_max, neg_min = (float(v) for v in tbuf.tolist())
    elapsed_ms_min = -neg_min
    tot_pushes = float(d.n_pushes)
    value = tot_pushes / (elapsed_ms_max * 1e-3)

    # ---- e2e through the host-buffer C ABI (pinned host memory, copies in the timed region)
    e2e = None
    if not args.no_e2e:
        hx = torch.empty((n, 3), dtype=torch.float64).pin_memory()
        hv, hw = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
        hb, hi, hf = (torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(3))
        hnp = torch.empty(n, dtype=torch.int64).pin_memory()
        for h, t in ((hx, res.x), (hv, res.vpar), (hw, res.vperp), (hb, res.init), (hi, res.ind), (hf, res.iface)):
            h.copy_(t)
        torch.cuda.synchronize()
        nx, nv, nw, nb, ni, nf, nnp = (t.numpy() for t in (hx, hv, hw, hb, hi, hf, hnp))
        e_steps = max(1, min(args.steps, 3))
        g.set_host_resort(bool(args.sort))     # the library sorts the uploaded batch itself (caller order is restored)
        g.orbit_timestep_gorilla(nx, nv, nw, t_step, nb, ni, nf, n_pushes=nnp)  # warm
        if barrier:
            barrier()
        t0 = time.perf_counter()
        e_push = 0
        for _ in range(e_steps):
            g.orbit_timestep_gorilla(nx, nv, nw, t_step, nb, ni, nf, n_pushes=nnp)
            e_push += int(nnp.sum())
        e_dt = time.perf_counter() - t0
        ebuf = torch.tensor([float(e_push)], dtype=torch.float64, device=dev)
        etim = torch.tensor([e_dt], dtype=torch.float64, device=dev)
        g.comm_allreduce_f64(ebuf, "sum", stream=res.stream)
        g.comm_allreduce_f64(etim, "max", stream=res.stream)
        torch.cuda.synchronize()
        h2d = n * (3 * 8 + 8 + 8 + 4 + 4 + 4)
        d2h = h2d + n * 8 + n * 8
        e2e = {"value": float(ebuf.item()) / float(etim.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e_steps, "host_resort": bool(args.sort)}

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    dfma_peak, muladd_peak = fp64_peak()

    # ---- variants (rank 0 of a single-GPU run only): short driver-clocked runs of the other kernels on the same mesh
    variants = None
    if world == 1 and not args.no_variants and args.workload in ("auto", "vmec_qi") and args.scaling == "weak" \
            and not (args.poly_order or args.ipusher or args.time_tracing or args.adaptive or args.optional_quantities):
        variants = []
        import dataclasses
        vlist = [("order4", dict(ipusher=2, poly_order=4), 300_000, wl["particles"], wl["name"]),
                 ("order3", dict(ipusher=2, poly_order=3), 300_000, wl["particles"], wl["name"]),
                 ("rk4", dict(ipusher=1), n, wl["particles"], wl["name"]),
                 ("order2_spread_start", dict(ipusher=2, poly_order=2), n, workloads.particles_vmec_alpha_spread,
                  wl["name"] + "_spread")]
        for label, kw, nv_, gen, wname in vlist:
            st = dataclasses.replace(settings, **kw)
            gv = Gorilla(mesh, st)
            gv.set_prefetch(args.prefetch)
            gv.set_gather(args.gather)
            xv, vv, wv = gen(nv_, 1000)
            rv = Resident(gv, xv, vv, wv, dev)
            smp = ClockSampler(sampler.index)
            mv = timed_steps(rv, t_step, 2, 3, args.sort, flush_buf, None, smp)
            dv = rv.diag()
            lm = mv["kernel_ms"] / 2
            rf = roofline_of(st, wname, bytes_per_crossing, mv["pushes"] / 2, lm, hbm_peak, peak_src, muladd_peak, dfma_peak,
                             mv["kernel_ms"] / mv["elapsed_ms"], has_phi, strong_e, False, args, gv.get_gather())
            variants.append({"variant": label, "workload": wname, "particles": nv_, "steps": 2, "warmup": 3,
                             "value": mv["pushes"] / (mv["elapsed_ms"] * 1e-3), "unit": UNIT,
                             "ms_per_step": mv["elapsed_ms"] / 2, "kernel": rf["kernel"], "bound": rf["bound"],
                             "frac": rf["frac"], "hbm_frac": (rf.get("hbm") or rf)["frac"], "fp64_frac": (rf.get("fp64") or rf)["frac"],
                             "dram_frac": rf["dram_frac"], "occupied_record_bytes": mv["occupied_tetra"] * 352,
                             "lost": dv.n_lost, "max_delta_energy": dv.max_delta_energy, "clocks": mv["clocks"]})
            gv.close()
            del rv

    if rank == 0:
        per_rank_pushes = m["pushes"] / max(1, args.steps)
        launch_ms = m["kernel_ms"] / max(1, args.steps)
        roofline = roofline_of(settings, wl["name"], bytes_per_crossing, per_rank_pushes, launch_ms, hbm_peak, peak_src,
                               muladd_peak, dfma_peak, m["kernel_ms"] / m["elapsed_ms"], has_phi, strong_e, ext, args,
                               g.get_gather())
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_s = 2000 * cores
            v, dt, p = cpu_run(wl, mesh, settings, n_s, t_step, max(1, min(args.steps, 3)), 1, cores)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample_particles": n_s,
                            "sample": f"{n_s} particles x {max(1, min(args.steps, 3))} steps of {t_step:g} s "
                                      f"({p} pushes, {dt:.1f} s wall)"}
        hot_rec = 352 + (160 if has_phi or strong_e else 0) + (256 if strong_e else 0)
        cfg = make_config(wl, settings, args, world, n, t_step, mesh)
        l2_bytes = torch.cuda.get_device_properties(dev).L2_cache_size
        footprint = {"mesh_hot_bytes": int(mesh.ntetr * hot_rec), "l2_bytes": int(l2_bytes),
                     "occupied_record_bytes_rank0": int(m["occupied_tetra"] * hot_rec),
                     "occupied_tetrahedra_rank0": int(m["occupied_tetra"]),
                     "note": "records under the particles of rank 0 at the end of the timed region; "
                             + ("smaller than the L2: after the first touches of a step the gather is served from the L2"
                                if m["occupied_tetra"] * hot_rec < 0.8 * l2_bytes else "exceeds the L2: a DRAM gather")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms_max / max(1, args.steps), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "e2e": e2e, "gpu_launches": int(m["launches"]), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "clocks": m["clocks"],
            "counters": {"pushes": tot_pushes, "lost": d.n_lost, "lost_outer": d.n_lost_outer, "lost_inner": d.n_lost_inner,
                         "failed": d.n_failed, "particles": d.n_particles,
                         "fallback_rank0": [int(v) for v in m["fallback"]], "adaptive_pushes_rank0": int(m["n_adaptive"]),
                         "located_rank0": n_located, "find_tetra_ms_rank0": res.find_ms, "mesh_build_s_rank0": t_mesh},
            "diag": {"reduced_by": "gorilla_b200_diag_reduce_dev (device reduction + NCCL all-reduce inside the library)"
                                   if world > 1 else "gorilla_b200_diag_reduce_dev (device reduction; one rank)",
                     "nranks": d.nranks, "n_sampled": d.n_sampled,
                     "max_delta_energy": d.max_delta_energy, "rms_delta_energy": d.rms_delta_energy,
                     "max_delta_perpinv": d.max_delta_perpinv, "rms_delta_perpinv": d.rms_delta_perpinv,
                     "max_delta_p_phi": d.max_delta_p_phi, "rms_delta_p_phi": d.rms_delta_p_phi},
            "footprint": footprint,
            "imbalance": {"rank_time_max_over_min": elapsed_ms_max / elapsed_ms_min if elapsed_ms_min > 0 else None},
            "variants": variants,
        }
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    g.close()
    if world > 1:
        dist.destroy_process_group()


RESULT_OUT = sys.stdout

if __name__ == "__main__":
    RESULT_OUT = _claim_stdout()
    main()
