#!/usr/bin/env python
"""BASELINE config 2 end to end: ASDEX Upgrade g-file, symmetry flux coordinates 100x40x40, N = 1e5 D+ of 3 keV,
poly_order = 4, 100 steps of 1e-4 s; after every step the max and rms of |E/E0 - 1|, |mu/mu0 - 1|, |p_phi/p_phi0 - 1| over the
confined particles (gorilla_b200_diag_reduce_dev: device reduction behind the C ABI) and the loss counters.  A sub-sample of
the same particles goes through the CPU restatement for all steps (identical final state).
Usage: tools/config2_conservation.py [particles [cpu_subsample [poly_order]]]"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import workloads  # noqa: E402
from gorilla_b200 import Gorilla, build_mesh  # noqa: E402
from oracle_binding import OracleMesh  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    n_sub = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    nsteps, t_step = 100, 1.0e-4
    grid, st = workloads.efit_flux(ROOT / "data" / "equilibria")
    st.poly_order = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    mesh = build_mesh(grid, st)
    g = Gorilla(mesh, st)
    dev = torch.device("cuda", 0)
    x, vpar, vperp = workloads.particles_flux(n, 2024)
    binit, ind, ifc = workloads.fresh_state(n)
    tt = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    xd, vd, wd, bd, it, fd = tt(x), tt(vpar), tt(vperp), tt(binit), tt(ind), tt(ifc)
    order = torch.arange(n, dtype=torch.int64, device=dev)
    g.orbit_timestep_gorilla_dev(xd, vd, wd, 0.0, bd, it, fd)
    e0, p0, m0 = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
    g.invariants_dev(xd, vd, wd, it, e0, p0, m0)
    g.diag_reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    curve, kernel_ms = [], 0.0
    for s in range(nsteps):
        perm = torch.empty(n, dtype=torch.int64, device=dev)
        g.resort_dev(xd, vd, wd, bd, it, fd, extra=[e0, p0, m0], perm_out=perm)   # the start values travel with their particles
        order = order[perm]
        g.orbit_timestep_gorilla_dev(xd, vd, wd, t_step, bd, it, fd)
        kernel_ms += g.counters().kernel_ms
        d = g.diag_reduce_dev(xd, vd, wd, it, e0, p0, m0)
        if (s + 1) % 10 == 0:
            curve.append({"step": s + 1, "confined": d.n_sampled, "max_dE": d.max_delta_energy, "rms_dE": d.rms_delta_energy,
                          "max_dmu": d.max_delta_perpinv, "max_dpphi": d.max_delta_p_phi, "rms_dpphi": d.rms_delta_p_phi})
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    inv = torch.empty_like(order)
    inv[order] = torch.arange(n, dtype=torch.int64, device=dev)
    xg, vg, wg, ig = (a[inv].cpu().numpy() for a in (xd, vd, wd, it))
    out = {"config": "BASELINE 2: efit symmetry flux 100x40x40, D+ 3 keV", "poly_order": st.poly_order, "particles": n, "steps": nsteps,
           "t_step_s": t_step, "crossings": d.n_pushes, "wall_s": wall, "kernel_s": kernel_ms * 1e-3,
           "crossings_per_s_wall": d.n_pushes / wall, "crossings_per_s_kernel": d.n_pushes / (kernel_ms * 1e-3),
           "lost": d.n_lost, "lost_outer": d.n_lost_outer, "lost_inner": d.n_lost_inner, "failed": d.n_failed,
           "conservation_every_10_steps": curve}
    xs, vs, ws = x[:n_sub].copy(), vpar[:n_sub].copy(), vperp[:n_sub].copy()
    bs, is_, fs = workloads.fresh_state(n_sub)
    om = OracleMesh(mesh, st)
    t1 = time.perf_counter()
    cpu_push = 0
    for s in range(nsteps):
        cpu_push += om.orbit_timestep_batch(xs, vs, ws, t_step, bs, is_, fs)
    cpu_wall = time.perf_counter() - t1
    alive = is_ > 0
    same_state = bool(np.array_equal(is_, ig[:n_sub]) and np.array_equal(xs[alive], xg[:n_sub][alive]) and
                      np.array_equal(vs[alive], vg[:n_sub][alive]) and np.array_equal(ws[alive], wg[:n_sub][alive]))
    out["cpu_subsample"] = {"particles": n_sub, "crossings": int(cpu_push), "wall_s": cpu_wall, "crossings_per_s": cpu_push / cpu_wall,
                            "final_state_bit_identical_to_gpu": same_state}
    print(json.dumps(out))
    g.close()
    assert same_state


if __name__ == "__main__":
    main()
