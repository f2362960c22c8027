#!/usr/bin/env python
"""BASELINE config 3 end to end: 10^6 alphas of 3.5 MeV on the QI stellarator mesh, 100 steps of 1e-4 s (1e-2 s total).
Prints the alpha loss fraction with its binomial error, the throughput of the whole run, and checks a sub-sample of the
same particles against the CPU restatement (identical lost set, identical final state).
Usage: tools/config3_loss.py [particles [cpu_subsample [i_time_tracing_option]]]"""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    n_sub = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
    nsteps, t_step = 100, 1.0e-4
    grid, st = workloads.vmec_qi(str(ROOT / "data" / "equilibria" / "netcdf_file_for_test.nc"))
    st.i_time_tracing_option = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    mesh = build_mesh(grid, st)
    g = Gorilla(mesh, st)
    dev = torch.device("cuda", 0)
    x, vpar, vperp = workloads.particles_vmec_alpha(n, 1000)
    binit, ind, ifc = workloads.fresh_state(n)
    tt = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    xd, vd, wd, bd, it, fd = tt(x), tt(vpar), tt(vperp), tt(binit), tt(ind), tt(ifc)
    order = torch.arange(n, device=dev)
    g.orbit_timestep_gorilla_dev(xd, vd, wd, 0.0, bd, it, fd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pushes, kernel_ms, lost_curve = 0, 0.0, []
    for s in range(nsteps):
        perm = torch.empty(n, dtype=torch.int64, device=dev)
        g.sort_permutation_dev(it, perm)
        xd, vd, wd, bd, it, fd, order = (a[perm].contiguous() for a in (xd, vd, wd, bd, it, fd, order))
        g.orbit_timestep_gorilla_dev(xd, vd, wd, t_step, bd, it, fd)
        c = g.counters()
        pushes += c.n_pushes
        kernel_ms += c.kernel_ms
        lost_curve.append(int((it < 1).sum()))
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    inv = torch.empty_like(order)
    inv[order] = torch.arange(n, device=dev)
    xg, vg, wg, ig = (a[inv].cpu().numpy() for a in (xd, vd, wd, it))
    p = lost_curve[-1] / n
    out = {"i_time_tracing_option": st.i_time_tracing_option, "particles": n, "steps": nsteps, "t_step_s": t_step, "crossings": pushes, "wall_s": wall, "kernel_s": kernel_ms * 1e-3,
           "crossings_per_s_wall": pushes / wall, "crossings_per_s_kernel": pushes / (kernel_ms * 1e-3),
           "loss_fraction": p, "loss_fraction_sigma": float(np.sqrt(p * (1 - p) / n)),
           "lost_after_step": lost_curve[9::10]}
    # the same first n_sub particles through the CPU restatement
    xs, vs, ws = x[:n_sub].copy(), vpar[:n_sub].copy(), vperp[:n_sub].copy()
    bs, is_, fs = workloads.fresh_state(n_sub)
    om = OracleMesh(mesh, st)
    t1 = time.perf_counter()
    cpu_push = 0
    for s in range(nsteps):
        cpu_push += om.orbit_timestep_batch(xs, vs, ws, t_step, bs, is_, fs)
    cpu_wall = time.perf_counter() - t1
    same_lost = bool(np.array_equal(is_ < 1, ig[:n_sub] < 1))
    alive = is_ > 0
    same_state = bool(np.array_equal(xs[alive], xg[:n_sub][alive]) and np.array_equal(vs[alive], vg[:n_sub][alive])
                      and np.array_equal(ws[alive], wg[:n_sub][alive]) and np.array_equal(is_, ig[:n_sub]))
    ps = float((is_ < 1).mean())
    out["cpu_subsample"] = {"particles": n_sub, "crossings": int(cpu_push), "wall_s": cpu_wall,
                            "crossings_per_s": cpu_push / cpu_wall, "loss_fraction": ps,
                            "loss_fraction_sigma": float(np.sqrt(ps * (1 - ps) / n_sub)),
                            "same_lost_set_as_gpu": same_lost, "final_state_bit_identical_to_gpu": same_state}
    print(json.dumps(out))
    g.close()
    assert same_lost and same_state


if __name__ == "__main__":
    main()
