"""Throughput of the event-capturing kernels (EXT = 2) with the full-orbit output switched on: BASELINE config 3 mesh,
order 2, host-pointer entry point; kernel time from the library's own CUDA events (gorilla_counters.kernel_ms).
    python tools/full_orbit_rate.py [n_particles] [n_skip_full_orbit]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import workloads  # noqa: E402
from gorilla_b200 import Gorilla, api, build_mesh  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
nskip = int(sys.argv[2]) if len(sys.argv) > 2 else 50
grid, settings = workloads.vmec_qi(str(ROOT / "data" / "equilibria" / "netcdf_file_for_test.nc"))
mesh = build_mesh(grid, settings)
g = Gorilla(mesh, settings)
x, vpar, vperp = workloads.particles_vmec_alpha(n, 1)
state = workloads.fresh_state(n)
g.orbit_timestep_gorilla(x, vpar, vperp, 0.0, *state)            # localisation
J, cv, cp = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
out = {}
for label, kw in (("events_off_plain_kernel", None),
                  ("phi0_vpar0_Jpar", dict()),
                  ("full_orbit_only", dict(boole_poincare_phi_0=False, boole_poincare_vpar_0=False, boole_J_par=False,
                                           boole_full_orbit=True, n_skip_full_orbit=nskip)),
                  ("all_kinds", dict(boole_full_orbit=True, n_skip_full_orbit=nskip))):
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        if kw is None:
            g.orbit_timestep_gorilla(x, vpar, vperp, 2e-5, *state)
            nev = 0
        else:
            ev, nev = g.orbit_timestep_gorilla_events(x, vpar, vperp, 2e-5, *state, J, cv, cp, 8_000_000, **kw)
        wall = time.perf_counter() - t0
        c = g.counters()
        rate = c.n_pushes / (c.kernel_ms * 1e-3)
        if best is None or rate > best["crossings_per_s_kernel"]:
            best = dict(crossings_per_s_kernel=rate, kernel_ms=c.kernel_ms, n_pushes=c.n_pushes, n_events=int(nev),
                        wall_s=wall)
    out[label] = best
print(json.dumps(dict(n_particles=n, n_skip_full_orbit=nskip, t_step=2e-5, workload="vmec_qi_alpha_3.5MeV_100x40x40, order 2",
                      results=out)))
g.close()
