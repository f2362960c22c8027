"""Regenerates tests/golden/orbits_v1.npz from the CPU oracle (oracle/gorilla_oracle.c).

The reference ships no golden vectors for this path and cannot be compiled in this image (no gfortran), so these
vectors do NOT pin the oracle to the Fortran binary ("parity unpinned").  What they pin is the oracle
itself -- and with it every parity test -- against drift: compiler or libm changes, accidental edits of the restatement,
changes of the mesh builders.  tests/test_golden.py checks the oracle (CPU) and the CUDA path (GPU) against them
bit for bit.  Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import workloads  # noqa: E402
from gorilla_b200 import build_mesh  # noqa: E402
from oracle_binding import OracleMesh  # noqa: E402

N, CAP = 48, 96

CASES = [  # name, settings overrides, t_step, seed
    ("k1", dict(poly_order=1), 1.0e-5, 11),
    ("k2", dict(poly_order=2), 2.0e-5, 12),
    ("k3", dict(poly_order=3), 2.0e-5, 13),
    ("k4", dict(poly_order=4), 2.0e-5, 14),
    ("k4_backward", dict(poly_order=4), -1.0e-5, 15),
    ("k2_noguess", dict(poly_order=2, boole_guess=False), 1.0e-5, 16),
    ("k4_noguess", dict(poly_order=4, boole_guess=False), 1.0e-5, 17),
    ("rk4", dict(ipusher=1), 2.0e-5, 18),
    ("k2_phi", dict(poly_order=2, eps_Phi=-1.0e-7), 2.0e-5, 19),
    ("k4_strongE", dict(poly_order=4, eps_Phi=-1.5e-5, boole_strong_electric_field=True), 2.0e-5, 20),
    ("rk4_strongE", dict(ipusher=1, eps_Phi=-1.5e-5, boole_strong_electric_field=True), 2.0e-5, 21),
]


def run_case(over, t_step, seed):
    grid, st = workloads.analytic_tokamak(10, 10, 10)
    st = type(st)(**{**st.__dict__, **over})
    mesh = build_mesh(grid, st)
    om = OracleMesh(mesh, st)
    x, vpar, vperp = workloads.particles_cyl(N, seed)
    binit, ind, ifc = workloads.fresh_state(N)
    r = om.orbit_timestep_trace(x, vpar, vperp, t_step, binit, ind, ifc, CAP)
    return dict(x=x, vpar=vpar, vperp=vperp, ind_tetr=ind, iface=ifc, trace_tetr=r["trace_tetr"].astype(np.int32),
                trace_face=r["trace_face"].astype(np.int8), n_pushes=r["n_pushes"].astype(np.int64),
                t_remain=r["t_remain"])


def main():
    out = {}
    for name, over, t_step, seed in CASES:
        for k, v in run_case(over, t_step, seed).items():
            out[f"{name}/{k}"] = v
    np.savez_compressed(Path(__file__).with_name("orbits_v1.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
