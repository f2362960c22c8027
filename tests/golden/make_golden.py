"""Regenerates tests/golden/orbits_v1.npz, orbits_v2.npz and orbits_v3.npz from the CPU oracle (oracle/gorilla_oracle.c).

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


# second file (orbits_v2.npz): Hamiltonian time tracing + optional quantities, adaptive sub-stepping, orbit events
_OQ = dict(boole_time_Hamiltonian=True, boole_gyrophase=True, boole_vpar_int=True, boole_vpar2_int=True)
CASES_V2 = [  # name, settings overrides, t_step, seed, kind
    ("k2_hamiltonian_optq", dict(poly_order=2, i_time_tracing_option=2, **_OQ), 2.0e-5, 31, "optq"),
    ("k4_hamiltonian_optq", dict(poly_order=4, i_time_tracing_option=2, **_OQ), 2.0e-5, 32, "optq"),
    ("k3_optq_backward", dict(poly_order=3, boole_vpar_int=True, boole_vpar2_int=True), -1.0e-5, 33, "optq"),
    ("k2_adaptive", dict(poly_order=2, boole_adaptive_time_steps=True, desired_delta_energy=1e-11,
                         max_n_intermediate_steps=30), 2.0e-5, 34, "plain"),
    ("k3_adaptive", dict(poly_order=3, boole_adaptive_time_steps=True, desired_delta_energy=1e-14,
                         max_n_intermediate_steps=30), 2.0e-5, 35, "plain"),
    ("k2_events", dict(poly_order=2), 6.0e-4, 36, "events"),
    ("k4_events", dict(poly_order=4), 6.0e-4, 37, "events"),
]
N_EV, EV_CAP = 24, 20000


def run_case_v2(over, t_step, seed, kind):
    grid, st = workloads.analytic_tokamak(10, 10, 10)
    st = type(st)(**{**st.__dict__, **over})
    mesh = build_mesh(grid, st)
    om = OracleMesh(mesh, st)
    n = N_EV if kind == "events" else N
    x, vpar, vperp = workloads.particles_cyl(n, seed)
    binit, ind, ifc = workloads.fresh_state(n)
    if kind == "events":
        J, cv, cp = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
        ev, nev, npush = om.orbit_timestep_events(x, vpar, vperp, t_step, binit, ind, ifc, J, cv, cp, EV_CAP, n_skip_phi_0=2)
        assert nev <= EV_CAP
        ev = ev[np.lexsort((ev["kind"], ev["push"], ev["particle"]))]
        return dict(x=x, vpar=vpar, vperp=vperp, ind_tetr=ind, iface=ifc, n_pushes=npush, par_adiab_inv=J, counter_vpar_0=cv,
                    counter_phi_0=cp, ev_particle=ev["particle"], ev_kind=ev["kind"], ev_counter=ev["counter"],
                    ev_push=ev["push"], ev_x=ev["x"], ev_value=ev["value"])
    r = om.orbit_timestep_trace(x, vpar, vperp, t_step, binit, ind, ifc, CAP)
    out = dict(x=x, vpar=vpar, vperp=vperp, ind_tetr=ind, iface=ifc, trace_tetr=r["trace_tetr"].astype(np.int32),
               trace_face=r["trace_face"].astype(np.int8), n_pushes=r["n_pushes"].astype(np.int64), t_remain=r["t_remain"])
    if kind == "optq":
        out["optional_quantities"] = r["optional_quantities"]
    return out


# third file (orbits_v3.npz): full-orbit output (event kind 3, with the elapsed time of every event) for both pushers
CASES_V3 = [  # name, settings overrides, t_step, seed, event switches
    ("k2_full_orbit", dict(poly_order=2), 2.0e-4, 41, dict(full_orbit=True, n_skip_full_orbit=3, n_skip_phi_0=2)),
    ("k4_full_orbit_only", dict(poly_order=4), 1.0e-4, 42,
     dict(poincare_phi_0=False, poincare_vpar_0=False, J_par=False, full_orbit=True, n_skip_full_orbit=1)),
    ("rk4_full_orbit", dict(ipusher=1), 2.0e-4, 43, dict(full_orbit=True, n_skip_full_orbit=2)),
]


def run_case_v3(over, t_step, seed, switches):
    grid, st = workloads.analytic_tokamak(10, 10, 10)
    st = type(st)(**{**st.__dict__, **over})
    om = OracleMesh(build_mesh(grid, st), st)
    n = N_EV
    x, vpar, vperp = workloads.particles_cyl(n, seed)
    binit, ind, ifc = workloads.fresh_state(n)
    J, cv, cp = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
    ev, nev, npush = om.orbit_timestep_events(x, vpar, vperp, t_step, binit, ind, ifc, J, cv, cp, EV_CAP, **switches)
    assert nev <= EV_CAP
    ev = ev[np.lexsort((ev["kind"], ev["push"], ev["particle"]))]
    return dict(x=x, vpar=vpar, vperp=vperp, ind_tetr=ind, iface=ifc, n_pushes=npush, par_adiab_inv=J, counter_vpar_0=cv,
                counter_phi_0=cp, ev_particle=ev["particle"], ev_kind=ev["kind"], ev_counter=ev["counter"],
                ev_push=ev["push"], ev_x=ev["x"], ev_value=ev["value"], ev_t=ev["t"])


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
    out = {}
    for name, over, t_step, seed, kind in CASES_V2:
        for k, v in run_case_v2(over, t_step, seed, kind).items():
            out[f"{name}/{k}"] = v
    np.savez_compressed(Path(__file__).with_name("orbits_v2.npz"), **out)
    print("wrote", len(out), "arrays (v2)")
    out = {}
    for name, over, t_step, seed, switches in CASES_V3:
        for k, v in run_case_v3(over, t_step, seed, switches).items():
            out[f"{name}/{k}"] = v
    np.savez_compressed(Path(__file__).with_name("orbits_v3.npz"), **out)
    print("wrote", len(out), "arrays (v3)")


if __name__ == "__main__":
    main()
