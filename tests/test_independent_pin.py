"""A pin that does NOT share the pusher restatement's reading of the Fortran.

tests/independent_orbit.py follows a particle by integrating dz/dtau = b + A z per tetrahedron with scipy's DOP853 and the
exact matrix-exponential flow (b, A from the record as SURVEY.md Appendix A states them; hand-over and time accounting in
numpy): no Taylor series in tau, no polynomial root solver, no fall-back ladder.  The order-4 pusher must visit the same
tetrahedra through the same faces and end the time step at the same point to 1e-10 relative (north_star's bound); orders 3
and 2 must converge towards it.  What this does and does not pin: it pins analytic_coeff / analytic_approx / the root-solver
chain / analytic_integration / t_pass / the stop-inside branch / pusher_handover2neighbour against an independent solution
of the SAME piecewise-linear equations; it cannot pin details in which the Fortran may differ from those equations, which
only a dump of the gfortran build can (tests/test_reference_dump.py)."""
from pathlib import Path

import numpy as np
import pytest

import workloads
from gorilla_b200 import build_mesh
from independent_orbit import independent_orbit
from oracle_binding import OracleMesh

ROOT = Path(__file__).resolve().parent.parent
TOL = 1.0e-10   # north_star: positions and velocities agree to 1e-10 relative


def _with(settings, **kw):
    return type(settings)(**{**settings.__dict__, **kw})


def _pusher_run(kind, mesh, settings, x, vpar, vperp, t_step, cap):
    """(trace, n_pushes, fallback-free mask is not available per particle -> total fallbacks) from the oracle or the CUDA path"""
    n = x.shape[0]
    b, i, f = workloads.fresh_state(n)
    if kind == "oracle":
        om = OracleMesh(mesh, settings)
        om.orbit_timestep_batch(x, vpar, vperp, 0.0, b, i, f, nthreads=1)
        start = i.copy()
        r = om.orbit_timestep_trace(x, vpar, vperp, t_step, b, i, f, cap)
        return start, r["trace_tetr"], r["trace_face"], r["n_pushes"], i
    from gorilla_b200 import Gorilla
    g = Gorilla(mesh, settings)
    g.orbit_timestep_gorilla(x, vpar, vperp, 0.0, b, i, f)
    start = i.copy()
    npu = np.zeros(n, np.int64)
    tt, tf = g.orbit_timestep_gorilla(x, vpar, vperp, t_step, b, i, f, n_pushes=npu, trace_cap=cap)
    g.close()
    return start, tt, tf, npu, i


def _compare(kind, mesh, settings, particles, t_step, cap=512, min_ok=0.8):
    x, vpar, vperp = particles
    x0, v0, w0 = x.copy(), vpar.copy(), vperp.copy()
    start, tt, tf, npu, ind_end = _pusher_run(kind, mesh, settings, x, vpar, vperp, t_step, cap)
    n = x.shape[0]
    scale = np.abs(x0).max(axis=0)            # per-coordinate size of the domain
    ok = 0
    err_x = err_v = 0.0
    for p in range(n):
        assert start[p] > 0
        o = independent_orbit(mesh, x0[p], v0[p], w0[p], int(start[p]), t_step, max_crossings=cap)
        k = int(npu[p])
        seq = list(zip(tt[p][:k].tolist(), tf[p][:k].tolist()))
        if o["margin"] < 1e-7 or k > cap:
            continue                          # an exit point within 1e-7 of an edge: the face choice is legitimately marginal
        assert seq == o["seq"], f"particle {p}: visited (tetrahedron, face) sequence differs from the independent integrator"
        assert int(ind_end[p]) == o["ind_tetr"]
        ex = float(np.max(np.abs(np.asarray(o["x"]) - x[p]) / scale))
        ev = max(abs(o["vpar"] - vpar[p]), abs(o["vperp"] - vperp[p])) / np.hypot(v0[p], w0[p])
        err_x, err_v = max(err_x, ex), max(err_v, ev)
        ok += 1
    assert ok >= min_ok * n, f"only {ok} of {n} particles had no marginal crossing"
    return err_x, err_v


def _cases(kind, product_lib):
    grid, settings = workloads.analytic_tokamak(20, 20, 20)
    mesh = build_mesh(grid, settings)
    out = {}
    for K in (2, 3, 4):
        out[K] = _compare(kind, mesh, _with(settings, poly_order=K), workloads.particles_cyl(24, 5), 1.0e-5)
    return out


def _check_orders(err):
    (x2, v2), (x3, v3), (x4, v4) = err[2], err[3], err[4]
    assert x4 <= TOL and v4 <= TOL, f"order 4 vs independent integrator: {x4:.2e} {v4:.2e}"
    assert x3 <= 1e-7 and v3 <= 1e-7
    assert x4 < x3 < x2 < 1e-3 and v2 < 1e-3, "orders 2 -> 3 -> 4 must converge to the independent solution"


def test_oracle_order4_matches_independent_integrator(product_lib, oracle_lib):
    _check_orders(_cases("oracle", product_lib))


def test_oracle_backward_time_and_potential(product_lib, oracle_lib):
    grid, settings = workloads.analytic_tokamak(16, 16, 16)
    settings.eps_Phi = -1.0e-7
    settings.poly_order = 4
    mesh = build_mesh(grid, settings)
    assert np.any(mesh.tetra_physics[:, 116:125] != 0.0)        # betmat live
    for t_step in (1.0e-5, -1.0e-5):
        ex, ev = _compare("oracle", mesh, settings, workloads.particles_cyl(16, 7), t_step)
        assert ex <= TOL and ev <= TOL, (t_step, ex, ev)


def test_oracle_flux_coordinates_periodic_handover(product_lib, oracle_lib):
    """VMEC mesh in symmetry-flux coordinates (sign_sqg = -1, theta and phi periodic boundaries), 3.5 MeV alphas."""
    grid, settings = workloads.vmec_qi(ROOT / "data" / "equilibria" / "netcdf_file_for_test.nc", n1=50, n2=20, n3=30,
                                       poly_order=4)
    mesh = build_mesh(grid, settings)
    ex, ev = _compare("oracle", mesh, settings, workloads.particles_vmec_alpha(16, 3), 1.5e-6, min_ok=0.7)
    assert ex <= TOL and ev <= TOL, (ex, ev)


def test_oracle_rk4_converges_to_the_independent_solution(product_lib, oracle_lib):
    """The RK4 pusher (ipusher = 1) integrates the same equations numerically: same sequence, positions to RK4 accuracy."""
    grid, settings = workloads.analytic_tokamak(20, 20, 20)
    mesh = build_mesh(grid, settings)
    ex, ev = _compare("oracle", mesh, _with(settings, ipusher=1), workloads.particles_cyl(16, 5), 1.0e-5)
    assert ex <= 1e-8 and ev <= 1e-8, (ex, ev)


@pytest.mark.gpu
def test_cuda_order4_matches_independent_integrator(product_lib, cuda_device):
    _check_orders(_cases("cuda", product_lib))


@pytest.mark.gpu
def test_cuda_flux_coordinates_and_rk4(product_lib, cuda_device):
    grid, settings = workloads.vmec_qi(ROOT / "data" / "equilibria" / "netcdf_file_for_test.nc", n1=50, n2=20, n3=30,
                                       poly_order=4)
    mesh = build_mesh(grid, settings)
    ex, ev = _compare("cuda", mesh, settings, workloads.particles_vmec_alpha(16, 3), 1.5e-6, min_ok=0.7)
    assert ex <= TOL and ev <= TOL, (ex, ev)
    grid, settings = workloads.analytic_tokamak(20, 20, 20)
    mesh = build_mesh(grid, settings)
    ex, ev = _compare("cuda", mesh, _with(settings, ipusher=1), workloads.particles_cyl(16, 5), 1.0e-5)
    assert ex <= 1e-8 and ev <= 1e-8, (ex, ev)
