"""Physics self-checks of the oracle (the substitutes for the golden vectors the reference does not have,
SURVEY.md 8c): the method conserves the magnetic moment exactly and, for an exact solution of the linearised
equations, total energy and (in axisymmetry) toroidal momentum (README.md:15)."""
import numpy as np

import workloads
from oracle_binding import OracleMesh


def _run(mesh, settings, K, n=100, steps=4, t_step=5e-5, seed=2):
    settings = type(settings)(**{**settings.__dict__, "poly_order": K})
    om = OracleMesh(mesh, settings)
    x, vpar, vperp = workloads.particles_cyl(n, seed)
    binit, ind, ifc = workloads.fresh_state(n)
    om.orbit_timestep_batch(x, vpar, vperp, 0.0, binit, ind, ifc)
    e0, p0, mu0 = om.invariants(x, vpar, vperp, ind)
    total = 0
    for _ in range(steps):
        total += om.orbit_timestep_batch(x, vpar, vperp, t_step, binit, ind, ifc, nthreads=4)
    e1, p1, mu1 = om.invariants(x, vpar, vperp, ind)
    ok = ind > 0
    return (np.abs(e1 / e0 - 1)[ok].max(), np.abs(p1 / p0 - 1)[ok].max(), np.abs(mu1 / mu0 - 1)[ok].max(), total, x,
            vpar, ind)


def test_invariants_order4(small_mesh):
    mesh, _, settings = small_mesh
    dE, dP, dMu, total, *_ = _run(mesh, settings, 4)
    assert total > 10000
    assert dMu < 1e-13   # perpinv is carried unchanged; vperp round-trips through sqrt
    assert dE < 1e-11    # order-4 Taylor solution of the linear ODE: energy drift at round-off level
    assert dP < 1e-9


def test_energy_error_shrinks_with_order(small_mesh):
    mesh, _, settings = small_mesh
    dE2 = _run(mesh, settings, 2)[0]
    dE3 = _run(mesh, settings, 3)[0]
    dE4 = _run(mesh, settings, 4)[0]
    assert dE2 > dE3 > dE4 or (dE3 < 1e-9 and dE4 < 1e-11)
    assert dE2 < 1e-4


def test_orders_converge_to_the_same_orbit(small_mesh):
    mesh, _, settings = small_mesh
    *_, x3, v3, i3 = _run(mesh, settings, 3, n=50, steps=1, t_step=2e-5)
    *_, x4, v4, i4 = _run(mesh, settings, 4, n=50, steps=1, t_step=2e-5)
    ok = (i3 > 0) & (i4 > 0)
    assert np.abs(x3[ok] - x4[ok]).max() < 1e-3
    assert np.abs(v3[ok] / v4[ok] - 1).max() < 1e-5
