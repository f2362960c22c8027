"""The device algorithm (gb_poly.cuh / gb_find.cuh compiled for the host, tests/host_mirror) against the
oracle on identical seeded inputs: visited-tetra sequence, positions, velocities and remaining time must be
IDENTICAL (same IEEE operations in the same order), for every polynomial order, through both the fast path and
the complete fall-back ladder.  The CUDA build of the same headers is checked by tests/test_gpu_parity.py."""
import numpy as np
import pytest

import workloads
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def run_pair(mesh, settings, n, seed, t_step, cap, force_full=False, nsteps=1, **pk):
    om, hm = OracleMesh(mesh, settings), HostMirror(mesh, settings)
    xa, va, wa = workloads.particles_cyl(n, seed, **pk)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    for _ in range(nsteps):
        ra = om.orbit_timestep_trace(xa, va, wa, t_step, ia, ta, fa, cap)
        rb = hm.orbit_timestep(xb, vb, wb, t_step, ib, tb, fb, cap, force_full=force_full)
        assert same(ra["trace_tetr"], rb["trace_tetr"]) and same(ra["trace_face"], rb["trace_face"])
        assert same(ra["n_pushes"], rb["n_pushes"])
        assert same(xa, xb) and same(va, vb) and same(wa, wb)
        assert same(ta, tb) and same(fa, fb) and same(ia, ib)
        assert same(ra["t_remain"], rb["t_remain"])
        assert same(ra["fallback"], rb["fallback"])
    return ra, ta


@pytest.mark.parametrize("K", [1, 2, 3, 4])
@pytest.mark.parametrize("force_full", [False, True])
def test_trace_and_state_identical(small_mesh, K, force_full):
    mesh, _, settings = small_mesh
    settings = type(settings)(**{**settings.__dict__, "poly_order": K})
    ra, ind = run_pair(mesh, settings, 200, 3, 2e-5, 256, force_full=force_full)
    assert ra["n_pushes"].sum() > 5000


@pytest.mark.parametrize("K", [2, 4])
def test_with_electrostatic_potential(small_mesh_phi, K):
    mesh, _, settings = small_mesh_phi
    assert np.any(mesh.tetra_physics[:, 116:125] != 0.0)  # betmat live -> PHI kernel variant
    settings = type(settings)(**{**settings.__dict__, "poly_order": K})
    run_pair(mesh, settings, 150, 5, 2e-5, 200)


@pytest.mark.parametrize("K", [2, 3, 4])
def test_backward_in_time_and_multiple_steps(small_mesh, K):
    mesh, _, settings = small_mesh
    settings = type(settings)(**{**settings.__dict__, "poly_order": K})
    run_pair(mesh, settings, 100, 8, -1e-5, 128, nsteps=3)


@pytest.mark.parametrize("K", [2, 4])
def test_losses_through_the_domain_boundary(small_mesh, K):
    """Starts close to the edge of the rectangular grid: many particles leave (ind_tetr = -1)."""
    mesh, _, settings = small_mesh
    settings = type(settings)(**{**settings.__dict__, "poly_order": K})
    ra, ind = run_pair(mesh, settings, 200, 13, 2e-4, 64, rmin_frac=0.9, rmax_frac=0.99, energy_ev=3.0e4)
    assert (ind == -1).sum() > 10 and (ind > 0).sum() > 10


def test_boole_guess_false(small_mesh):
    mesh, _, settings = small_mesh
    settings = type(settings)(**{**settings.__dict__, "poly_order": 4, "boole_guess": False})
    run_pair(mesh, settings, 100, 21, 1e-5, 128)


def test_zero_time_step_only_localises(small_mesh):
    mesh, _, settings = small_mesh
    ra, ind = run_pair(mesh, settings, 64, 4, 0.0, 4)
    assert ra["n_pushes"].sum() == 0 and np.all(ind > 0)


def test_start_on_cell_faces(small_mesh):
    """Start points exactly on grid planes (R, Z and phi planes of the rectangular mesh) exercise find_tetra's
    face-degenerate branch (find_tetra_mod.f90:478-581)."""
    mesh, grid, settings = small_mesh
    settings = type(settings)(**{**settings.__dict__, "poly_order": 2})
    om, hm = OracleMesh(mesh, settings), HostMirror(mesh, settings)
    n = 120
    xa, va, wa = workloads.particles_cyl(n, 17, rmin_frac=0.2, rmax_frac=0.6)
    hr, hz, hphi = 100.0 / grid.n1, 100.0 / grid.n3, 2 * np.pi / grid.n2
    xa[0::3, 0] = 120.0 + hr * np.round((xa[0::3, 0] - 120.0) / hr)
    xa[1::3, 2] = -50.0 + hz * np.round((xa[1::3, 2] + 50.0) / hz)
    xa[2::3, 1] = hphi * np.floor(xa[2::3, 1] / hphi)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    ra = om.orbit_timestep_trace(xa, va, wa, 1e-5, ia, ta, fa, 64)
    rb = hm.orbit_timestep(xb, vb, wb, 1e-5, ib, tb, fb, 64)
    assert same(ra["trace_tetr"], rb["trace_tetr"]) and same(xa, xb) and same(va, vb) and same(ta, tb)
    assert (fa != 0).sum() + (ta > 0).sum() > 0
