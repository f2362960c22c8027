"""Seeded synthetic workloads shared by tests, smoke() and bench.py (no file under /root/reference is read)."""
from __future__ import annotations

import numpy as np

from gorilla_b200 import GorillaSettings, TetraGridSettings

EV2ERG = 1.6022e-12  # constants_mod.f90
AMP = 1.6726e-24


def analytic_tokamak(n1=40, n2=80, n3=40):
    """EXAMPLES/example_8 (tetra_grid.inp / gorilla.inp): analytic circular tokamak, grid_kind = 5."""
    grid = TetraGridSettings(grid_kind=5, n1=n1, n2=n2, n3=n3, boole_n_field_periods=True,
                             R0_analytic_circ=170.0, a_analytic_circ=50.0, B0_analytic_circ=20000.0,
                             q0_analytic_circ=1.1, q1_analytic_circ=2.0)
    settings = GorillaSettings(eps_Phi=0.0, coord_system=1, ispecies=2, boole_periodic_relocation=False, ipusher=2,
                               poly_order=4, boole_guess=True)
    return grid, settings


def particles_cyl(n, seed, energy_ev=3.0e3, mass=2.0 * AMP, R0=170.0, a=50.0, rmin_frac=0.1, rmax_frac=0.85):
    """Start positions uniform in minor radius / poloidal angle / toroidal angle, pitch uniform in [-1,1]
    (gorilla_plot_mod.f90:198,210-211 for vmod/vpar/vperp)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rho = a * (rmin_frac + (rmax_frac - rmin_frac) * rng.random(n))
    th = 2 * np.pi * rng.random(n)
    x = np.empty((n, 3))
    x[:, 0] = R0 + rho * np.cos(th)
    x[:, 1] = 2 * np.pi * rng.random(n)
    x[:, 2] = rho * np.sin(th)
    lam = 2.0 * rng.random(n) - 1.0
    vmod = np.sqrt(2.0 * energy_ev * EV2ERG / mass)
    vpar = lam * vmod
    vperp = np.sqrt(vmod ** 2 - vpar ** 2)
    return x, vpar, vperp


def fresh_state(n):
    return np.zeros(n, np.int32), np.full(n, -1, np.int32), np.full(n, -1, np.int32)


def vmec_qi(netcdf_path, n1=100, n2=40, n3=40, poly_order=2):
    """BASELINE config 3: QI stellarator netcdf_file_for_test.nc (VMEC), field-aligned grid 100x40x40
    (INPUT/tetra_grid.inp defaults), 3.5 MeV alphas (ispecies = 3), symmetry-flux coordinates."""
    grid = TetraGridSettings(grid_kind=3, n1=n1, n2=n2, n3=n3, boole_n_field_periods=True, sfc_s_min=0.1,
                             i_radial_spacing=1, netcdf_filename=str(netcdf_path))
    settings = GorillaSettings(eps_Phi=0.0, coord_system=2, ispecies=3, boole_periodic_relocation=True, ipusher=2,
                               poly_order=poly_order, boole_guess=True)
    return grid, settings


def particles_vmec_alpha(n, seed, energy_ev=3.5e6, s0=0.5, nfp=5):
    """Alpha particles started on the flux surface s0: theta, phi uniform, pitch uniform in [-1, 1]
    (vmod = sqrt(2 E / m), gorilla_plot_mod.f90:198)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.empty((n, 3))
    x[:, 0] = s0
    x[:, 1] = 2 * np.pi * rng.random(n)
    x[:, 2] = 2 * np.pi / nfp * rng.random(n)
    lam = 2.0 * rng.random(n) - 1.0
    vmod = np.sqrt(2.0 * energy_ev * EV2ERG / (4.0 * AMP))
    vpar = lam * vmod
    vperp = np.sqrt(vmod ** 2 - vpar ** 2)
    return x, vpar, vperp


def particles_vmec_alpha_spread(n, seed, s_lo=0.15, s_hi=0.95, **kw):
    """As particles_vmec_alpha, started uniformly in s over [s_lo, s_hi] instead of on one flux surface: the batch then
    touches (nearly) every record of the mesh in every step, which does not fit the L2."""
    x, vpar, vperp = particles_vmec_alpha(n, seed, **kw)
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    x[:, 0] = s_lo + (s_hi - s_lo) * rng.random(n)
    return x, vpar, vperp


def west_soledge3x(data_dir, n2=60, strong=True, ipusher=1, poly_order=2):
    """BASELINE config 4: WEST equilibrium table + SOLEDGE3X-EIRENE triangle mesh extruded to n2 toroidal
    slices (4.24 M tetrahedra at n2 = 60), strong-electric-field mode with eps_Phi = -1.5e-5 (MATLAB/example_8.m:42-48),
    RK4 pusher, cylindrical coordinates, W74+ (ispecies = 4: the heavy-impurity case the strong-field terms exist for;
    600 keV deuterons have banana widths of the minor radius in this equilibrium and are lost promptly)."""
    d = str(data_dir)
    grid = TetraGridSettings(grid_kind=4, n1=100, n2=n2, n3=60, boole_n_field_periods=True,
                             g_file_filename=d + "/g_file_for_test_WEST",
                             convex_wall_filename=d + "/convex_wall_for_test_WEST.dat",
                             knots_SOLEDGE3X_EIRENE_filename=d + "/MESH_SOLEDGE3X_EIRENE/knots_for_test.dat",
                             triangles_SOLEDGE3X_EIRENE_filename=d + "/MESH_SOLEDGE3X_EIRENE/triangles_for_test.dat")
    settings = GorillaSettings(eps_Phi=-1.5e-5 if strong else 0.0, coord_system=1, ispecies=4, boole_periodic_relocation=True,
                               ipusher=ipusher, poly_order=poly_order, boole_guess=True,
                               boole_strong_electric_field=bool(strong))
    return grid, settings


def particles_on_triangles(data_dir, n, seed, energy_ev=6.0e5, mass=184.0 * AMP):
    """Start points uniform over the poloidal mesh area (triangle picked with probability ~ area, uniform barycentric
    point inside), toroidal angle and pitch uniform."""
    d = str(data_dir)
    knots = np.loadtxt(d + "/MESH_SOLEDGE3X_EIRENE/knots_for_test.dat", skiprows=1)
    tri = np.loadtxt(d + "/MESH_SOLEDGE3X_EIRENE/triangles_for_test.dat", skiprows=1, dtype=np.int64) - 1
    a, b, c = knots[tri[:, 0]], knots[tri[:, 1]], knots[tri[:, 2]]
    area = 0.5 * np.abs((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))
    rng = np.random.Generator(np.random.PCG64(seed))
    t = rng.choice(len(tri), size=n, p=area / area.sum())
    u, v = rng.random(n), rng.random(n)
    flip = u + v > 1
    u[flip], v[flip] = 1 - u[flip], 1 - v[flip]
    # keep a little away from the triangle edges so that no start point sits on a cell face
    u, v = 0.02 + 0.94 * u, 0.02 + 0.94 * v * (1 - 0.0)
    p = a[t] + u[:, None] * (b[t] - a[t]) + np.minimum(v, 0.98 - u)[:, None] * (c[t] - a[t])
    x = np.empty((n, 3))
    x[:, 0], x[:, 2] = p[:, 0], p[:, 1]
    x[:, 1] = 2 * np.pi * rng.random(n)
    lam = 2.0 * rng.random(n) - 1.0
    vmod = np.sqrt(2.0 * energy_ev * EV2ERG / mass)
    vpar = lam * vmod
    vperp = np.sqrt(vmod ** 2 - vpar ** 2)
    return x, vpar, vperp


def efit_flux(data_dir, n1=100, n2=40, n3=40, poly_order=2):
    """BASELINE configs 1/2: ASDEX Upgrade g_file_for_test, field-aligned grid in symmetry flux coordinates
    (grid_kind = 2, coord_system = 2), 100x40x40 = 960 000 tetrahedra, deuterons."""
    d = str(data_dir)
    grid = TetraGridSettings(grid_kind=2, n1=n1, n2=n2, n3=n3, boole_n_field_periods=True, sfc_s_min=0.1,
                             g_file_filename=d + "/g_file_for_test", convex_wall_filename=d + "/convex_wall_for_test.dat")
    settings = GorillaSettings(eps_Phi=0.0, coord_system=2, ispecies=2, boole_periodic_relocation=True, ipusher=2,
                               poly_order=poly_order, boole_guess=True)
    return grid, settings


def particles_flux(n, seed, energy_ev=3.0e3, mass=2.0 * AMP, s_lo=0.2, s_hi=0.9, nfp=1):
    """s uniform in [s_lo, s_hi], theta and phi uniform, pitch uniform in [-1, 1]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.empty((n, 3))
    x[:, 0] = s_lo + (s_hi - s_lo) * rng.random(n)
    x[:, 1] = 2 * np.pi * rng.random(n)
    x[:, 2] = 2 * np.pi / nfp * rng.random(n)
    lam = 2.0 * rng.random(n) - 1.0
    vmod = np.sqrt(2.0 * energy_ev * EV2ERG / mass)
    vpar = lam * vmod
    vperp = np.sqrt(vmod ** 2 - vpar ** 2)
    return x, vpar, vperp
