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
