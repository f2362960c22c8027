"""Options of the host mesh builder that change what the records hold, not how they are pushed:
the analytical helical perturbation of gorilla.inp (boole_helical_pert, vector_potential_sthetaphi,
SRC/tetra_physics_mod.f90:1158-1161) and the optional bmod_multiplier argument of initialize_gorilla
(SRC/orbit_timestep_gorilla.f90:151,258-262; SRC/tetra_physics_mod.f90:281-286,1095,1145,1201)."""
import dataclasses
from pathlib import Path

import numpy as np
import pytest

import workloads
from gorilla_b200 import GorillaSettings, TetraGridSettings, build_mesh, load_gorilla_inp
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh

DATA = Path(__file__).resolve().parent.parent / "data" / "equilibria"
BMOD1, ATHETA1, APHI1, H1 = 24, 25, 26, 27
GB, GAPHI, CURLA = 59, 77, 21


def flux_grid(**kw):
    return TetraGridSettings(grid_kind=2, n1=16, n2=8, n3=16, boole_n_field_periods=True, sfc_s_min=0.1,
                             g_file_filename=str(DATA / "g_file_for_test"),
                             convex_wall_filename=str(DATA / "convex_wall_for_test.dat"), **kw)


def flux_settings(**kw):
    return GorillaSettings(eps_Phi=0.0, coord_system=2, ispecies=2, boole_periodic_relocation=True, ipusher=2,
                           poly_order=2, boole_guess=True, **kw)


@pytest.fixture(scope="module")
def plain(product_lib):
    return build_mesh(flux_grid(), flux_settings())


@pytest.fixture(scope="module")
def helical(product_lib):
    st = flux_settings(boole_helical_pert=True, helical_pert_eps_Aphi=0.05, helical_pert_m_fourier=2,
                       helical_pert_n_fourier=3)
    return build_mesh(flux_grid(), st), st


def test_helical_perturbation_of_the_vertex_potential(plain, helical):
    mesh, st = helical
    a0, a1 = plain.tetra_physics, mesh.tetra_physics
    # first vertex of every tetrahedron: A_phi -> A_phi + A_phi eps cos(m theta + n phi); nothing else at the vertex moves
    # (theta / phi of x1 may carry the + 2 pi of the periodic boundary: m, n are integers)
    th, ph = a0[:, 1], a0[:, 2]
    expect = a0[:, APHI1] + a0[:, APHI1] * 0.05 * np.cos(2 * th + 3 * ph)
    assert np.abs(a1[:, APHI1] - expect).max() <= 4e-15 * np.abs(expect).max()
    inner = a0[:, APHI1] != 0.0
    assert np.abs(a1[inner, APHI1] / a0[inner, APHI1] - 1).max() > 0.04
    for k in (0, 1, 2, BMOD1, ATHETA1, H1, H1 + 1, H1 + 2):
        assert np.array_equal(a0[:, k], a1[:, k]), k
    assert np.array_equal(plain.tetra_grid, mesh.tetra_grid)
    # the perturbed field has a radial component: curl A gets an s-component ~ d(A_phi)/d(theta) that the axisymmetric
    # equilibrium in flux coordinates does not have
    assert np.abs(a0[:, CURLA]).max() < 1e-6 * np.abs(a0[:, CURLA + 1]).max()
    assert np.abs(a1[:, CURLA]).max() > 1e-3 * np.abs(a1[:, CURLA + 1]).max()


def test_helical_perturbation_breaks_p_phi_but_not_the_energy(plain, helical):
    """Orbits on the perturbed mesh: oracle == device algorithm (host compile) bit for bit; the toroidal momentum is no
    longer an invariant (n != 0), magnetic moment and energy still are."""
    mesh, st = helical
    drift = {}
    for label, m in (("plain", plain), ("helical", mesh)):
        om, hm = OracleMesh(m, st), HostMirror(m, st)
        n = 96
        xa, va, wa = workloads.particles_flux(n, 77, s_lo=0.3, s_hi=0.7)
        xb, vb, wb = xa.copy(), va.copy(), wa.copy()
        ia, ta, fa = workloads.fresh_state(n)
        ib, tb, fb = workloads.fresh_state(n)
        om.orbit_timestep_batch(xa, va, wa, 0.0, ia, ta, fa)
        hm.orbit_timestep(xb, vb, wb, 0.0, ib, tb, fb, 0)
        assert np.array_equal(ta, tb) and (ta > 0).all()
        e0, p0, mu0 = om.invariants(xa, va, wa, ta)
        ra = om.orbit_timestep_trace(xa, va, wa, 1e-4, ia, ta, fa, 128)
        rb = hm.orbit_timestep(xb, vb, wb, 1e-4, ib, tb, fb, 128)
        assert ra["n_pushes"].sum() > 5000
        assert np.array_equal(ra["trace_tetr"], rb["trace_tetr"]) and np.array_equal(ra["trace_face"], rb["trace_face"])
        assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(ta, tb)
        ok = ta > 0
        e1, p1, mu1 = om.invariants(xa, va, wa, ta)
        assert np.abs(mu1 / mu0 - 1)[ok].max() < 1e-13
        assert np.abs(e1 / e0 - 1)[ok].max() < 1e-3
        drift[label] = np.abs(p1 - p0)[ok].max() / np.abs(p0[ok]).mean()
    assert drift["helical"] > 20 * drift["plain"], drift


@pytest.mark.gpu
def test_gpu_parity_on_the_helically_perturbed_mesh(helical, cuda_device):
    from gorilla_b200 import Gorilla
    mesh, st = helical
    g, om = Gorilla(mesh, st), OracleMesh(mesh, st)
    n = 400
    xa, va, wa = workloads.particles_flux(n, 78, s_lo=0.3, s_hi=0.7)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    ra = om.orbit_timestep_trace(xa, va, wa, 2e-5, *sa, 128)
    tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, 2e-5, *sb, trace_cap=128)
    assert np.array_equal(ra["trace_tetr"], tt) and np.array_equal(ra["trace_face"], tf)
    assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(sa[1], sb[1])
    g.close()


def test_helical_settings_come_from_the_namelist(tmp_path):
    (tmp_path / "gorilla.inp").write_text(
        "&GORILLANML\n boole_helical_pert = .true. ,\n helical_pert_eps_Aphi = 1.d-1 ,\n helical_pert_m_fourier = 2 ,\n"
        " helical_pert_n_fourier = 2 ,\n/\n")
    s = load_gorilla_inp(tmp_path / "gorilla.inp")
    assert s.boole_helical_pert is True and s.helical_pert_eps_Aphi == 0.1
    assert (s.helical_pert_m_fourier, s.helical_pert_n_fourier) == (2, 2)


@pytest.mark.parametrize("kind", ["analytic", "efit_flux", "vmec"])
def test_bmod_multiplier_scales_the_field_modulus_only(product_lib, kind):
    """|B| at the vertices times a power of two: bmod1 and grad |B| scale exactly, the unit vector h = B / |B| by the exact
    inverse, the vector potential (hence curl A) not at all; 0 (not given) and 1 are the same mesh."""
    if kind == "analytic":
        grid, st = workloads.analytic_tokamak(6, 6, 6)
    elif kind == "efit_flux":
        grid, st = flux_grid(), flux_settings()
    else:
        grid, st = workloads.vmec_qi(DATA / "netcdf_file_for_test.nc", 6, 6, 8)
    base = build_mesh(grid, st).tetra_physics.copy()
    same = build_mesh(dataclasses.replace(grid, bmod_multiplier=0.0), st).tetra_physics
    assert np.array_equal(base, same)
    big = build_mesh(dataclasses.replace(grid, bmod_multiplier=4.0), st).tetra_physics
    assert np.array_equal(big[:, BMOD1], 4.0 * base[:, BMOD1])
    assert np.array_equal(big[:, GB:GB + 3], 4.0 * base[:, GB:GB + 3])
    assert np.array_equal(big[:, H1:H1 + 3], 0.25 * base[:, H1:H1 + 3])
    for k in (ATHETA1, APHI1, CURLA, CURLA + 1, CURLA + 2, 0, 1, 2):
        assert np.array_equal(big[:, k], base[:, k]), k


def test_psi_window_filter_of_field_divB0_inp(product_lib):
    """nwindow_r / nwindow_z (field_divB0.inp; bdivfree.f90:1144-1164, window_filter utils_bdivfree.f90:859-872): the psi(R, Z)
    table is smoothed by a centred moving average over R, then over Z, before it is splined.  Independent check: the same
    filter in numpy on an independent parse of the g-file + scipy's quintic spline against A_phi at the vertices."""
    from scipy.interpolate import RectBivariateSpline
    from test_efit_mesh import GFILE, parse_gfile
    st = GorillaSettings(eps_Phi=0.0, coord_system=1, ispecies=2, boole_periodic_relocation=False, ipusher=2,
                         poly_order=2, boole_guess=True)
    grid = TetraGridSettings(grid_kind=1, n1=32, n2=4, n3=48, boole_n_field_periods=True, g_file_filename=str(GFILE),
                             convex_wall_filename=str(DATA / "convex_wall_for_test.dat"))
    base = build_mesh(grid, st).tetra_physics.copy()
    assert np.array_equal(base, build_mesh(dataclasses.replace(grid, nwindow_r=0, nwindow_z=0), st).tetra_physics)
    nwr, nwz = 3, 2
    filt = build_mesh(dataclasses.replace(grid, nwindow_r=nwr, nwindow_z=nwz), st).tetra_physics
    nw, nh, v = parse_gfile(GFILE)
    xdim, zdim, rzero, r1, zmid = v[0:5]
    psi_axis = v[7]
    psi = v[20 + 4 * nw:20 + 4 * nw + nw * nh].reshape(nh, nw).T          # psi[i, j] = psiRZ(i+1, j+1)

    def window(a, half):                                                   # along axis 0
        out = np.empty_like(a)
        n = a.shape[0]
        for i in range(n):
            k = min(half, i, n - 1 - i)
            out[i] = a[i - k:i + k + 1].sum(axis=0) / (2 * k + 1)
        return out

    psi_f = window(window(psi, nwr).T, nwz).T
    rad = (r1 + np.arange(nw) * (xdim / (nw - 1))) * 1e2
    zet = (zmid - zdim / 2 + np.arange(nh) * (zdim / (nh - 1))) * 1e2
    spl = RectBivariateSpline(rad, zet, (psi_f - psi_axis) * 1e8, kx=5, ky=5)
    R, Z = filt[:, 0], filt[:, 2]
    core = (R > 115) & (R < 215) & (Z > -100) & (Z < 100)                  # inside the convex wall: no stretching
    span = np.abs((psi - psi_axis) * 1e8).max()
    assert np.abs(filt[core, APHI1] - spl.ev(R[core], Z[core])).max() < 2e-6 * span
    # and the filter did something: the smoothed flux differs from the raw one by much more than that
    assert np.abs(filt[core, APHI1] - base[core, APHI1]).max() > 1e-4 * span


def test_invalid_mesh_options_are_refused(product_lib):
    from gorilla_b200 import api
    with pytest.raises(api.GorillaError) as ei:
        build_mesh(flux_grid(theta_geom_flux=3), flux_settings())
    assert ei.value.code == 1 and "theta_geom_flux" in str(ei.value)


def test_binding_checks_its_struct_layouts_against_the_library(product_lib):
    """gorilla_b200_abi_struct_sizes: sizeof of every struct of the header as compiled; load_library() compares them with the
    ctypes mirrors (and the Fortran module with its bind(C) types) so that a stale binding fails at start-up."""
    import ctypes as C
    from gorilla_b200 import api
    sizes = (C.c_int64 * 7)()
    assert product_lib.gorilla_b200_abi_struct_sizes(sizes) == 0
    assert tuple(sizes) == (C.sizeof(api._Settings), C.sizeof(api._MeshDesc), C.sizeof(api._Counters), C.sizeof(api._Diag),
                            C.sizeof(api._GridSettings), api.EVENT_DTYPE.itemsize, C.sizeof(api._EventSettings))
    assert api.EVENT_DTYPE.itemsize == 72 and C.sizeof(api._EventSettings) == 32
    assert product_lib.gorilla_b200_abi_struct_sizes(None) == 1


def test_vertex_noise_options(product_lib):
    """boole_axi_noise_vector_pot / boole_non_axi_noise_vector_pot / boole_axi_noise_elec_pot (gorilla.inp:84-109;
    tetra_physics_mod.f90:256-261,400-415,441-444): A_k += A_k eps r with r uniform in [0, 1), the same r in every poloidal
    plane (axisymmetric) or fresh per vertex and component; the electrostatic potential follows the noisy A_2 and can get
    its own axisymmetric noise.  The stream is the library's own (deterministic per noise_seed).  On a field-aligned grid,
    where nvert / grid_size(2) is the number of vertices of a poloidal plane (on the rectangular grid, with its n2 + 1
    planes, the reference's index arithmetic does not repeat plane by plane; the library follows the same arithmetic)."""
    grid, st0 = flux_grid(), dataclasses.replace(flux_settings(), eps_Phi=-1e-5)
    base = build_mesh(grid, st0)
    tp0, tg = base.tetra_physics, base.tetra_grid
    # poloidal position of the first vertex of every tetrahedron -> group index (the same (R, Z) in every phi plane)
    rz = np.round(np.column_stack([tp0[:, 31], tp0[:, 32]]), 7)
    _, group = np.unique(rz, axis=0, return_inverse=True)
    group = group.ravel()

    def spread_within_groups(v):
        lo = np.full(group.max() + 1, np.inf); hi = np.full(group.max() + 1, -np.inf)
        np.minimum.at(lo, group, v); np.maximum.at(hi, group, v)
        return (hi - lo).max()

    sel = np.abs(tp0[:, APHI1]) > 1e-3 * np.abs(tp0[:, APHI1]).max()

    def ratio(tp):   # (A_3' / A_3 - 1) at the first vertex of every tetrahedron, A_3 = A_phi = psi_pol in flux coordinates
        return tp[sel, APHI1] / tp0[sel, APHI1] - 1

    eps = 0.05
    axi = build_mesh(grid, dataclasses.replace(st0, boole_axi_noise_vector_pot=True, axi_noise_eps_A=eps)).tetra_physics
    r = ratio(axi) / eps
    assert r.min() >= -1e-12 and r.max() < 1 and 0.3 < r.mean() < 0.7 and r.std() > 0.2
    # axisymmetric: the same factor at corresponding tetrahedra of every phi slice
    full = axi[:, APHI1] / np.where(tp0[:, APHI1] == 0, 1, tp0[:, APHI1])
    assert np.bincount(group).min() >= grid.n2 and spread_within_groups(full) < 1e-13
    # the potential follows: Phi_1 = A_2 eps_Phi at the first vertex (TP_PHI1 = 30)
    assert np.allclose(axi[sel, 30], axi[sel, ATHETA1] * -1e-5, rtol=1e-13, atol=0)
    # deterministic, and a different seed gives different numbers
    again = build_mesh(grid, dataclasses.replace(st0, boole_axi_noise_vector_pot=True, axi_noise_eps_A=eps)).tetra_physics
    other = build_mesh(grid, dataclasses.replace(st0, boole_axi_noise_vector_pot=True, axi_noise_eps_A=eps, noise_seed=7)).tetra_physics
    assert np.array_equal(axi, again) and not np.array_equal(axi, other)
    # non-axisymmetric: differs between slices
    non = build_mesh(grid, dataclasses.replace(st0, boole_non_axi_noise_vector_pot=True, non_axi_noise_eps_A=eps)).tetra_physics
    r = ratio(non) / eps
    assert r.min() >= -1e-12 and r.max() < 1 and 0.3 < r.mean() < 0.7
    fulln = non[:, APHI1] / np.where(tp0[:, APHI1] == 0, 1, tp0[:, APHI1])
    assert spread_within_groups(fulln) > 0.3 * eps
    # noise on the electrostatic potential alone leaves A untouched
    pot = build_mesh(grid, dataclasses.replace(st0, boole_axi_noise_elec_pot=True, axi_noise_eps_Phi=0.3)).tetra_physics
    assert np.array_equal(pot[:, APHI1], tp0[:, APHI1]) and np.array_equal(pot[:, BMOD1], tp0[:, BMOD1])
    rp = (pot[sel, 30] / tp0[sel, 30] - 1) / 0.3
    assert rp.min() >= -1e-12 and rp.max() < 1 and rp.std() > 0.2
    assert np.array_equal(tg, build_mesh(grid, dataclasses.replace(st0, boole_axi_noise_elec_pot=True)).tetra_grid)


def test_orbits_on_a_noisy_mesh(product_lib):
    """What the options are for: a rough field.  Oracle and device algorithm (host compile) stay bit-identical on it."""
    grid, st0 = workloads.analytic_tokamak(10, 10, 10)
    st = dataclasses.replace(st0, boole_non_axi_noise_vector_pot=True, non_axi_noise_eps_A=1e-4)
    mesh = build_mesh(grid, st)
    om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
    n = 100
    xa, va, wa = workloads.particles_cyl(n, 3)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    ra = om.orbit_timestep_trace(xa, va, wa, 1e-4, ia, ta, fa, 256)
    rb = hm.orbit_timestep(xb, vb, wb, 1e-4, ib, tb, fb, 256)
    assert ra["n_pushes"].sum() > 8000
    assert np.array_equal(ra["trace_tetr"], rb["trace_tetr"]) and np.array_equal(ra["trace_face"], rb["trace_face"])
    assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(ta, tb)
