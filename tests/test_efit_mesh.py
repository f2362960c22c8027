"""grid_kind = 1: rectangular (R, phi, Z) grid over an EFIT equilibrium (g-file reader, quintic psi spline, fpol spline,
convex-wall coordinate stretching; gorilla_b200/csrc/host/mesh_efit.cpp).  The reader and the interpolant are checked
against an independent Python parse of the same g-file and scipy's quintic spline (different end conditions: compared
away from the box edge), then the mesh carries orbits through the oracle and the device algorithm."""
from pathlib import Path

import numpy as np
import pytest

import workloads
from gorilla_b200 import GorillaSettings, TetraGridSettings, build_mesh
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh

DATA = Path(__file__).resolve().parent.parent / "data" / "equilibria"
GFILE = DATA / "g_file_for_test"


def parse_gfile(path):
    lines = path.read_text().splitlines()
    nw, nh = int(lines[0][52:56]), int(lines[0][56:60])
    vals = []
    for ln in lines[1:]:
        for k in range(0, len(ln), 16):
            f = ln[k:k + 16]
            if f.strip():
                try:
                    vals.append(float(f))
                except ValueError:
                    return nw, nh, np.array(vals)
    return nw, nh, np.array(vals)


@pytest.fixture(scope="module")
def efit_mesh(product_lib):
    grid = TetraGridSettings(grid_kind=1, n1=48, n2=12, n3=72, boole_n_field_periods=True,
                             g_file_filename=str(GFILE), convex_wall_filename=str(DATA / "convex_wall_for_test.dat"))
    st = GorillaSettings(eps_Phi=0.0, coord_system=1, ispecies=2, boole_periodic_relocation=False, ipusher=2,
                         poly_order=2, boole_guess=True)
    return build_mesh(grid, st), grid, st


def test_vertex_fields_agree_with_an_independent_spline(efit_mesh):
    from scipy.interpolate import RectBivariateSpline
    mesh, grid, _ = efit_mesh
    nw, nh, v = parse_gfile(GFILE)
    xdim, zdim, rzero, r1, zmid = v[0:5]
    psi_axis, psi_sep, bt0 = v[7], v[8], v[9]
    fpol = v[20:20 + nw]
    psi = v[20 + 4 * nw:20 + 4 * nw + nw * nh].reshape(nh, nw).T            # psi[i, j] = psiRZ(i+1, j+1)
    rad = (r1 + np.arange(nw) * (xdim / (nw - 1))) * 1e2
    zet = (zmid - zdim / 2 + np.arange(nh) * (zdim / (nh - 1))) * 1e2
    psi_cgs = (psi - psi_axis) * 1e8
    d = mesh.desc()
    assert d.Rmin == rad[0] and d.Rmax == rad[-1] and d.Zmin == zet[0] and d.Zmax == zet[-1]
    spl = RectBivariateSpline(rad, zet, psi_cgs, kx=5, ky=5)
    tp = mesh.tetra_physics
    R, Z = tp[:, 0], tp[:, 2]                                                # first vertex x1 = (R, phi, Z)
    core = (R > 115) & (R < 215) & (Z > -100) & (Z < 100)                    # inside the convex wall: no stretching
    assert core.sum() > 10000
    span = psi_cgs.max() - psi_cgs.min()
    assert np.abs(tp[core, 26] - spl.ev(R[core], Z[core])).max() < 2e-6 * span   # A_phi at the first vertex = psi
    dpr, dpz = spl.ev(R[core], Z[core], dx=1), spl.ev(R[core], Z[core], dy=1)
    psihat = np.clip(spl.ev(R[core], Z[core]) / ((psi_sep - psi_axis) * 1e8), 0, None)
    from scipy.interpolate import CubicSpline
    F = np.where(psihat > 1, fpol[-1], CubicSpline(np.linspace(0, 1, nw), fpol)(np.minimum(psihat, 1))) * 1e6
    Bmod = np.sqrt((dpz / R[core]) ** 2 + (dpr / R[core]) ** 2 + (F / R[core]) ** 2)
    assert np.abs(tp[core, 24] / Bmod - 1).max() < 1e-5
    # unit vector h: covariant components, |h|^2 = h_R^2 + (h_phi/R)^2 + h_Z^2 = 1
    h = tp[:, 27:30]
    assert np.abs(h[:, 0] ** 2 + (h[:, 1] / R) ** 2 + h[:, 2] ** 2 - 1).max() < 1e-12
    assert mesh.n_overlaps == 0 if hasattr(mesh, "n_overlaps") else True


def test_orbits_oracle_vs_device_algorithm_and_invariants(efit_mesh):
    mesh, _, st = efit_mesh
    for K in (2, 4):
        s = type(st)(**{**st.__dict__, "poly_order": K})
        om, hm = OracleMesh(mesh, s), HostMirror(mesh, s)
        n = 120
        xa, va, wa = workloads.particles_cyl(n, 5, R0=165.0, a=40.0)
        xb, vb, wb = xa.copy(), va.copy(), wa.copy()
        ia, ta, fa = workloads.fresh_state(n)
        ib, tb, fb = workloads.fresh_state(n)
        om.orbit_timestep_batch(xa, va, wa, 0.0, ia, ta, fa)
        e0, p0, mu0 = om.invariants(xa, va, wa, ta)
        ra = om.orbit_timestep_trace(xa, va, wa, 2e-5, ia, ta, fa, 256)
        rb = hm.orbit_timestep(xb, vb, wb, 2e-5, ib, tb, fb, 256)
        assert (ta > 0).sum() > 110 and ra["n_pushes"].sum() > 3000
        assert np.array_equal(ra["trace_tetr"], rb["trace_tetr"]) and np.array_equal(ra["trace_face"], rb["trace_face"])
        assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(ta, tb)
        e1, p1, mu1 = om.invariants(xa, va, wa, ta)
        ok = ta > 0
        assert np.abs(mu1 / mu0 - 1)[ok].max() < 1e-13
        dp = np.abs(p1 - p0)[ok].max() / np.abs(p0[ok]).mean()                 # axisymmetric: p_phi conserved
        assert dp < (1e-2 if K == 2 else 1e-9), dp     # order 2 on 12 toroidal cells: ~1e-3, as on the analytic grid
        assert np.abs(e1 / e0 - 1)[ok].max() < (1e-4 if K == 2 else 1e-11)


@pytest.mark.gpu
def test_gpu_parity_on_the_efit_mesh(efit_mesh, cuda_device):
    from gorilla_b200 import Gorilla
    mesh, _, st = efit_mesh
    om, g = OracleMesh(mesh, st), Gorilla(mesh, st)
    n = 500
    xa, va, wa = workloads.particles_cyl(n, 6, R0=165.0, a=40.0)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    ra = om.orbit_timestep_trace(xa, va, wa, 2e-5, ia, ta, fa, 128)
    tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, 2e-5, ib, tb, fb, trace_cap=128)
    assert np.array_equal(ra["trace_tetr"], tt) and np.array_equal(ra["trace_face"], tf)
    assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(ta, tb)
    g.close()


# ---------------------------------------------------------------------------------------------- grid_kind = 2
@pytest.fixture(scope="module")
def efit_flux_mesh(product_lib):
    grid = TetraGridSettings(grid_kind=2, n1=40, n2=8, n3=32, boole_n_field_periods=True, sfc_s_min=0.1,
                             g_file_filename=str(GFILE), convex_wall_filename=str(DATA / "convex_wall_for_test.dat"))
    st = GorillaSettings(eps_Phi=0.0, coord_system=2, ispecies=2, boole_periodic_relocation=True, ipusher=2,
                         poly_order=2, boole_guess=True)
    return build_mesh(grid, st), grid, st


def test_flux_coordinates_reproduce_the_equilibrium(efit_flux_mesh):
    """The symmetry flux coordinates are constructed by field-line integration.  Independent checks against the g-file:
    the poloidal flux stored at every vertex equals the psi(R, Z) table at the vertex position, the safety factor
    q = d(psi_tor)/d(psi_pol) of the constructed surfaces follows the g-file's own q profile, and the poloidal angle is
    a straight-field-line angle (the Jacobian R^2 psitor_max / (h_phi B) used for dt/dtau is positive and smooth)."""
    from scipy.interpolate import RectBivariateSpline
    mesh, grid, _ = efit_flux_mesh
    nw, nh, v = parse_gfile(GFILE)
    xdim, zdim, rzero, r1, zmid = v[0:5]
    psi_axis, psi_sep = v[7], v[8]
    psi = v[20 + 4 * nw:20 + 4 * nw + nw * nh].reshape(nh, nw).T
    qpsi = v[20 + 4 * nw + nw * nh:20 + 5 * nw + nw * nh]
    rad = (r1 + np.arange(nw) * (xdim / (nw - 1))) * 1e2
    zet = (zmid - zdim / 2 + np.arange(nh) * (zdim / (nh - 1))) * 1e2
    spl = RectBivariateSpline(rad, zet, (psi - psi_axis) * 1e8, kx=5, ky=5)
    tp = mesh.tetra_physics
    s, R, Z, Aphi, Ath = tp[:, 0], tp[:, 31], tp[:, 32], tp[:, 26], tp[:, 25]
    span = abs(psi_sep - psi_axis) * 1e8
    # (1) psi at the vertex position; the axis found by field-line averaging is not exactly the table's minimum
    assert np.abs(Aphi - spl.ev(R, Z)).max() < 2e-4 * span
    # (2) q profile: ring-to-ring finite difference of A_theta = s psitor_max against A_phi = psi_pol
    rings = np.unique(np.round(s, 12))
    a_phi = np.array([Aphi[np.isclose(s, r)].mean() for r in rings])
    a_th = np.array([Ath[np.isclose(s, r)].mean() for r in rings])
    q_mesh = np.abs(np.diff(a_th) / np.diff(a_phi))
    psin_mid = 0.5 * (a_phi[1:] + a_phi[:-1]) / span
    q_file = np.interp(np.abs(psin_mid), np.linspace(0, 1, nw), np.abs(qpsi))
    inner = np.abs(psin_mid) < 0.9
    assert inner.sum() > 20
    assert np.abs(q_mesh[inner] / q_file[inner] - 1).max() < 0.03
    # (3) flux surfaces: psi is constant on a ring to interpolation accuracy
    for r in rings[::5]:
        sel = np.isclose(s, r)
        assert np.ptp(Aphi[sel]) < 2e-4 * span
    assert np.all(tp[:, 40] > 0) and mesh.desc().sign_sqg == 1     # dt/dtau = sqrt(g) B > 0


def test_orbits_on_the_field_aligned_efit_mesh(efit_flux_mesh):
    mesh, grid, st = efit_flux_mesh
    for K in (2, 4):
        sK = type(st)(**{**st.__dict__, "poly_order": K})
        om, hm = OracleMesh(mesh, sK), HostMirror(mesh, sK)
        n = 120
        rng = np.random.Generator(np.random.PCG64(9))
        xa = np.column_stack([0.2 + 0.6 * rng.random(n), 2 * np.pi * rng.random(n), 2 * np.pi * rng.random(n)])
        lam = 2 * rng.random(n) - 1
        vmod = np.sqrt(2.0 * 3.0e3 * workloads.EV2ERG / (2.0 * workloads.AMP))
        va, wa = lam * vmod, vmod * np.sqrt(1 - lam ** 2)
        xb, vb, wb = xa.copy(), va.copy(), wa.copy()
        ia, ta, fa = workloads.fresh_state(n)
        ib, tb, fb = workloads.fresh_state(n)
        om.orbit_timestep_batch(xa, va, wa, 0.0, ia, ta, fa)
        hm.orbit_timestep(xb, vb, wb, 0.0, ib, tb, fb, 0)
        assert ia.all() and np.array_equal(ta, tb)
        e0, p0, mu0 = om.invariants(xa, va, wa, ta)
        ra = om.orbit_timestep_trace(xa, va, wa, 2e-5, ia, ta, fa, 256)
        rb = hm.orbit_timestep(xb, vb, wb, 2e-5, ib, tb, fb, 256)
        assert ra["n_pushes"].sum() > 2000 and (ta > 0).sum() > 100
        assert np.array_equal(ra["trace_tetr"], rb["trace_tetr"]) and np.array_equal(ra["trace_face"], rb["trace_face"])
        assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(ta, tb)
        e1, p1, mu1 = om.invariants(xa, va, wa, ta)
        ok = ta > 0
        assert np.abs(mu1 / mu0 - 1)[ok].max() < 1e-13
        assert np.abs(e1 / e0 - 1)[ok].max() < (1e-4 if K == 2 else 1e-10)
        dp = np.abs(p1 - p0)[ok].max() / np.abs(p0[ok]).mean()
        assert dp < (1e-2 if K == 2 else 1e-8), dp


@pytest.mark.gpu
def test_gpu_parity_on_the_field_aligned_efit_mesh(efit_flux_mesh, cuda_device):
    from gorilla_b200 import Gorilla
    mesh, _, st = efit_flux_mesh
    s4 = type(st)(**{**st.__dict__, "poly_order": 4})
    for s in (st, s4):
        om, g = OracleMesh(mesh, s), Gorilla(mesh, s)
        n = 400
        xa, va, wa = workloads.particles_flux(n, 6)
        xb, vb, wb = xa.copy(), va.copy(), wa.copy()
        ia, ta, fa = workloads.fresh_state(n)
        ib, tb, fb = workloads.fresh_state(n)
        ra = om.orbit_timestep_trace(xa, va, wa, 2e-5, ia, ta, fa, 128)
        tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, 2e-5, ib, tb, fb, trace_cap=128)
        assert np.array_equal(ra["trace_tetr"], tt) and np.array_equal(ra["trace_face"], tf)
        assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(ta, tb)
        g.close()


# ------------------------------------------------------------- grid_kind = 2, theta_geom_flux = 2 (points_2d.f90:139-149)
@pytest.fixture(scope="module")
def efit_flux_geom_mesh(product_lib):
    grid = TetraGridSettings(grid_kind=2, n1=24, n2=6, n3=32, boole_n_field_periods=True, sfc_s_min=0.1, theta_geom_flux=2,
                             g_file_filename=str(GFILE), convex_wall_filename=str(DATA / "convex_wall_for_test.dat"))
    st = GorillaSettings(eps_Phi=0.0, coord_system=2, ispecies=2, boole_periodic_relocation=True, ipusher=2,
                         poly_order=2, boole_guess=True)
    return build_mesh(grid, st), grid, st


def test_theta_geom_flux_2_places_the_vertices_at_equidistant_geometrical_angles(efit_flux_geom_mesh, efit_flux_mesh):
    """theta_geom2theta_flux (SRC/points_2d.f90:254-373): with theta_geom_flux = 2 the n3 vertices of every ring sit at
    equidistant GEOMETRICAL poloidal angles around the magnetic axis, measured from the axis -> X-point ray; their flux
    angles (verts_sthetaphi) are what the inversion returned.  Independent check from the vertex positions alone."""
    mesh, grid, _ = efit_flux_geom_mesh
    n1 = mesh.desc().grid_size[0]
    n3 = grid.n3
    vr, vs = mesh.verts_rphiz[:(n1 + 1) * n3], mesh.verts_sthetaphi[:(n1 + 1) * n3]     # first phi slice
    # the magnetic axis is the ONE centre around which every ring is equidistant in the geometrical angle: fit it on three
    # rings, then check all of them
    from scipy.optimize import least_squares

    def steps(c, ring):
        R, Z = vr[ring * n3:(ring + 1) * n3, 0], vr[ring * n3:(ring + 1) * n3, 2]
        ang = np.unwrap(np.arctan2(Z - c[1], R - c[0]))
        return np.diff(np.append(ang, ang[0] + np.sign(ang[1] - ang[0]) * 2 * np.pi))

    fit = least_squares(lambda c: np.concatenate([np.abs(steps(c, r)) - 2 * np.pi / n3 for r in (2, n1 // 2, n1)]),
                        x0=[vr[:n3, 0].mean(), vr[:n3, 2].mean()], xtol=1e-14, ftol=1e-14)
    assert abs(fit.x[0] - vr[:n3, 0].mean()) < 1.0 and abs(fit.x[1] - vr[:n3, 2].mean()) < 1.0      # [cm]
    for ring in range(n1 + 1):
        th = vs[ring * n3:(ring + 1) * n3, 1]
        step = steps(fit.x, ring)
        assert np.all(step > 0) or np.all(step < 0)
        assert np.abs(np.abs(step) - 2 * np.pi / n3).max() < 2e-5, (ring, np.abs(np.abs(step) - 2 * np.pi / n3).max())
        # the flux angles are a monotonic re-parametrisation that starts at exactly 0 and is NOT equidistant
        assert th[0] == 0.0 and np.all(np.diff(th) > 0) and th[-1] < 2 * np.pi
    outer = vs[n1 * n3:(n1 + 1) * n3, 1]
    assert np.abs(np.diff(outer) - 2 * np.pi / n3).max() > 0.02
    # same surfaces as the flux-angle grid: s of the rings, psi on them
    m1 = efit_flux_mesh[0]
    assert np.array_equal(np.unique(mesh.verts_sthetaphi[:, 2]).size, grid.n2)
    tp = mesh.tetra_physics
    s, Aphi = tp[:, 0], tp[:, 26]
    span = np.abs(m1.tetra_physics[:, 26]).max()
    for r in np.unique(np.round(s, 12))[::4]:
        sel = np.isclose(s, r)
        assert np.ptp(Aphi[sel]) < 2e-4 * span
    assert np.all(tp[:, 40] > 0)


def test_orbits_on_the_geometrical_angle_grid(efit_flux_geom_mesh):
    """Oracle and device algorithm (host compile) carry the same orbits over the theta_geom_flux = 2 mesh bit for bit, the
    particles cross the theta = 0 / 2 pi seam, and the invariants hold as on the flux-angle grid."""
    mesh, grid, st = efit_flux_geom_mesh
    om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
    n = 120
    rng = np.random.Generator(np.random.PCG64(19))
    xa = np.column_stack([0.25 + 0.5 * rng.random(n), 2 * np.pi * rng.random(n), 2 * np.pi * rng.random(n)])
    lam = 2 * rng.random(n) - 1
    vmod = np.sqrt(2.0 * 3.0e3 * workloads.EV2ERG / (2.0 * workloads.AMP))
    va, wa = lam * vmod, vmod * np.sqrt(1 - lam ** 2)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    om.orbit_timestep_batch(xa, va, wa, 0.0, ia, ta, fa)
    hm.orbit_timestep(xb, vb, wb, 0.0, ib, tb, fb, 0)
    assert ia.all() and np.array_equal(ta, tb) and (ta > 0).all()
    e0, p0, mu0 = om.invariants(xa, va, wa, ta)
    ra = om.orbit_timestep_trace(xa, va, wa, 4e-5, ia, ta, fa, 256)
    rb = hm.orbit_timestep(xb, vb, wb, 4e-5, ib, tb, fb, 256)
    assert ra["n_pushes"].sum() > 2000 and (ta > 0).sum() > 100
    assert np.array_equal(ra["trace_tetr"], rb["trace_tetr"]) and np.array_equal(ra["trace_face"], rb["trace_face"])
    assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(ta, tb)
    e1, p1, mu1 = om.invariants(xa, va, wa, ta)
    ok = ta > 0
    assert np.abs(mu1 / mu0 - 1)[ok].max() < 1e-13
    assert np.abs(e1 / e0 - 1)[ok].max() < 1e-4
    assert np.abs(p1 - p0)[ok].max() / np.abs(p0[ok]).mean() < 1e-2


@pytest.mark.gpu
def test_gpu_parity_on_the_geometrical_angle_grid(efit_flux_geom_mesh, cuda_device):
    from gorilla_b200 import Gorilla
    mesh, _, st = efit_flux_geom_mesh
    g, om = Gorilla(mesh, st), OracleMesh(mesh, st)
    n = 400
    xa, va, wa = workloads.particles_flux(n, 23)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    ra = om.orbit_timestep_trace(xa, va, wa, 2e-5, *sa, 128)
    tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, 2e-5, *sb, trace_cap=128)
    assert np.array_equal(ra["trace_tetr"], tt) and np.array_equal(ra["trace_face"], tf)
    assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(sa[1], sb[1])
    g.close()
