"""handover_processing_kind = 2: position exchange between tetrahedra via Cartesian skew coordinates,
pusher_handover2neighbour, pusher_tetra_func_mod.f90:59-89; the tetra_skew_coord records, tetra_physics_mod.f90:89-99,946-1011.
Polynomial pusher (EXT = 2 kernels)."""
import numpy as np
import pytest

import workloads
from gorilla_b200 import api, build_mesh
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def _with(settings, **kw):
    return type(settings)(**{**settings.__dict__, **kw})


@pytest.fixture(scope="module")
def skew_mesh(product_lib):
    grid, settings = workloads.analytic_tokamak(14, 14, 14)
    settings.handover_processing_kind = 2
    return build_mesh(grid, settings), grid, settings


def test_skew_records_follow_the_reference_formulas(skew_mesh):
    mesh, _, _ = skew_mesh
    S, tp, tg = mesh.tetra_skew_coord, mesh.tetra_physics, mesh.tetra_grid
    assert S.shape == (mesh.ntetr, 168) and np.isfinite(S).all()
    mat = lambda off: S[:, off:off + 36].reshape(-1, 4, 3, 3).transpose(0, 1, 3, 2)   # [t, face, i, j] from column-major
    A, C, Ai, Ci = mat(0), mat(36), mat(72), mat(108)
    eye = np.eye(3)
    assert np.abs(np.einsum("tkij,tkjl->tkil", A, Ai) - eye).max() < 1e-12
    assert np.abs(np.einsum("tkij,tkjl->tkil", C, Ci) - eye).max() < 1e-12
    rx, rc = S[:, 144:156].reshape(-1, 4, 3), S[:, 156:168].reshape(-1, 4, 3)
    # the reference vertex of face k is vertex modulo(k,4)+1 (never the vertex opposite the face); face 4 -> vertex 1 = x1
    assert same(rx[:, 3], tp[:, 0:3])
    v = mesh.verts_rphiz[tg[:, 0] - 1]
    assert same(rc[:, 3], np.stack([v[:, 0] * np.cos(v[:, 1]), v[:, 0] * np.sin(v[:, 1]), v[:, 2]], 1))
    # two of the three skew vectors of face k lie in the plane of face k: n_k . column = 0 for them
    an = tp[:, 9:21].reshape(-1, 4, 3)
    for k in range(4):
        dots = np.einsum("ti,tij->tj", an[:, k], A[:, k])
        assert (np.abs(dots) < 1e-9 * np.abs(an[:, k]).max()).sum(axis=1).min() >= 2


def test_same_orbits_as_kind_1(skew_mesh):
    """Both tetrahedra share the vertices of the hand-over face, so the linear map through Cartesian space reproduces the
    plain hand-over (with its periodic shifts) to rounding."""
    mesh, _, settings = skew_mesh
    out = {}
    for kind in (1, 2):
        om = OracleMesh(mesh, _with(settings, poly_order=2, handover_processing_kind=kind))
        x, vpar, vperp = workloads.particles_cyl(200, 3)
        s = workloads.fresh_state(200)
        r = om.orbit_timestep_trace(x, vpar, vperp, 1e-5, *s, 24)
        out[kind] = (x, vpar, r, s[1])
    ok = (out[1][3] > 0) & (out[2][3] > 0)
    d = out[1][0][ok] - out[2][0][ok]
    d[:, 1] = (d[:, 1] + np.pi) % (2 * np.pi) - np.pi
    assert ok.sum() > 190 and np.abs(d).max() < 1e-10 and not same(out[1][0], out[2][0])
    assert same(out[1][2]["trace_tetr"][ok], out[2][2]["trace_tetr"][ok])


@pytest.mark.parametrize("K", [0, 1, 2, 3, 4])
def test_host_mirror_parity(skew_mesh, K):
    mesh, _, settings = skew_mesh
    if K == 0:      # RK4 pusher
        cases = ((_with(settings, ipusher=1), False, 5e-6), (_with(settings, ipusher=1), True, -4e-6))
    else:
        cases = ((_with(settings, poly_order=K), False, 5e-6),
                 (_with(settings, poly_order=K, i_time_tracing_option=2, boole_time_Hamiltonian=True), True, -4e-6))
    for st, ff, t in cases:
        om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
        xa, va, wa = workloads.particles_cyl(120, 4)
        xb, vb, wb = xa.copy(), va.copy(), wa.copy()
        sa, sb = workloads.fresh_state(120), workloads.fresh_state(120)
        for _ in range(2):
            ra = om.orbit_timestep_trace(xa, va, wa, t, *sa, 32)
            rb = hm.orbit_timestep(xb, vb, wb, t, *sb, trace_cap=32, force_full=ff, optional=st.boole_time_Hamiltonian)
            assert same(ra["trace_tetr"], rb["trace_tetr"]) and same(ra["trace_face"], rb["trace_face"])
            assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(ra["t_remain"], rb["t_remain"])
            assert same(sa[1], sb[1]) and same(sa[2], sb[2])
            if st.boole_time_Hamiltonian:
                assert same(ra["optional_quantities"], rb["optional_quantities"])
        assert ra["n_pushes"].sum() > 500


def _face_starts(grid, n, seed):
    """start points of which three in five lie exactly on a cell face (R, Z or phi plane of the vertex grid)"""
    xa, va, wa = workloads.particles_cyl(n, seed, rmin_frac=0.05, rmax_frac=0.95)
    hr, hz, hphi = 100.0 / grid.n1, 100.0 / grid.n3, 2 * np.pi / grid.n2
    xa[0::5, 0] = 120.0 + hr * np.round((xa[0::5, 0] - 120.0) / hr)
    xa[1::5, 2] = -50.0 + hz * np.round((xa[1::5, 2] + 50.0) / hz)
    xa[2::5, 1] = hphi * np.floor(xa[2::5, 1] / hphi)
    return xa, va, wa


def test_start_on_a_face_is_handed_over_with_skew_coordinates(skew_mesh):
    """find_tetra's neighbour hop for a start point on a face goes through pusher_handover2neighbour, which
    honours handover_processing_kind = 2 (find_tetra_mod.f90:283-600); oracle <-> host compile of the device headers."""
    mesh, grid, settings = skew_mesh
    st = _with(settings, poly_order=2)
    om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
    n = 1500
    xa, va, wa = _face_starts(grid, n, 17)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    ra = om.orbit_timestep_trace(xa, va, wa, 2e-6, *sa, 16)
    rb = hm.orbit_timestep(xb, vb, wb, 2e-6, *sb, trace_cap=16)
    assert same(ra["trace_tetr"], rb["trace_tetr"]) and same(ra["trace_face"], rb["trace_face"])
    assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(sa[1], sb[1]) and same(sa[2], sb[2])
    assert (sa[1] > 0).sum() > 0.9 * n


@pytest.mark.gpu
def test_gpu_start_on_a_face(skew_mesh, cuda_device):
    from gorilla_b200 import Gorilla
    mesh, grid, settings = skew_mesh
    st = _with(settings, poly_order=2)
    om, g = OracleMesh(mesh, st), Gorilla(mesh, st)
    n = 3000
    xa, va, wa = _face_starts(grid, n, 19)
    xb = xa.copy()
    ta, fa = om.find_tetra(xa, va, wa)
    tb, fb = g.find_tetra(xb, va, wa)
    assert same(ta, tb) and same(fa, fb) and same(xa, xb) and (fa > 0).sum() > 100
    g.close()


def test_settings_rules(product_lib, skew_mesh, small_mesh):
    mesh, _, settings = skew_mesh
    for bad, code in ((_with(settings, boole_adaptive_time_steps=True, desired_delta_energy=1e-10, max_n_intermediate_steps=10), 2),
                      (_with(settings, handover_processing_kind=3), 1)):
        with pytest.raises(api.GorillaError) as ei:
            api.Gorilla(mesh, bad)
        assert ei.value.code == code
    plain_mesh, _, plain_settings = small_mesh          # built without the skew records
    with pytest.raises(api.GorillaError) as ei:
        api.Gorilla(plain_mesh, _with(plain_settings, handover_processing_kind=2))
    assert ei.value.code == 1


@pytest.mark.gpu
@pytest.mark.parametrize("K", [0, 2, 4])
def test_gpu_parity(skew_mesh, cuda_device, K):
    from gorilla_b200 import Gorilla
    mesh, _, settings = skew_mesh
    st = _with(settings, ipusher=1) if K == 0 else _with(settings, poly_order=K)
    om, g = OracleMesh(mesh, st), Gorilla(mesh, st)
    n = 500
    xa, va, wa = workloads.particles_cyl(n, 4)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    for _ in range(2):
        ra = om.orbit_timestep_trace(xa, va, wa, 6e-6, *sa, 64)
        tro, npu = np.zeros(n), np.zeros(n, np.int64)
        tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, 6e-6, *sb, t_remain_out=tro, n_pushes=npu, trace_cap=64)
        assert same(ra["trace_tetr"], tt) and same(ra["trace_face"], tf), "visited tetra sequence differs"
        assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(ra["t_remain"], tro) and same(ra["n_pushes"], npu)
        assert same(sa[1], sb[1]) and same(sa[2], sb[2])
    assert g.counters().n_pushes > 3000
    g.close()
