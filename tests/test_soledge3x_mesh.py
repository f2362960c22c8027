"""grid_kind = 4: SOLEDGE3X-EIRENE triangle mesh extruded toroidally over the WEST equilibrium
(gorilla_b200/csrc/host/mesh_soledge3x.cpp), and BASELINE config 4 on it: RK4 pusher with strong-electric-field terms."""
import collections
from pathlib import Path

import numpy as np
import pytest

import workloads
from gorilla_b200 import build_mesh
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh

DATA = Path(__file__).resolve().parent.parent / "data" / "equilibria"


@pytest.fixture(scope="module")
def west_mesh(product_lib):
    grid, st = workloads.west_soledge3x(DATA, n2=6)
    return build_mesh(grid, st), grid, st


def test_topology(west_mesh):
    mesh, grid, _ = west_mesh
    tri = np.loadtxt(DATA / "MESH_SOLEDGE3X_EIRENE" / "triangles_for_test.dat", skiprows=1, dtype=int)
    nt, n2 = len(tri), grid.n2
    tps = 3 * nt
    assert mesh.ntetr == tps * n2
    tg = mesh.tetra_grid
    knot, nb, nf, pphi = tg[:, 0:4], tg[:, 4:8], tg[:, 8:12], tg[:, 12:16]
    # every connection is mutual, over the same face, with opposite periodic flags, and the two cells share the 3
    # vertices of that face
    idx = np.arange(mesh.ntetr)
    for f in range(4):
        ok = nb[:, f] > 0
        j, g = nb[ok, f] - 1, nf[ok, f] - 1
        assert np.all(nb[j, g] == idx[ok] + 1) and np.all(nf[j, g] == f + 1) and np.all(pphi[j, g] == -pphi[ok, f])
        mine = np.sort(np.delete(knot[ok], f, axis=1), axis=1)
        theirs = np.sort(np.where(np.arange(4)[None, :] == g[:, None], -1, knot[j]), axis=1)[:, 1:]
        assert np.array_equal(mine, theirs)
    # in-plane: every pair of triangles sharing an edge is connected (the repair pass leaves no gap), and the only
    # open faces lie over the boundary edges of the 2-D mesh, two per edge and slice
    edges = collections.defaultdict(list)
    for t, (a, b, c) in enumerate(tri):
        for u, v in ((a, b), (b, c), (c, a)):
            edges[(min(u, v), max(u, v))].append(t)
    n_boundary_edges = sum(1 for e in edges.values() if len(e) == 1)
    assert (nb == -1).sum() == 2 * n_boundary_edges * n2
    adj = {(min(e), max(e)) for e in edges.values() if len(e) == 2}
    conn = set()
    first = nb[:tps]
    for t in range(tps):
        for f in range(4):
            q = first[t, f]
            if 0 < q <= tps and (q - 1) // 3 != t // 3:
                conn.add((min(t // 3, (q - 1) // 3), max(t // 3, (q - 1) // 3)))
    assert conn == adj
    # periodic boundary: slice 1 -> last slice carries -1, last -> first +1
    assert (pphi[:tps] == -1).sum() == nt and (pphi[-tps:] == 1).sum() == nt and (pphi[tps:-tps] != 0).sum() == 0
    tp = mesh.tetra_physics
    assert not np.isnan(tp).any() and np.all(tp[:, 4:8] > 0)          # positive cell heights: no degenerate tetrahedra


def test_config4_rk4_strong_field_oracle_vs_device_algorithm(west_mesh):
    mesh, _, st = west_mesh
    om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
    n = 150
    xa, va, wa = workloads.particles_on_triangles(DATA, n, 3)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    ra = om.orbit_timestep_trace(xa, va, wa, 2e-5, ia, ta, fa, 256)
    rb = hm.orbit_timestep(xb, vb, wb, 2e-5, ib, tb, fb, 256)
    assert ia.all() and ra["n_pushes"].sum() > 3000                    # every start point was located
    assert np.array_equal(ra["trace_tetr"], rb["trace_tetr"]) and np.array_equal(ra["trace_face"], rb["trace_face"])
    assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(ta, tb)
    assert np.array_equal(ra["fallback"], rb["fallback"])


@pytest.mark.parametrize("K", [2, 4])
def test_polynomial_pusher_invariants(west_mesh, K):
    mesh, _, st = west_mesh
    s = type(st)(**{**st.__dict__, "ipusher": 2, "poly_order": K, "ispecies": 2})
    om = OracleMesh(mesh, s)
    n = 100
    x, v, w = workloads.particles_on_triangles(DATA, n, 4, energy_ev=3.0e4, mass=2.0 * workloads.AMP)
    b, t, f = workloads.fresh_state(n)
    om.orbit_timestep_batch(x, v, w, 0.0, b, t, f)
    e0, p0, mu0 = om.invariants(x, v, w, t)
    total = om.orbit_timestep_batch(x, v, w, 1e-5, b, t, f, nthreads=4)
    e1, p1, mu1 = om.invariants(x, v, w, t)
    ok = t > 0
    assert total > 2000 and ok.sum() > 30          # start points in the scrape-off layer leave along open field lines
    assert np.abs(mu1 / mu0 - 1)[ok].max() < 1e-13
    assert np.abs(e1 / e0 - 1)[ok].max() < (1e-3 if K == 2 else 1e-6)


@pytest.mark.gpu
def test_gpu_parity_config4(west_mesh, cuda_device):
    from gorilla_b200 import Gorilla
    mesh, _, st = west_mesh
    om, g = OracleMesh(mesh, st), Gorilla(mesh, st)
    n = 400
    xa, va, wa = workloads.particles_on_triangles(DATA, n, 5)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    ra = om.orbit_timestep_trace(xa, va, wa, 2e-5, ia, ta, fa, 128)
    tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, 2e-5, ib, tb, fb, trace_cap=128)
    assert np.array_equal(ra["trace_tetr"], tt) and np.array_equal(ra["trace_face"], tf)
    assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(ta, tb)
    assert np.array_equal(ia, ib) and np.array_equal(fa, fb)
    g.close()
