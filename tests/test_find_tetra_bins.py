"""find_tetra on the slice-wise grids (kinds 2, 3, 4) uses 2-D bins built from the face planes of the records instead of
scanning the whole phi slice (gb_repack.hpp build_find_bins, gb_find.cuh).  The result must be the reference's: the FIRST
tetrahedron of the scan order that contains the point, including starts on slice boundaries, on cell faces and at the
periodic seams.  The oracle scans like the reference; the device algorithm (host mirror here, CUDA in the gpu test) bins."""
from pathlib import Path

import numpy as np
import pytest

import workloads
from gorilla_b200 import build_mesh
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh

DATA = Path(__file__).resolve().parent.parent / "data" / "equilibria"


def _points_flux(n, seed, n2, n3, nfp):
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.column_stack([0.1 + 0.9 * rng.random(n), 2 * np.pi * rng.random(n), 2 * np.pi / nfp * rng.random(n)])
    hphi, hth = 2 * np.pi / nfp / n2, 2 * np.pi / n3
    x[0::7, 2] = hphi * np.floor(x[0::7, 2] / hphi)              # exactly on a phi slice boundary
    x[1::7, 1] = hth * np.floor(x[1::7, 1] / hth)                # exactly on a theta grid line
    x[2::7, 1] = 0.0                                             # periodic seam in theta
    x[3::7, 2] = 0.0                                             # periodic seam in phi
    x[4::7, 0] = 0.1 + 0.9 * np.round((x[4::7, 0] - 0.1) / 0.9 * 8) / 8   # on flux-surface grid rings (n1 = 8)
    x[5::7, 0] = 1.0 - 1e-14                                     # at the outer boundary
    return x


def _check(mesh, st, x):
    n = len(x)
    rng = np.random.Generator(np.random.PCG64(5))
    lam = 2 * rng.random(n) - 1
    vmod = 5.0e7
    v, w = lam * vmod, vmod * np.sqrt(1 - lam ** 2)
    xa, xb = x.copy(), x.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    OracleMesh(mesh, st).orbit_timestep_batch(xa, v.copy(), w.copy(), 0.0, ia, ta, fa)
    HostMirror(mesh, st).orbit_timestep(xb, v.copy(), w.copy(), 0.0, ib, tb, fb, 0)
    assert np.array_equal(ta, tb) and np.array_equal(fa, fb) and np.array_equal(ia, ib) and np.array_equal(xa, xb)
    return ta, fa


def test_vmec_field_aligned(product_lib):
    grid, st = workloads.vmec_qi(str(DATA / "netcdf_file_for_test.nc"), n1=8, n2=6, n3=12)
    mesh = build_mesh(grid, st)
    t, f = _check(mesh, st, _points_flux(4000, 1, 6, 12, 5))
    assert (t > 0).mean() > 0.85 and (f > 0).sum() > 300


def test_efit_field_aligned(product_lib):
    grid, st = workloads.efit_flux(DATA, n1=8, n2=6, n3=12)
    mesh = build_mesh(grid, st)
    t, f = _check(mesh, st, _points_flux(4000, 2, 6, 12, 1))
    assert (t > 0).mean() > 0.85 and (f > 0).sum() > 300


def test_soledge3x(product_lib):
    grid, st = workloads.west_soledge3x(DATA, n2=5)
    mesh = build_mesh(grid, st)
    x, _, _ = workloads.particles_on_triangles(DATA, 3000, 3)
    knots = np.loadtxt(DATA / "MESH_SOLEDGE3X_EIRENE" / "knots_for_test.dat", skiprows=1)
    x[0::5, 1] = 2 * np.pi / 5 * np.floor(x[0::5, 1] / (2 * np.pi / 5))          # on a slice boundary
    x[1::5, 0], x[1::5, 2] = knots[: len(x[1::5]), 0], knots[: len(x[1::5]), 1]  # exactly on mesh knots
    t, f = _check(mesh, st, x)
    assert (t > 0).mean() > 0.9


@pytest.mark.gpu
def test_gpu_binned_equals_full_scan_and_oracle(cuda_device, product_lib):
    from gorilla_b200 import Gorilla
    grid, st = workloads.efit_flux(DATA, n1=8, n2=6, n3=12)
    mesh = build_mesh(grid, st)
    x = _points_flux(20000, 3, 6, 12, 1)
    v = np.full(len(x), 3.0e7)
    w = np.full(len(x), 4.0e7)
    g = Gorilla(mesh, st)
    xb = x.copy()
    tb, fb = g.find_tetra(xb, v, w)
    g._debug_find_bins(False)
    xs = x.copy()
    ts, fs = g.find_tetra(xs, v, w)
    g.close()
    assert np.array_equal(tb, ts) and np.array_equal(fb, fs) and np.array_equal(xb, xs)
    sub = slice(0, 3000)
    xo = x[sub].copy()
    to, fo = OracleMesh(mesh, st).find_tetra(xo, v[sub], w[sub])
    assert np.array_equal(to, tb[sub]) and np.array_equal(fo, fb[sub])
