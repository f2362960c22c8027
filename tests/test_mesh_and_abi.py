"""Host logic: mesh builder consistency, settings/namelist handling, and the C-ABI surface (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import workloads
from gorilla_b200 import GorillaSettings, api, load_gorilla_inp, load_tetra_grid_inp

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(product_lib):
    header = (ROOT / "include" / "gorilla_b200.h").read_text()
    declared = set(re.findall(r"\b(gorilla_(?:b200|mesh)_[a-z0-9_]+)\s*\(", header))
    assert declared == set(api.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(product_lib, name), name


def test_struct_layouts_match_header_sizes():
    assert C.sizeof(api._Settings) == 8 + 4 * 20 + 8 + 8 + 8 + 4 * 4 + 3 * 8 + 4 * 4
    assert C.sizeof(api._Counters) == 8 * 9 + 16 + 8 + 16
    assert C.sizeof(api._Diag) == 8 * 7 + 8 * 4 + 8 * 2 + 8 * 6 + 8
    assert api._MeshDesc.tetra_physics.offset == 8 and api._MeshDesc.Rmin.offset % 8 == 0


def test_unsupported_settings_are_refused_without_touching_the_gpu(product_lib, small_mesh):
    mesh, _, settings = small_mesh
    for field, value in (("i_precomp", 3),):
        bad = type(settings)(**{**settings.__dict__, field: value})
        with pytest.raises(api.GorillaError) as ei:
            api.Gorilla(mesh, bad)
        assert ei.value.code == 2, field
    with pytest.raises(api.GorillaError) as ei:
        api.Gorilla(mesh, type(settings)(**{**settings.__dict__, "poly_order": 7}))
    assert ei.value.code == 1


def test_unknown_grid_kind_is_refused(product_lib):
    grid, settings = workloads.analytic_tokamak(4, 4, 4)
    grid.grid_kind = 2
    with pytest.raises(api.GorillaError):
        api.build_mesh(grid, settings)


def test_namelist_round_trip(tmp_path):
    (tmp_path / "gorilla.inp").write_text(
        "! comment\n&GORILLANML\n eps_Phi = -1.5d-5 ,\n coord_system = 1 ,\n ispecies = 3 ,\n"
        " boole_periodic_relocation = .false. ,\n ipusher = 2 ,\n poly_order = 4 ,\n boole_guess = .true. ,\n"
        " filename_electric_field = 'electric_field.dat' ,\n/\n")
    s = load_gorilla_inp(tmp_path / "gorilla.inp")
    assert s.eps_Phi == -1.5e-5 and s.coord_system == 1 and s.ispecies == 3 and s.poly_order == 4
    assert s.boole_periodic_relocation is False and s.boole_guess is True
    (tmp_path / "tetra_grid.inp").write_text(
        "&TETRA_GRID_NML\n grid_kind = 5 ,\n n1 = 40 ,\n n2 = 80 ,\n n3 = 40 ,\n R0_analytic_circ = 170.0 ,\n"
        " netcdf_filename = 'MHD_EQUILIBRIA/netcdf_file_for_test.nc' ,\n/\n")
    g = load_tetra_grid_inp(tmp_path / "tetra_grid.inp")
    assert (g.grid_kind, g.n1, g.n2, g.n3) == (5, 40, 80, 40) and g.R0_analytic_circ == 170.0
    assert g.netcdf_filename.endswith("netcdf_file_for_test.nc")
    # blueprint defaults of the reference (SRC/TESTS/test_tetra_grid_settings_mod.f90)
    d = type(g)()
    assert (d.grid_kind, d.n1, d.n2, d.n3, d.sfc_s_min) == (3, 100, 40, 40, 0.1)
    assert GorillaSettings().eps_Phi == 0.0


def test_rect_mesh_topology(small_mesh):
    """Every interior face is shared by exactly two tetrahedra that point at each other, normals oppose
    (check_tetra_overlaps found nothing), periodic flags pair up, volumes tile the torus."""
    mesh, grid, _ = small_mesh
    tg, tp = mesh.tetra_grid, mesh.tetra_physics
    nt = mesh.ntetr
    assert nt == 6 * grid.n1 * grid.n2 * grid.n3
    nb, nf, pp = tg[:, 4:8], tg[:, 8:12], tg[:, 12:16]
    has = nb > 0
    assert (nf[has] >= 1).all() and (nf[~has] == -1).all()
    t_idx, f_idx = np.nonzero(has)
    back_t = nb[nb[t_idx, f_idx] - 1, nf[t_idx, f_idx] - 1]
    back_f = nf[nb[t_idx, f_idx] - 1, nf[t_idx, f_idx] - 1]
    assert (back_t == t_idx + 1).all() and (back_f == f_idx + 1).all()
    assert (pp[t_idx, f_idx] == -pp[nb[t_idx, f_idx] - 1, nf[t_idx, f_idx] - 1]).all()
    assert (pp != 0).sum() == 2 * 2 * grid.n1 * grid.n3
    # boundary faces: 2 triangles per boundary quad on the R and Z faces
    assert (~has).sum() == 2 * 2 * grid.n2 * (grid.n1 + grid.n3)
    an = tp[:, 9:21].reshape(nt, 4, 3)
    dots = np.einsum("ij,ij->i", an[t_idx, f_idx], an[nb[t_idx, f_idx] - 1, nf[t_idx, f_idx] - 1])
    assert (dots < 0).all()
    # dist_ref = 6 * volume (in coordinate space): cells tile [Rmin,Rmax] x [0,2pi] x [Zmin,Zmax]
    np.testing.assert_allclose(tp[:, 3].sum() / 6.0, 100.0 * 2 * np.pi * 100.0, rtol=1e-12)
    assert np.isfinite(tp).all()


def test_linearised_field_matches_analytic_field(small_mesh):
    """|B| and grad|B| of the records reproduce the analytic circular-tokamak field at cell vertices."""
    mesh, grid, _ = small_mesh
    tp = mesh.tetra_physics
    R, Z = tp[:, 0], tp[:, 2]
    rho2 = (R - 170.0) ** 2 + Z ** 2
    q = 1.1 + 2.0 * rho2 / 50.0 ** 2
    Bp, Br, Bz = 2e4 * 170.0 / R, -2e4 * Z / (R * q), 2e4 * (R - 170.0) / (R * q)
    np.testing.assert_allclose(tp[:, 24], np.sqrt(Br ** 2 + Bp ** 2 + Bz ** 2), rtol=1e-13)
    # dt_dtau_const = <R |B|> over the 4 vertices ~ R|B| at the first vertex
    np.testing.assert_allclose(tp[:, 40], R * tp[:, 24], rtol=0.1)
    # curl A = B (contravariant components times sqrt(g) = R): curlA_phi = R * B^phi = Bp
    np.testing.assert_allclose(tp[:, 22], Bp, rtol=0.05)
