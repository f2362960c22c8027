"""The .gmesh on-disk mesh format: gorilla_mesh_save / gorilla_mesh_load -- round trip, versioning,
corruption detection, and that a pusher started from a loaded mesh behaves like one started from the built mesh."""
import struct

import numpy as np
import pytest

import workloads
from gorilla_b200 import GorillaError, Mesh, build_mesh, load_mesh
from oracle_binding import OracleMesh

HEADER_BYTES = 8 + 4 * 4 + 2 * 8 + 8 * 4 + 8 * 8 + 8


def test_round_trip_is_bit_identical(small_mesh, tmp_path):
    mesh, _, settings = small_mesh
    path = tmp_path / "tokamak.gmesh"
    mesh.save(path)
    assert path.stat().st_size == HEADER_BYTES + mesh.ntetr * (142 * 8 + 20 * 4) + mesh.verts_rphiz.size * 8
    back = load_mesh(path)
    assert np.array_equal(back.tetra_physics, mesh.tetra_physics) and np.array_equal(back.tetra_grid, mesh.tetra_grid)
    assert np.array_equal(back.verts_rphiz, mesh.verts_rphiz) and back.verts_sthetaphi is None
    assert back.scalars == mesh.scalars
    # same orbits from the loaded mesh (oracle: the file carries everything the hot path reads)
    res = []
    for m in (mesh, back):
        om = OracleMesh(m, settings)
        x, vpar, vperp = workloads.particles_cyl(40, 5)
        s = workloads.fresh_state(40)
        r = om.orbit_timestep_trace(x, vpar, vperp, 1e-5, *s, 32)
        res.append((x, vpar, r["trace_tetr"]))
    assert all(np.array_equal(a, b) for a, b in zip(res[0], res[1]))


def test_arrays_from_elsewhere_can_be_saved(small_mesh, tmp_path):
    """A mesh that did not come from the library's builder (e.g. dumped from a Fortran run): no vertex tables."""
    mesh, _, _ = small_mesh
    m2 = Mesh.from_arrays(mesh.tetra_physics[:500].copy(), mesh.tetra_grid[:500].copy(), **mesh.scalars)
    path = tmp_path / "part.gmesh"
    m2.save(path)
    back = load_mesh(path)
    assert back.ntetr == 500 and back.verts_rphiz is None
    assert np.array_equal(back.tetra_physics, m2.tetra_physics) and np.array_equal(back.tetra_grid, m2.tetra_grid)


def test_bad_files_are_refused(small_mesh, tmp_path):
    mesh, _, _ = small_mesh
    m2 = Mesh.from_arrays(mesh.tetra_physics[:200].copy(), mesh.tetra_grid[:200].copy(), **mesh.scalars)
    good = tmp_path / "good.gmesh"
    m2.save(good)
    raw = bytearray(good.read_bytes())

    def refused(data, what):
        p = tmp_path / "bad.gmesh"
        p.write_bytes(bytes(data))
        with pytest.raises(GorillaError) as ei:
            load_mesh(p)
        assert ei.value.code == 5 and what in str(ei.value), str(ei.value)

    refused(b"not a mesh file at all" * 20, "magic")
    refused(raw[:40], "header")
    v3 = bytearray(raw); v3[8:12] = struct.pack("<I", 3)
    refused(v3, "version 3")
    fl = bytearray(raw); fl[40:44] = struct.pack("<i", 4)
    refused(fl, "flags")
    v1 = bytearray(raw); v1[8:12] = struct.pack("<I", 1)     # version-1 files (no skew records) are still read
    (tmp_path / "v1.gmesh").write_bytes(bytes(v1))
    assert load_mesh(tmp_path / "v1.gmesh").ntetr == 200
    sw = bytearray(raw); sw[12:16] = struct.pack(">I", 0x01020304)
    refused(sw, "byte order")
    rs = bytearray(raw); rs[16:20] = struct.pack("<I", 141)
    refused(rs, "record sizes")
    refused(raw[:-9], "truncated")
    refused(raw + b"x", "trailing")
    flip = bytearray(raw); flip[HEADER_BYTES + 1234] ^= 0x10
    refused(flip, "checksum")
    with pytest.raises(GorillaError):
        load_mesh(tmp_path / "does_not_exist.gmesh")


def test_kind2_mesh_keeps_its_skew_records(product_lib, tmp_path):
    """a mesh built (or dumped) with handover_processing_kind = 2 carries tetra_skew_coord ([ntetr][168],
    tetra_physics_mod.f90:89-99); the file is the complete resume state, so the records travel with it and are covered by the
    checksum."""
    grid, settings = workloads.analytic_tokamak(8, 8, 8)
    settings.handover_processing_kind = 2
    settings.poly_order = 2
    mesh = build_mesh(grid, settings)
    assert mesh.tetra_skew_coord is not None
    path = tmp_path / "skew.gmesh"
    mesh.save(path)
    assert path.stat().st_size == HEADER_BYTES + mesh.ntetr * (142 * 8 + 20 * 4 + 168 * 8) + mesh.verts_rphiz.size * 8
    back = load_mesh(path)
    assert back.tetra_skew_coord is not None and np.array_equal(back.tetra_skew_coord, mesh.tetra_skew_coord)
    assert np.array_equal(back.tetra_physics, mesh.tetra_physics)
    res = []
    for m in (mesh, back):
        om = OracleMesh(m, settings)
        x, vpar, vperp = workloads.particles_cyl(30, 5)
        s = workloads.fresh_state(30)
        r = om.orbit_timestep_trace(x, vpar, vperp, 1e-5, *s, 24)
        res.append((x, vpar, r["trace_tetr"]))
    assert all(np.array_equal(a, b) for a, b in zip(res[0], res[1]))
    raw = bytearray(path.read_bytes())
    raw[-100] ^= 0x01                      # inside the skew payload (last in the file)
    bad = tmp_path / "bad.gmesh"
    bad.write_bytes(bytes(raw))
    with pytest.raises(GorillaError) as ei:
        load_mesh(bad)
    assert "checksum" in str(ei.value)


@pytest.mark.gpu
def test_gpu_pusher_from_a_loaded_mesh(small_mesh, cuda_device, tmp_path):
    from gorilla_b200 import Gorilla
    mesh, _, settings = small_mesh
    path = tmp_path / "m.gmesh"
    mesh.save(path)
    out = []
    for m in (mesh, load_mesh(path)):
        g = Gorilla(m, settings)
        x, vpar, vperp = workloads.particles_cyl(2000, 8)
        s = workloads.fresh_state(2000)
        g.orbit_timestep_gorilla(x, vpar, vperp, 1e-5, *s)
        out.append((x, vpar, vperp, s[1]))
        g.close()
    assert all(np.array_equal(a, b) for a, b in zip(out[0], out[1]))
