"""The reference dump pipeline (SURVEY.md section 7 step 9 / 8c pin 5): gorilla_b200/fortran/gorilla_reference_dump.f90
writes the mesh and per-push traces of the UNMODIFIED gfortran build; tests/reference_dump.py reads them and checks the
oracle and the CUDA path against them.

No gfortran exists in this image, so no real dump is committed yet ("parity unpinned").  What is tested here:
the file format (a Python writer of the same byte layout stands in for the Fortran program), the mesh hand-over of a
dumped mesh into `Mesh.from_arrays` / .gmesh, the comparison, and that a tampered dump is detected.  Any real dump dropped
into tests/golden/reference_dumps/*.bin is picked up by `test_committed_reference_dumps_*` and pins the oracle (CPU) and
the kernels (GPU) against the Fortran binary.
"""
import dataclasses
from pathlib import Path

import numpy as np
import pytest

import reference_dump as rd
import workloads

DUMP_DIR = Path(__file__).resolve().parent / "golden" / "reference_dumps"
REAL_DUMPS = sorted(DUMP_DIR.glob("*.bin")) if DUMP_DIR.is_dir() else []


def _synthetic_dump(path, mesh, settings, n=120, seed=11, t_step=1.5e-5, cap=48, n_steps=1):
    """A dump in the Fortran program's layout whose results come from the oracle (stand-in for the gfortran run)."""
    x0, vpar0, vperp0 = workloads.particles_cyl(n, seed)
    x0[0, 0] = 1.0e4   # one particle outside the grid: find_tetra gives ind_tetr = -1, state untouched
    inputs = dict(x0=x0, vpar0=vpar0, vperp0=vperp0)
    skeleton = rd.ReferenceDump({}, {k: getattr(settings, k) for k in rd._SETTING_INTS + ("eps_Phi", "desired_delta_energy", "coord_system")},
                                dict(mesh.scalars), mesh.tetra_physics, mesh.tetra_grid, mesh.verts_rphiz, mesh.verts_sthetaphi,
                                mesh.tetra_skew_coord, t_step, cap, n_steps, inputs, {})
    results = rd.run_oracle(skeleton)
    rd.write_dump(path, mesh, settings, t_step, cap, inputs, results, n_steps)
    return results


@pytest.mark.parametrize("variant", ["poly2", "poly4_phi", "rk4", "handover2"])
def test_dump_format_round_trip_and_oracle_check(tmp_path, small_mesh, small_mesh_phi, variant):
    mesh, grid, settings = small_mesh_phi if variant == "poly4_phi" else small_mesh
    st = dataclasses.replace(settings, **dict(poly2=dict(poly_order=2), poly4_phi=dict(poly_order=4), rk4=dict(ipusher=1),
                                              handover2=dict(poly_order=2, handover_processing_kind=2))[variant])
    if variant == "handover2":
        from gorilla_b200 import build_mesh
        mesh = build_mesh(grid, st)   # builds the tetra_skew_coord records
        assert mesh.tetra_skew_coord is not None
    p = tmp_path / "gorilla_reference_dump.bin"
    res = _synthetic_dump(p, mesh, st)
    d = rd.read_dump(p)
    # header, settings and mesh survive byte for byte
    assert d.head["ntetr"] == mesh.ntetr and d.head["grid_kind"] == 5 and d.head["has_skew"] == int(variant == "handover2")
    assert d.settings["ipusher"] == st.ipusher and d.settings["poly_order"] == st.poly_order and d.settings["eps_Phi"] == st.eps_Phi
    assert np.array_equal(d.tetra_physics, mesh.tetra_physics) and np.array_equal(d.tetra_grid, mesh.tetra_grid)
    assert np.array_equal(d.verts_rphiz, mesh.verts_rphiz)
    m2, st2 = rd.mesh_and_settings(d)
    assert st2 == st and m2.scalars["grid_size"] == tuple(mesh.scalars["grid_size"])
    assert np.array_equal(d.results["trace_ind_tetr"], res["trace_ind_tetr"])
    assert d.results["ind_tetr"][0] == -1 and d.results["boole_initialized"][0] == 0 and d.results["n_pushes"][0] == 0
    assert d.results["n_pushes"].sum() > 1000
    # the check: oracle on the dumped mesh reproduces the dumped results
    assert rd.check_oracle(d) == {}
    assert rd.check_host_mirror(d) == {}   # the kernels' algorithm (device headers compiled for the host)
    # ... and a dump whose reference results differ in one bit / one cell is caught
    d.results["vpar"][5] = np.nextafter(d.results["vpar"][5], np.inf)
    d.results["trace_ind_tetr"][7, 3] += 1
    bad = rd.check_oracle(d)
    assert bad == {"vpar": 1, "trace_ind_tetr": 1}


def test_multi_step_dump(tmp_path, small_mesh):
    """n_steps successive calls: the first locates the particle, the later ones start from (ind_tetr, iface); pushes add up,
    the traces run on across the calls, particles that left the domain are not pushed again."""
    from oracle_binding import OracleMesh
    mesh, _, settings = small_mesh
    st = dataclasses.replace(settings, poly_order=3)
    p = tmp_path / "d.bin"
    n, t_step, cap = 150, 4e-6, 40
    _synthetic_dump(p, mesh, st, n=n, t_step=t_step, cap=cap, n_steps=3, seed=5)
    d = rd.read_dump(p)
    assert d.n_steps == 3
    assert rd.check_oracle(d) == {} and rd.check_host_mirror(d) == {}
    # independent of _run_steps: three plain batched calls on all particles that stay inside
    x, vpar, vperp = d.inputs["x0"].copy(), d.inputs["vpar0"].copy(), d.inputs["vperp0"].copy()
    binit, ind, ifc = workloads.fresh_state(n)
    om = OracleMesh(mesh, st)
    total = np.zeros(n, np.int64)
    first = None
    for k in range(3):
        r = om.orbit_timestep_trace(x, vpar, vperp, t_step, binit, ind, ifc, cap)
        total += r["n_pushes"]
        first = r if first is None else first
    inside = (ind > 0) & (d.results["ind_tetr"] > 0)
    assert inside.sum() > n // 2
    assert np.array_equal(d.results["x"][inside], x[inside]) and np.array_equal(d.results["vpar"][inside], vpar[inside])
    assert np.array_equal(d.results["n_pushes"][inside], total[inside])
    # the trace continues over the call boundary: a particle with fewer than cap pushes in call 1 has later slots filled by call 2
    short = inside & (first["n_pushes"] < cap) & (total > first["n_pushes"])
    assert short.any()
    j = int(np.nonzero(short)[0][0])
    k1 = int(first["n_pushes"][j])
    assert np.array_equal(d.results["trace_ind_tetr"][j, :k1], first["trace_tetr"][j, :k1]) and d.results["trace_ind_tetr"][j, k1] > 0


def test_dumped_mesh_to_gmesh(tmp_path, small_mesh):
    """A dumped mesh becomes a .gmesh (built once, loaded by later runs) with identical arrays."""
    from gorilla_b200 import api
    mesh, _, settings = small_mesh
    p = tmp_path / "d.bin"
    _synthetic_dump(p, mesh, dataclasses.replace(settings, poly_order=2), n=8, cap=4)
    d = rd.read_dump(p)
    m, _ = rd.mesh_and_settings(d)
    m.save(tmp_path / "d.gmesh")
    m2 = api.load_mesh(tmp_path / "d.gmesh")
    assert np.array_equal(m2.tetra_physics, mesh.tetra_physics) and np.array_equal(m2.tetra_grid, mesh.tetra_grid)
    assert m2.scalars["cm_over_e"] == mesh.scalars["cm_over_e"] and m2.scalars["grid_kind"] == 5


def test_mesh_diff_of_builder_against_dump(tmp_path, small_mesh):
    """mesh-diff: a mesh built by the host builders against a dumped one -- identical here (same builder), and a perturbed
    member / a changed neighbour entry show up under the right name."""
    mesh, _, settings = small_mesh
    p = tmp_path / "d.bin"
    _synthetic_dump(p, mesh, dataclasses.replace(settings, poly_order=2), n=8, cap=4)
    d = rd.read_dump(p)
    r = rd.mesh_diff(d, mesh)
    assert r["tetra_grid_identical_records"] == 1.0 and r["tetra_physics_identical_records"] == 1.0
    assert set(r["tetra_physics_max_dev_rel_to_member_scale"].values()) == {0.0} and r["cm_over_e_equal"]
    d.tetra_physics[17, 24] *= 1.0 + 1e-9     # bmod1 of one tetrahedron
    d.tetra_grid[5, 7] += 1
    r = rd.mesh_diff(d, mesh)
    dev = r["tetra_physics_max_dev_rel_to_member_scale"]
    assert 1e-10 < dev["bmod1"] < 1e-8 and all(v == 0.0 for k, v in dev.items() if k != "bmod1")
    assert r["tetra_grid_identical_records"] == 1.0 - 1.0 / mesh.ntetr


def test_bad_dumps_are_rejected(tmp_path, small_mesh):
    mesh, _, settings = small_mesh
    p = tmp_path / "d.bin"
    _synthetic_dump(p, mesh, dataclasses.replace(settings, poly_order=2), n=8, cap=4)
    raw = p.read_bytes()
    (tmp_path / "magic.bin").write_bytes(b"GREFDMP9" + raw[8:])
    with pytest.raises(ValueError, match="not a reference dump"):
        rd.read_dump(tmp_path / "magic.bin")
    (tmp_path / "short.bin").write_bytes(raw[:-5])
    with pytest.raises(ValueError, match="truncated"):
        rd.read_dump(tmp_path / "short.bin")
    (tmp_path / "long.bin").write_bytes(raw + b"\0\0")
    with pytest.raises(ValueError, match="trailing"):
        rd.read_dump(tmp_path / "long.bin")
    bad_rec = bytearray(raw)
    bad_rec[8:12] = np.asarray([141], "<i4").tobytes()
    (tmp_path / "rec.bin").write_bytes(bytes(bad_rec))
    with pytest.raises(ValueError, match="record sizes"):
        rd.read_dump(tmp_path / "rec.bin")


def test_particle_file_layout(tmp_path):
    """dump_particles.bin as the Fortran program reads it: int32 n, cap, n_steps; f64 t_step; x0[n][3], vpar0[n], vperp0[n]."""
    x, vpar, vperp = workloads.particles_cyl(5, 1)
    rd.write_particles(tmp_path / "p.bin", x, vpar, vperp, 2e-5, 16, 3)
    raw = (tmp_path / "p.bin").read_bytes()
    assert len(raw) == 12 + 8 + 5 * 5 * 8
    assert tuple(np.frombuffer(raw, "<i4", 3)) == (5, 16, 3) and np.frombuffer(raw, "<f8", 1, 12)[0] == 2e-5
    assert np.array_equal(np.frombuffer(raw, "<f8", 15, 20).reshape(5, 3), x)
    assert np.array_equal(np.frombuffer(raw, "<f8", 5, 20 + 15 * 8 + 40), vperp)


@pytest.mark.skipif(not REAL_DUMPS, reason="no dump of the gfortran build committed yet: parity unpinned (DESIGN.md section 4)")
@pytest.mark.parametrize("path", REAL_DUMPS, ids=lambda p: p.name)
def test_committed_reference_dumps_pin_the_oracle(path, product_lib, oracle_lib, host_mirror_lib):
    d = rd.read_dump(path)
    assert rd.check_oracle(d) == {}
    assert rd.check_host_mirror(d) == {}


@pytest.mark.gpu
def test_device_reproduces_a_dump(tmp_path, small_mesh, cuda_device):
    mesh, _, settings = small_mesh
    for k, st in enumerate((dataclasses.replace(settings, poly_order=2), dataclasses.replace(settings, poly_order=4),
                            dataclasses.replace(settings, ipusher=1))):
        p = tmp_path / f"d{k}.bin"
        _synthetic_dump(p, mesh, st, n=400, cap=64)
        assert rd.check_device(rd.read_dump(p)) == {}


@pytest.mark.gpu
@pytest.mark.skipif(not REAL_DUMPS, reason="no dump of the gfortran build committed yet")
@pytest.mark.parametrize("path", REAL_DUMPS, ids=lambda p: p.name)
def test_committed_reference_dumps_pin_the_device(path, cuda_device):
    assert rd.check_device(rd.read_dump(path)) == {}
