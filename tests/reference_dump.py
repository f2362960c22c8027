"""Reader / writer of the reference dump (`gorilla_reference_dump.bin`) and the check that pins the oracle to it.

The dump is written by gorilla_b200/fortran/gorilla_reference_dump.f90, a driver linked against the UNMODIFIED
GORILLA library on a machine that has gfortran (this image has none): the mesh the reference
built (tetra_physics, tetra_grid, vertices, module scalars), the settings of its gorilla.inp, and for a set of
particles the state after one orbit_timestep_gorilla call (SRC/orbit_timestep_gorilla.f90:19-147) plus the
(ind_tetr, iface) pair after each of the first trace_cap pushes.  The layout is documented in the header of the
Fortran file; `write_dump` here produces the same bytes so that the reader and the check are tested without
gfortran (tests/test_reference_dump.py).

Test infrastructure: `check_oracle` runs the C oracle, `check_host_mirror` the device headers compiled for the host,
`check_device` the CUDA path through the C ABI.

    python tests/reference_dump.py particles OUT.bin --n 2000 --trace-cap 64 --t-step 1e-5 --kind cyl|flux ...
    python tests/reference_dump.py check gorilla_reference_dump.bin [--device] [--gmesh OUT.gmesh]
    python tests/reference_dump.py mesh-diff gorilla_reference_dump.bin --tetra-grid-inp tetra_grid.inp --gorilla-inp gorilla.inp
"""
from __future__ import annotations

import argparse
import sys
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

MAGIC = b"GREFDMP1"
_HEAD_INTS = ("ndoubles", "nints", "ntetr", "nvert", "has_sthetaphi", "has_skew", "sign_sqg", "coord_system",
              "n_field_periods", "grid_kind", "grid_size0", "grid_size1", "grid_size2")
_SETTING_INTS = ("ispecies", "boole_periodic_relocation", "ipusher", "boole_pusher_ode45", "boole_dt_dtau",
                 "boole_newton_precalc", "poly_order", "i_precomp", "boole_guess", "i_time_tracing_option",
                 "handover_processing_kind", "boole_adaptive_time_steps", "boole_strong_electric_field",
                 "max_n_intermediate_steps")
_DOUBLES = ("cm_over_e", "particle_mass", "particle_charge", "Rmin", "Rmax", "Zmin", "Zmax", "sfc_s_min", "eps_Phi",
            "desired_delta_energy")


@dataclass
class ReferenceDump:
    head: dict                       # _HEAD_INTS
    settings: dict                   # _SETTING_INTS + eps_Phi, desired_delta_energy
    scalars: dict                    # mesh scalars (gorilla_mesh_desc)
    tetra_physics: np.ndarray        # [ntetr,142] f64
    tetra_grid: np.ndarray           # [ntetr,20] i32
    verts_rphiz: np.ndarray          # [nvert,3]
    verts_sthetaphi: np.ndarray | None
    tetra_skew_coord: np.ndarray | None
    t_step: float = 0.0
    trace_cap: int = 0
    n_steps: int = 1                 # successive orbit_timestep_gorilla calls of t_step each
    inputs: dict = field(default_factory=dict)    # x0 [n,3], vpar0, vperp0
    results: dict = field(default_factory=dict)   # x, vpar, vperp, t_remain, boole_initialized, ind_tetr, iface, n_pushes,
    #                                               trace_ind_tetr [n,cap], trace_iface [n,cap]


class _Cursor:
    def __init__(self, buf: bytes):
        self.buf, self.pos = buf, 0

    def take(self, dtype, count, shape=None):
        nbytes = np.dtype(dtype).itemsize * int(count)
        if self.pos + nbytes > len(self.buf):
            raise ValueError(f"reference dump truncated at byte {self.pos} (wanted {nbytes} more, file has {len(self.buf)})")
        a = np.frombuffer(self.buf, dtype=dtype, count=int(count), offset=self.pos).copy()
        self.pos += nbytes
        return a if shape is None else a.reshape(shape)


def read_dump(path) -> ReferenceDump:
    buf = Path(path).read_bytes()
    if buf[:8] != MAGIC:
        raise ValueError(f"{path}: not a reference dump (magic {buf[:8]!r})")
    c = _Cursor(buf)
    c.pos = 8
    hi = c.take("<i4", len(_HEAD_INTS))
    si = c.take("<i4", len(_SETTING_INTS))
    dd = c.take("<f8", len(_DOUBLES))
    head = dict(zip(_HEAD_INTS, map(int, hi)))
    if head["ndoubles"] != 142 or head["nints"] != 20:
        raise ValueError(f"{path}: record sizes {head['ndoubles']}/{head['nints']} are not those of tetrahedron_physics / "
                         "tetrahedron_grid (142 / 20) this build restates")
    dbl = dict(zip(_DOUBLES, map(float, dd)))
    settings = dict(zip(_SETTING_INTS, map(int, si)))
    settings["eps_Phi"], settings["desired_delta_energy"] = dbl["eps_Phi"], dbl["desired_delta_energy"]
    settings["coord_system"] = head["coord_system"]
    nt, nv = head["ntetr"], head["nvert"]
    scalars = {k: dbl[k] for k in ("cm_over_e", "particle_mass", "particle_charge", "Rmin", "Rmax", "Zmin", "Zmax", "sfc_s_min")}
    scalars.update(sign_sqg=head["sign_sqg"], coord_system=head["coord_system"], n_field_periods=head["n_field_periods"],
                   grid_kind=head["grid_kind"], grid_size=(head["grid_size0"], head["grid_size1"], head["grid_size2"]))
    tp = c.take("<f8", nt * 142, (nt, 142))
    tg = c.take("<i4", nt * 20, (nt, 20))
    vr = c.take("<f8", nv * 3, (nv, 3))
    vs = c.take("<f8", nv * 3, (nv, 3)) if head["has_sthetaphi"] else None
    sk = c.take("<f8", nt * 168, (nt, 168)) if head["has_skew"] else None
    n, cap, n_steps = map(int, c.take("<i4", 3))
    t_step = float(c.take("<f8", 1)[0])
    inputs = dict(x0=c.take("<f8", n * 3, (n, 3)), vpar0=c.take("<f8", n), vperp0=c.take("<f8", n))
    results = dict(x=c.take("<f8", n * 3, (n, 3)), vpar=c.take("<f8", n), vperp=c.take("<f8", n), t_remain=c.take("<f8", n))
    for k in ("boole_initialized", "ind_tetr", "iface", "n_pushes"):
        results[k] = c.take("<i4", n)
    results["trace_ind_tetr"] = c.take("<i4", n * cap, (n, cap))
    results["trace_iface"] = c.take("<i4", n * cap, (n, cap))
    if c.pos != len(buf):
        raise ValueError(f"{path}: {len(buf) - c.pos} trailing bytes after the last array")
    return ReferenceDump(head, settings, scalars, tp, tg, vr, vs, sk, t_step, cap, n_steps, inputs, results)


def write_dump(path, mesh, settings, t_step, trace_cap, inputs, results, n_steps=1) -> None:
    """The bytes gorilla_reference_dump.f90 writes, from a gorilla_b200.Mesh + GorillaSettings + result arrays."""
    s = mesh.scalars
    nt = mesh.ntetr
    vr = mesh.verts_rphiz if mesh.verts_rphiz is not None else np.zeros((0, 3))
    has_s = int(mesh.verts_sthetaphi is not None and len(vr) > 0)
    has_k = int(settings.handover_processing_kind == 2)
    head = [142, 20, nt, len(vr), has_s, has_k, s["sign_sqg"], s["coord_system"], s["n_field_periods"], s["grid_kind"], *s["grid_size"]]
    sett = [int(getattr(settings, k)) for k in _SETTING_INTS]
    dbl = [float(s.get(k, 0.0)) for k in _DOUBLES[:8]] + [settings.eps_Phi, settings.desired_delta_energy]
    n = inputs["x0"].shape[0]
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(np.asarray(head, "<i4").tobytes())
        f.write(np.asarray(sett, "<i4").tobytes())
        f.write(np.asarray(dbl, "<f8").tobytes())
        f.write(np.ascontiguousarray(mesh.tetra_physics, "<f8").tobytes())
        f.write(np.ascontiguousarray(mesh.tetra_grid, "<i4").tobytes())
        f.write(np.ascontiguousarray(vr, "<f8").tobytes())
        if has_s:
            f.write(np.ascontiguousarray(mesh.verts_sthetaphi, "<f8").tobytes())
        if has_k:
            f.write(np.ascontiguousarray(mesh.tetra_skew_coord, "<f8").tobytes())
        f.write(np.asarray([n, trace_cap, n_steps], "<i4").tobytes())
        f.write(np.asarray([t_step], "<f8").tobytes())
        for k in ("x0", "vpar0", "vperp0"):
            f.write(np.ascontiguousarray(inputs[k], "<f8").tobytes())
        for k in ("x", "vpar", "vperp", "t_remain"):
            f.write(np.ascontiguousarray(results[k], "<f8").tobytes())
        for k in ("boole_initialized", "ind_tetr", "iface", "n_pushes", "trace_ind_tetr", "trace_iface"):
            f.write(np.ascontiguousarray(results[k], "<i4").tobytes())


def write_particles(path, x0, vpar0, vperp0, t_step, trace_cap, n_steps=1) -> None:
    """dump_particles.bin, the input of gorilla_reference_dump.f90: int32 n, trace_cap, n_steps; f64 t_step; x0[n][3], vpar0,
    vperp0."""
    assert trace_cap >= 1 and n_steps >= 1
    with open(path, "wb") as f:
        f.write(np.asarray([x0.shape[0], trace_cap, n_steps], "<i4").tobytes())
        f.write(np.asarray([t_step], "<f8").tobytes())
        for a in (x0, vpar0, vperp0):
            f.write(np.ascontiguousarray(a, "<f8").tobytes())


def mesh_and_settings(d: ReferenceDump):
    """gorilla_b200.Mesh (wrapping the dumped arrays) and GorillaSettings of a dump."""
    from gorilla_b200 import GorillaSettings
    from gorilla_b200.api import Mesh
    m = Mesh.from_arrays(d.tetra_physics, d.tetra_grid, **d.scalars)
    m.verts_rphiz, m.verts_sthetaphi, m.tetra_skew_coord = d.verts_rphiz, d.verts_sthetaphi, d.tetra_skew_coord
    st = GorillaSettings()
    for k, v in d.settings.items():
        cur = getattr(st, k)
        setattr(st, k, bool(v) if isinstance(cur, bool) else type(cur)(v))
    return m, st


_STATE_KEYS = ("x", "vpar", "vperp", "t_remain", "boole_initialized", "ind_tetr", "iface", "n_pushes", "trace_ind_tetr",
               "trace_iface")


def _fresh(d: ReferenceDump):
    n = d.inputs["x0"].shape[0]
    return (d.inputs["x0"].copy(), d.inputs["vpar0"].copy(), d.inputs["vperp0"].copy(),
            np.zeros(n, np.int32), np.full(n, -1, np.int32), np.full(n, -1, np.int32))


def _compare(d: ReferenceDump, got: dict) -> dict:
    """Per-key count of particles whose value differs from the dump (bit comparison; NaN == NaN)."""
    bad = {}
    for k in _STATE_KEYS:
        a, b = d.results[k], got[k]
        neq = ~((a == b) | ((a != a) & (b != b))) if a.dtype.kind == "f" else (a != b)
        if k == "t_remain":
            # a particle find_tetra could not place returns before t_remain_out is assigned (orbit_timestep_gorilla.f90:57-59):
            # the value is undefined in the reference (the dump program writes 0 there), so it is not compared
            neq = neq & (d.results["boole_initialized"] != 0)
        cnt = int(np.count_nonzero(neq.reshape(neq.shape[0], -1).any(axis=1))) if neq.size else 0
        if cnt:
            bad[k] = cnt
    return bad


def _run_steps(d: ReferenceDump, step) -> dict:
    """d.n_steps successive orbit_timestep_gorilla calls, as the dump program makes them.  `step(x, vpar, vperp, binit, ind, ifc,
    cap)` performs ONE batched call in place and returns (t_remain [m], n_pushes [m], trace_ind_tetr [m,cap], trace_iface [m,cap]).
    A particle that was not placed by find_tetra or has left the domain is not passed on to the next call (the reference would
    index tetra_physics(-1)); n_pushes adds up over the calls, the traces hold the first trace_cap pushes of all calls together,
    t_remain is that of the last call made for the particle."""
    x, vpar, vperp, binit, ind, ifc = _fresh(d)
    n, cap = x.shape[0], max(d.trace_cap, 1)
    t_rem, npush = np.zeros(n), np.zeros(n, np.int64)
    tt, tf = np.zeros((n, cap), np.int32), np.zeros((n, cap), np.int32)
    act = np.arange(n)
    for _ in range(d.n_steps):
        if act.size == 0:
            break
        sub = [np.ascontiguousarray(a[act]) for a in (x, vpar, vperp, binit, ind, ifc)]
        tr, np1, t1, f1 = step(*sub, cap)
        for a, b in zip((x, vpar, vperp, binit, ind, ifc), sub):
            a[act] = b
        t_rem[act] = tr
        before = npush[act]
        for j in np.nonzero((before < cap) & (np1 > 0))[0]:
            m = int(min(cap - before[j], np1[j]))
            tt[act[j], before[j]:before[j] + m] = t1[j, :m]
            tf[act[j], before[j]:before[j] + m] = f1[j, :m]
        npush[act] += np1
        act = act[(sub[4] != -1) & (sub[3] != 0)]
    return dict(x=x, vpar=vpar, vperp=vperp, t_remain=t_rem, boole_initialized=binit, ind_tetr=ind, iface=ifc,
                n_pushes=npush.astype(np.int32), trace_ind_tetr=tt, trace_iface=tf)


def run_oracle(d: ReferenceDump) -> dict:
    from oracle_binding import OracleMesh
    mesh, st = mesh_and_settings(d)
    om = OracleMesh(mesh, st)

    def step(x, vpar, vperp, binit, ind, ifc, cap):
        r = om.orbit_timestep_trace(x, vpar, vperp, d.t_step, binit, ind, ifc, cap)
        return r["t_remain"], r["n_pushes"], r["trace_tetr"], r["trace_face"]
    return _run_steps(d, step)


def run_device(d: ReferenceDump) -> dict:
    from gorilla_b200 import Gorilla
    mesh, st = mesh_and_settings(d)
    g = Gorilla(mesh, st)

    def step(x, vpar, vperp, binit, ind, ifc, cap):
        t_rem, npush = np.zeros(x.shape[0]), np.zeros(x.shape[0], np.int64)
        tt, tf = g.orbit_timestep_gorilla(x, vpar, vperp, d.t_step, binit, ind, ifc, t_remain_out=t_rem, n_pushes=npush,
                                          trace_cap=cap)
        return t_rem, npush, tt, tf
    try:
        return _run_steps(d, step)
    finally:
        g.close()


def run_host_mirror(d: ReferenceDump) -> dict:
    """The device headers compiled for the host (tests/host_mirror): the kernels' algorithm without a GPU."""
    from host_mirror_binding import HostMirror
    mesh, st = mesh_and_settings(d)
    hm = HostMirror(mesh, st)

    def step(x, vpar, vperp, binit, ind, ifc, cap):
        r = hm.orbit_timestep(x, vpar, vperp, d.t_step, binit, ind, ifc, cap)
        return r["t_remain"], r["n_pushes"], r["trace_tetr"], r["trace_face"]
    return _run_steps(d, step)


def check_host_mirror(d: ReferenceDump) -> dict:
    return _compare(d, run_host_mirror(d))


def check_oracle(d: ReferenceDump) -> dict:
    return _compare(d, run_oracle(d))


def check_device(d: ReferenceDump) -> dict:
    return _compare(d, run_device(d))


# first column of the named members of type tetrahedron_physics (tetra_physics_mod.f90:9-83) inside the 142-double record, as
# the repack reads them (gorilla_b200/csrc/gb_repack.hpp); a range ends where the next named member starts
_TP_OFFSETS = dict(x1=0, dist_ref=3, tetra_dist_ref=8, anorm=9, curlA=21, bmod1=24, Aphi1=26, h1_1=27, h2_1=28, h3_1=29, Phi1=30,
                   R1=31, vE2_1=34, v2Emod_1=36, Er_mod=37, v_E_mod_average=38, dt_dtau_const=40, gBxcurlA=41, gPhixcurlA=42,
                   gv2EmodxcurlA=43, gBxcurlvE=44, gPhixcurlvE=45, gv2EmodxcurlvE=46, spalpmat=47, spbetmat=48, spgammat=49,
                   gBxh1=50, gPhixh1=53, gv2Emodxh1=56, gB=59, gPhi=62, gAphi=77, gh1=80, gh2=83, gh3=86, curlh=89, gvE2=95,
                   curlvE=101, gv2Emod=104, alpmat=107, betmat=116, gammat=125)


def mesh_diff(d: ReferenceDump, mesh) -> dict:
    """Deviation of a mesh built by this library's host builders from the mesh the reference built (same namelists).
    The host builders restate ~6 k lines of spline / mesh Fortran; bit equality with the Fortran is not claimed for them
    (SURVEY.md H6), so this reports, per member of tetrahedron_physics, the largest deviation relative to the member's
    largest magnitude, and for tetra_grid (vertex indices, neighbours, faces, periodic-boundary flags) the share of
    identical records.  The push itself is pinned on the DUMPED mesh (check_oracle / check_device)."""
    out = {"ntetr_reference": int(d.head["ntetr"]), "ntetr_built": int(mesh.ntetr)}
    if d.head["ntetr"] != mesh.ntetr:
        return out
    out["tetra_grid_identical_records"] = float(np.mean(np.all(d.tetra_grid == mesh.tetra_grid, axis=1)))
    names = sorted(_TP_OFFSETS, key=_TP_OFFSETS.get)
    dev = {}
    for k, name in enumerate(names):
        lo = _TP_OFFSETS[name]
        hi = _TP_OFFSETS[names[k + 1]] if k + 1 < len(names) else 134
        a, b = d.tetra_physics[:, lo:hi], mesh.tetra_physics[:, lo:hi]
        scale = float(np.max(np.abs(a)))
        dev[name] = 0.0 if scale == 0.0 and not np.any(b) else float(np.max(np.abs(a - b)) / (scale or 1.0))
    out["tetra_physics_max_dev_rel_to_member_scale"] = dev
    out["tetra_physics_identical_records"] = float(np.mean(np.all(d.tetra_physics == mesh.tetra_physics, axis=1)))
    for k in ("cm_over_e", "particle_mass", "particle_charge", "sign_sqg", "n_field_periods"):
        out[k + "_equal"] = bool(d.scalars[k] == mesh.scalars[k])
    return out


def _cli():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("particles", help="write dump_particles.bin for gorilla_reference_dump.x")
    p.add_argument("out")
    p.add_argument("--n", type=int, default=2000)
    p.add_argument("--trace-cap", type=int, default=64)
    p.add_argument("--t-step", type=float, default=1e-5)
    p.add_argument("--n-steps", type=int, default=3, help="successive orbit_timestep_gorilla calls per particle")
    p.add_argument("--seed", type=int, default=2024)
    p.add_argument("--kind", choices=["cyl", "flux"], default="cyl",
                   help="cyl: (R,phi,Z) around --R0/--a (coord_system 1); flux: (s,theta,phi) with s in [0.2,0.9] (coord_system 2)")
    p.add_argument("--R0", type=float, default=170.0)
    p.add_argument("--a", type=float, default=50.0)
    p.add_argument("--nfp", type=int, default=1)
    p.add_argument("--energy-ev", type=float, default=3.0e3)
    p.add_argument("--mass-amu", type=float, default=2.0)
    c = sub.add_parser("check", help="compare the oracle (and the CUDA path) with a dump")
    c.add_argument("dump")
    c.add_argument("--device", action="store_true")
    c.add_argument("--gmesh", default="", help="also save the dumped mesh as a .gmesh file")
    m = sub.add_parser("mesh-diff", help="build the mesh of the same namelists with this library and compare it with the dumped one")
    m.add_argument("dump")
    m.add_argument("--tetra-grid-inp", required=True)
    m.add_argument("--gorilla-inp", required=True)
    a = ap.parse_args()
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    if a.cmd == "particles":
        rng = np.random.Generator(np.random.PCG64(a.seed))
        x = np.empty((a.n, 3))
        if a.kind == "cyl":
            rho, th = a.a * (0.1 + 0.75 * rng.random(a.n)), 2 * np.pi * rng.random(a.n)
            x[:, 0], x[:, 1], x[:, 2] = a.R0 + rho * np.cos(th), 2 * np.pi * rng.random(a.n), rho * np.sin(th)
        else:
            x[:, 0], x[:, 1], x[:, 2] = 0.2 + 0.7 * rng.random(a.n), 2 * np.pi * rng.random(a.n), 2 * np.pi / a.nfp * rng.random(a.n)
        vmod = np.sqrt(2.0 * a.energy_ev * 1.6022e-12 / (a.mass_amu * 1.6726e-24))
        vpar = (2.0 * rng.random(a.n) - 1.0) * vmod
        write_particles(a.out, x, vpar, np.sqrt(vmod ** 2 - vpar ** 2), a.t_step, a.trace_cap, a.n_steps)
        print(f"wrote {a.out}: {a.n} particles, trace_cap {a.trace_cap}, {a.n_steps} steps of {a.t_step} s")
        return 0
    if a.cmd == "mesh-diff":
        import json
        from gorilla_b200 import build_mesh, load_gorilla_inp, load_tetra_grid_inp
        built = build_mesh(load_tetra_grid_inp(a.tetra_grid_inp), load_gorilla_inp(a.gorilla_inp))
        print(json.dumps(mesh_diff(read_dump(a.dump), built), indent=1))
        return 0
    d = read_dump(a.dump)
    print(f"{a.dump}: ntetr {d.head['ntetr']}, grid_kind {d.head['grid_kind']}, ipusher {d.settings['ipusher']}, "
          f"poly_order {d.settings['poly_order']}, {d.inputs['x0'].shape[0]} particles x {d.n_steps} steps, "
          f"{int(d.results['n_pushes'].sum())} pushes")
    if a.gmesh:
        mesh_and_settings(d)[0].save(a.gmesh)
    bad = check_oracle(d)
    print("oracle vs reference:", "bit-identical" if not bad else f"DIFFERS {bad}")
    bh = check_host_mirror(d)
    print("device headers on the host vs reference:", "bit-identical" if not bh else f"DIFFERS {bh}")
    bad.update({"host_mirror_" + k: v for k, v in bh.items()})
    if a.device:
        bd = check_device(d)
        print("CUDA path vs reference:", "bit-identical" if not bd else f"DIFFERS {bd}")
        bad.update(bd)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(_cli())
