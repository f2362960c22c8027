"""Host-side mirror of the reference's public interface for the orbit-pusher hot path.

Reference interface (SRC/orbit_timestep_gorilla.f90:10):
    use orbit_timestep_gorilla_mod, only: initialize_gorilla, orbit_timestep_gorilla, check_coordinate_domain
    call orbit_timestep_gorilla(x,vpar,vperp,t_step,boole_initialized,ind_tetr,iface,t_remain_out)

Same names, argument meaning, 1-based indices and in-band loss signalling (ind_tetr = -1) here; the
arrays are batched (x is [n,3]).  Everything below the ctypes boundary is CUDA: a missing
libgorilla_b200.so or a missing GPU raises, nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from .settings import GorillaSettings, TetraGridSettings

import os

# GORILLA_B200_LIB selects an alternative build of the same library (tuning experiments); default = in-tree build
_LIB_PATH = Path(os.environ.get("GORILLA_B200_LIB", Path(__file__).resolve().parent / "lib" / "libgorilla_b200.so"))
_lib = None


class GorillaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gorilla_b200 error {code}: {msg}")
        self.code = code


class _Settings(C.Structure):
    _fields_ = [("eps_Phi", C.c_double)] + [(n, C.c_int32) for n in (
        "coord_system", "ispecies", "boole_periodic_relocation", "ipusher", "boole_pusher_ode45", "boole_dt_dtau",
        "boole_newton_precalc", "poly_order", "i_precomp", "boole_guess", "i_time_tracing_option",
        "handover_processing_kind", "boole_adaptive_time_steps", "boole_strong_electric_field",
        "boole_grid_for_find_tetra", "boole_time_Hamiltonian", "boole_gyrophase", "boole_vpar_int",
        "boole_vpar2_int", "max_n_intermediate_steps")] + [("desired_delta_energy", C.c_double), ("rel_err_ode45", C.c_double),
                                                        ("helical_pert_eps_Aphi", C.c_double), ("boole_helical_pert", C.c_int32),
                                                        ("helical_pert_m_fourier", C.c_int32),
                                                        ("helical_pert_n_fourier", C.c_int32), ("reserved0", C.c_int32),
                                                        ("axi_noise_eps_A", C.c_double), ("axi_noise_eps_Phi", C.c_double),
                                                        ("non_axi_noise_eps_A", C.c_double),
                                                        ("boole_axi_noise_vector_pot", C.c_int32),
                                                        ("boole_axi_noise_elec_pot", C.c_int32),
                                                        ("boole_non_axi_noise_vector_pot", C.c_int32), ("noise_seed", C.c_int32)]


class _MeshDesc(C.Structure):
    _fields_ = [
        ("ntetr", C.c_int64), ("tetra_physics", C.POINTER(C.c_double)), ("tetra_grid", C.POINTER(C.c_int32)),
        ("cm_over_e", C.c_double), ("particle_mass", C.c_double), ("particle_charge", C.c_double),
        ("sign_sqg", C.c_int32), ("coord_system", C.c_int32), ("n_field_periods", C.c_int32),
        ("grid_kind", C.c_int32), ("grid_size", C.c_int32 * 3), ("pad0", C.c_int32),
        ("Rmin", C.c_double), ("Rmax", C.c_double), ("Zmin", C.c_double), ("Zmax", C.c_double),
        ("sfc_s_min", C.c_double), ("tetra_skew_coord", C.POINTER(C.c_double)),
    ]


class _GridSettings(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("grid_kind", "n1", "n2", "n3", "boole_n_field_periods",
                                         "n_field_periods_manual", "i_radial_spacing", "theta_geom_flux")] + \
               [(n, C.c_double) for n in ("sfc_s_min", "theta0_at_xpoint", "R0_analytic_circ", "a_analytic_circ",
                                          "B0_analytic_circ", "q0_analytic_circ", "q1_analytic_circ")] + \
               [(n, C.c_char_p) for n in ("g_file_filename", "convex_wall_filename", "netcdf_filename",
                                          "knots_SOLEDGE3X_EIRENE_filename", "triangles_SOLEDGE3X_EIRENE_filename")] + \
               [("bmod_multiplier", C.c_double), ("nwindow_r", C.c_int32), ("nwindow_z", C.c_int32)]


class _Counters(C.Structure):
    _fields_ = [("n_particles", C.c_int64), ("n_pushes", C.c_int64), ("n_lost", C.c_int64),
                ("n_finished", C.c_int64), ("n_fallback", C.c_int64 * 4), ("n_domain_errors", C.c_int64),
                ("kernel_ms", C.c_double), ("find_ms", C.c_double), ("n_adaptive", C.c_int64),
                ("n_lost_inner", C.c_int64), ("n_failed", C.c_int64)]


class _Diag(C.Structure):   # struct gorilla_diag
    _fields_ = [(n, C.c_int64) for n in ("n_particles", "n_pushes", "n_lost", "n_lost_outer", "n_lost_inner", "n_failed",
                                         "n_finished")] + [("n_fallback", C.c_int64 * 4), ("n_adaptive", C.c_int64),
                                                           ("n_sampled", C.c_int64)] + \
               [(n, C.c_double) for n in ("max_delta_energy", "rms_delta_energy", "max_delta_perpinv", "rms_delta_perpinv",
                                          "max_delta_p_phi", "rms_delta_p_phi")] + [("nranks", C.c_int32), ("reserved", C.c_int32)]


COMM_ID_BYTES = 128


class _EventSettings(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("boole_poincare_phi_0", "n_skip_phi_0", "boole_poincare_vpar_0", "boole_J_par",
                                         "n_skip_vpar_0", "boole_full_orbit", "n_skip_full_orbit", "reserved")]


# struct gorilla_event (include/gorilla_b200.h)
EVENT_DTYPE = np.dtype([("particle", np.int64), ("kind", np.int32), ("counter", np.int32), ("push", np.int64),
                        ("x", np.float64, 3), ("value", np.float64, 2), ("t", np.float64)])
EVENT_PHI_0, EVENT_VPAR_0, EVENT_FULL_ORBIT = 1, 2, 3


@dataclass
class Counters:
    n_particles: int
    n_pushes: int
    n_lost: int
    n_finished: int
    n_fallback: tuple
    n_domain_errors: int
    kernel_ms: float
    find_ms: float
    n_adaptive: int = 0
    n_lost_inner: int = 0
    n_failed: int = 0


@dataclass
class Diag:
    """struct gorilla_diag: counters since diag_reset and conservation statistics, reduced over all ranks."""
    n_particles: int
    n_pushes: int
    n_lost: int
    n_lost_outer: int
    n_lost_inner: int
    n_failed: int
    n_finished: int
    n_fallback: tuple
    n_adaptive: int
    n_sampled: int
    max_delta_energy: float
    rms_delta_energy: float
    max_delta_perpinv: float
    rms_delta_perpinv: float
    max_delta_p_phi: float
    rms_delta_p_phi: float
    nranks: int


# every symbol include/gorilla_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = (
    "gorilla_b200_init", "gorilla_b200_free", "gorilla_b200_last_error", "gorilla_b200_launch_count",
    "gorilla_b200_orbit_timestep", "gorilla_b200_orbit_timestep_dev", "gorilla_b200_orbit_timestep_trace",
    "gorilla_b200_orbit_timestep_optional", "gorilla_b200_orbit_timestep_optional_dev",
    "gorilla_b200_orbit_timestep_events", "gorilla_b200_orbit_timestep_events_dev",
    "gorilla_b200_find_tetra", "gorilla_b200_invariants", "gorilla_b200_invariants_dev",
    "gorilla_b200_get_counters", "gorilla_b200_sort_permutation_dev", "gorilla_b200_resort_dev",
    "gorilla_b200_set_host_resort", "gorilla_b200_set_launch_config", "gorilla_b200_fp64_peak", "gorilla_b200_set_prefetch", "gorilla_b200_set_gather", "gorilla_b200_get_gather",
    "gorilla_b200_comm_unique_id", "gorilla_b200_comm_init", "gorilla_b200_comm_free", "gorilla_b200_comm_allreduce_f64",
    "gorilla_b200_shard_range", "gorilla_b200_diag_reset", "gorilla_b200_diag_reduce_dev", "gorilla_b200_diag_reduce",
    "gorilla_mesh_build", "gorilla_mesh_get_desc", "gorilla_mesh_get_vertices", "gorilla_mesh_free",
    "gorilla_mesh_save", "gorilla_mesh_load", "gorilla_b200_abi_struct_sizes",
)


def load_library():
    """dlopen libgorilla_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise ImportError(f"{_LIB_PATH} not found: build it with `python -m gorilla_b200.build` "
                          "(there is no CPU fall-back for the orbit pusher)")
    lib = C.CDLL(str(_LIB_PATH))
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    lib.gorilla_b200_last_error.restype = C.c_char_p
    lib.gorilla_b200_launch_count.restype = i64
    lib.gorilla_b200_init.argtypes = [C.POINTER(_MeshDesc), C.POINTER(_Settings), C.POINTER(vp)]
    lib.gorilla_b200_free.argtypes = [vp]
    lib.gorilla_b200_free.restype = None
    lib.gorilla_b200_orbit_timestep.argtypes = [vp, i64, vp, vp, vp, dbl, vp, vp, vp, vp, vp]
    lib.gorilla_b200_orbit_timestep_dev.argtypes = [vp, i64, vp, vp, vp, dbl, vp, vp, vp, vp, vp, vp]
    lib.gorilla_b200_orbit_timestep_trace.argtypes = [vp, i64, vp, vp, vp, dbl, vp, vp, vp, vp, vp, i32, vp, vp]
    lib.gorilla_b200_orbit_timestep_optional.argtypes = [vp, i64, vp, vp, vp, dbl, vp, vp, vp, vp, vp, vp]
    lib.gorilla_b200_orbit_timestep_optional_dev.argtypes = [vp, i64, vp, vp, vp, dbl, vp, vp, vp, vp, vp, vp, vp]
    lib.gorilla_b200_debug_orbit_timestep_trace_optional.argtypes = [vp, i64, vp, vp, vp, dbl, vp, vp, vp, vp, vp, i32, vp,
                                                                     vp, vp]
    lib.gorilla_b200_orbit_timestep_events.argtypes = [vp, i64, vp, vp, vp, dbl, vp, vp, vp, vp, vp,
                                                       C.POINTER(_EventSettings), vp, vp, vp, vp, i64, C.POINTER(i64)]
    lib.gorilla_b200_orbit_timestep_events_dev.argtypes = [vp, i64, vp, vp, vp, dbl, vp, vp, vp, vp, vp,
                                                           C.POINTER(_EventSettings), vp, vp, vp, vp, i64, vp, vp]
    lib.gorilla_b200_find_tetra.argtypes = [vp, i64, vp, vp, vp, vp, vp, i32]
    lib.gorilla_b200_invariants.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.gorilla_b200_invariants_dev.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.gorilla_b200_get_counters.argtypes = [vp, C.POINTER(_Counters)]
    lib.gorilla_b200_sort_permutation_dev.argtypes = [vp, i64, vp, vp, vp]
    lib.gorilla_b200_resort_dev.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, i32, C.POINTER(vp), vp, vp]
    lib.gorilla_b200_set_host_resort.argtypes = [vp, i32]
    lib.gorilla_b200_comm_unique_id.argtypes = [vp]
    lib.gorilla_b200_comm_init.argtypes = [vp, vp, i32, i32]
    lib.gorilla_b200_comm_free.argtypes = [vp]
    lib.gorilla_b200_comm_allreduce_f64.argtypes = [vp, vp, i64, i32, vp]
    lib.gorilla_b200_shard_range.argtypes = [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    lib.gorilla_b200_diag_reset.argtypes = [vp, vp]
    lib.gorilla_b200_diag_reduce_dev.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(_Diag), vp]
    lib.gorilla_b200_diag_reduce.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(_Diag)]
    lib.gorilla_b200_set_launch_config.argtypes = [vp, i32, i32]
    lib.gorilla_b200_set_prefetch.argtypes = [vp, i32]
    lib.gorilla_b200_set_gather.argtypes = [vp, i32]
    lib.gorilla_b200_get_gather.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.gorilla_b200_debug_force_full.argtypes = [vp, i32]
    lib.gorilla_b200_debug_find_bins.argtypes = [vp, i32]
    lib.gorilla_b200_debug_use_group.argtypes = [vp, i32]
    lib.gorilla_b200_fp64_peak.argtypes = [C.POINTER(dbl), C.POINTER(dbl)]
    lib.gorilla_mesh_build.argtypes = [C.POINTER(_GridSettings), C.POINTER(_Settings), C.POINTER(vp)]
    lib.gorilla_mesh_get_desc.argtypes = [vp, C.POINTER(_MeshDesc)]
    lib.gorilla_mesh_get_vertices.argtypes = [vp, C.POINTER(i64), C.POINTER(C.POINTER(dbl)), C.POINTER(C.POINTER(dbl))]
    lib.gorilla_mesh_free.argtypes = [vp]
    lib.gorilla_mesh_free.restype = None
    lib.gorilla_mesh_save.argtypes = [C.POINTER(_MeshDesc), i64, vp, vp, C.c_char_p]
    lib.gorilla_mesh_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    # the ctypes mirrors above against the layout the library was compiled with
    sizes = (i64 * 7)()
    lib.gorilla_b200_abi_struct_sizes.argtypes = [C.POINTER(i64)]
    if lib.gorilla_b200_abi_struct_sizes(sizes) != 0:
        raise ImportError("gorilla_b200_abi_struct_sizes failed")
    mine = (C.sizeof(_Settings), C.sizeof(_MeshDesc), C.sizeof(_Counters), C.sizeof(_Diag), C.sizeof(_GridSettings),
            EVENT_DTYPE.itemsize, C.sizeof(_EventSettings))
    if tuple(sizes) != mine:
        raise ImportError(f"struct layouts of {_LIB_PATH.name} {tuple(sizes)} differ from this binding's {mine}: "
                          "rebuild the library (python -m gorilla_b200.build)")
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise GorillaError(rc, load_library().gorilla_b200_last_error().decode(errors="replace"))


def launch_count() -> int:
    """Kernels launched by the library in this process (bench.py's gpu_launches)."""
    return int(load_library().gorilla_b200_launch_count())


def fp64_peak() -> tuple[float, float]:
    """(DFMA, DMUL+DADD) thread-instructions per second of the current device, measured."""
    a, b = C.c_double(), C.c_double()
    _check(load_library().gorilla_b200_fp64_peak(C.byref(a), C.byref(b)))
    return a.value, b.value


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0); hand the bytes to the other ranks."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(load_library().gorilla_b200_comm_unique_id(buf))
    return buf.raw


def shard_range(n_total: int, rank: int, nranks: int) -> tuple[int, int]:
    """(first, count) of the contiguous shard [r N/G, (r+1) N/G) (gorilla_b200_shard_range)."""
    a, b = C.c_int64(), C.c_int64()
    _check(load_library().gorilla_b200_shard_range(int(n_total), int(rank), int(nranks), C.byref(a), C.byref(b)))
    return a.value, b.value


def _require(a, dtype, shape=None, name="array", allow_none=False):
    """Every array that crosses the C ABI as a raw pointer: right dtype, C-contiguous, right shape -- or a TypeError
    instead of a silent overflow."""
    if a is None:
        if allow_none:
            return None
        raise TypeError(f"{name} must not be None")
    if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dtype) or not a.flags.c_contiguous:
        raise TypeError(f"{name} must be a C-contiguous numpy array of {np.dtype(dtype).name}")
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise TypeError(f"{name} must have shape {tuple(shape)}, not {tuple(a.shape)}")
    return a


def _c_settings(s: GorillaSettings) -> _Settings:
    cs = _Settings()
    for name, typ in _Settings._fields_:
        v = getattr(s, name, 0)
        setattr(cs, name, float(v) if typ is C.c_double else int(v))
    return cs


def _c_grid(g: TetraGridSettings) -> _GridSettings:
    cg = _GridSettings()
    for name, typ in _GridSettings._fields_:
        v = getattr(g, name)
        if typ is C.c_char_p:
            setattr(cg, name, str(v).encode() if v else None)
        elif typ is C.c_double:
            setattr(cg, name, float(v))
        else:
            setattr(cg, name, int(v))
    return cg


class Mesh:
    """Host mesh in the reference's own AoS layout (tetra_physics [ntetr,142] f64, tetra_grid [ntetr,20] i32).

    Either built by the library (build_mesh: make_tetra_grid + make_tetra_physics + check_tetra_overlaps)
    or wrapped around arrays that came from elsewhere (from_arrays), e.g. dumped from a Fortran run.
    """

    def __init__(self):
        self._handle = None
        self.tetra_physics: np.ndarray | None = None
        self.tetra_grid: np.ndarray | None = None
        self.verts_rphiz: np.ndarray | None = None
        self.verts_sthetaphi: np.ndarray | None = None
        self.tetra_skew_coord: np.ndarray | None = None   # [ntetr,168], handover_processing_kind = 2 only
        self.scalars: dict = {}

    @classmethod
    def from_arrays(cls, tetra_physics, tetra_grid, **scalars) -> "Mesh":
        m = cls()
        m.tetra_physics = np.ascontiguousarray(tetra_physics, dtype=np.float64)
        m.tetra_grid = np.ascontiguousarray(tetra_grid, dtype=np.int32)
        assert m.tetra_physics.shape[1] == 142 and m.tetra_grid.shape[1] == 20
        m.scalars = dict(scalars)
        return m

    @property
    def ntetr(self) -> int:
        return int(self.tetra_physics.shape[0])

    def desc(self) -> _MeshDesc:
        d = _MeshDesc()
        d.ntetr = self.ntetr
        d.tetra_physics = self.tetra_physics.ctypes.data_as(C.POINTER(C.c_double))
        d.tetra_grid = self.tetra_grid.ctypes.data_as(C.POINTER(C.c_int32))
        s = self.scalars
        for k in ("cm_over_e", "particle_mass", "particle_charge", "Rmin", "Rmax", "Zmin", "Zmax", "sfc_s_min"):
            setattr(d, k, float(s.get(k, 0.0)))
        for k in ("sign_sqg", "coord_system", "n_field_periods", "grid_kind"):
            setattr(d, k, int(s[k]))
        for i in range(3):
            d.grid_size[i] = int(s["grid_size"][i])
        if self.tetra_skew_coord is not None:
            d.tetra_skew_coord = self.tetra_skew_coord.ctypes.data_as(C.POINTER(C.c_double))
        return d

    def save(self, path) -> None:
        """Write the mesh as a versioned .gmesh file (gorilla_mesh_save)."""
        d = self.desc()
        nv = 0 if self.verts_rphiz is None else int(self.verts_rphiz.shape[0])
        _check(load_library().gorilla_mesh_save(C.byref(d), nv, _ptr(self.verts_rphiz) if nv else None,
                                                _ptr(self.verts_sthetaphi) if nv and self.verts_sthetaphi is not None else None,
                                                str(path).encode()))



class _MeshOwner:
    """Owns the C-side gorilla_mesh.  The numpy views handed out by _mesh_from_handle keep a reference to it (through
    their base), so the backing memory outlives every view, not only the Mesh object."""

    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        if self.handle is not None and _lib is not None:
            _lib.gorilla_mesh_free(self.handle)
            self.handle = None


class _View(np.ndarray):
    """ndarray view that carries a reference to the owner of its memory."""
    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._owner = getattr(obj, "_owner", None)


def _owned_view(ptr, shape, owner):
    v = np.ctypeslib.as_array(ptr, shape=shape).view(_View)
    v._owner = owner
    return v


def _mesh_from_handle(h) -> Mesh:
    lib = load_library()
    d = _MeshDesc()
    _check(lib.gorilla_mesh_get_desc(h, C.byref(d)))
    m = Mesh()
    owner = _MeshOwner(h)
    m._handle = owner
    nt = int(d.ntetr)
    m.tetra_physics = _owned_view(d.tetra_physics, (nt, 142), owner)
    m.tetra_grid = _owned_view(d.tetra_grid, (nt, 20), owner)
    if d.tetra_skew_coord:
        m.tetra_skew_coord = _owned_view(d.tetra_skew_coord, (nt, 168), owner)
    nv = C.c_int64()
    pr, ps = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
    _check(lib.gorilla_mesh_get_vertices(h, C.byref(nv), C.byref(pr), C.byref(ps)))
    if nv.value > 0:
        m.verts_rphiz = _owned_view(pr, (nv.value, 3), owner)
        if ps:
            m.verts_sthetaphi = _owned_view(ps, (nv.value, 3), owner)
    m.scalars = dict(
        cm_over_e=d.cm_over_e, particle_mass=d.particle_mass, particle_charge=d.particle_charge,
        sign_sqg=d.sign_sqg, coord_system=d.coord_system, n_field_periods=d.n_field_periods,
        grid_kind=d.grid_kind, grid_size=tuple(d.grid_size), Rmin=d.Rmin, Rmax=d.Rmax, Zmin=d.Zmin, Zmax=d.Zmax,
        sfc_s_min=d.sfc_s_min,
    )
    return m


def load_mesh(path) -> Mesh:
    """Read a .gmesh file written by Mesh.save / gorilla_mesh_save (raises GorillaError on a bad or corrupted file)."""
    h = C.c_void_p()
    _check(load_library().gorilla_mesh_load(str(path).encode(), C.byref(h)))
    return _mesh_from_handle(h)


def build_mesh(grid: TetraGridSettings, settings: GorillaSettings) -> Mesh:
    """Grid + physics half of initialize_gorilla (host, runs once)."""
    lib = load_library()
    h = C.c_void_p()
    cg, cs = _c_grid(grid), _c_settings(settings)
    _check(lib.gorilla_mesh_build(C.byref(cg), C.byref(cs), C.byref(h)))
    return _mesh_from_handle(h)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Gorilla:
    """A mesh resident on one GPU plus the settings: what initialize_gorilla leaves behind."""

    def __init__(self, mesh: Mesh, settings: GorillaSettings):
        lib = load_library()
        self.mesh = mesh
        self.settings = settings
        self._h = C.c_void_p()
        d, cs = mesh.desc(), _c_settings(settings)
        _check(lib.gorilla_b200_init(C.byref(d), C.byref(cs), C.byref(self._h)))

    def close(self):
        if self._h is not None and self._h.value:
            load_library().gorilla_b200_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference API -------------------------------------------------------------------------
    def check_coordinate_domain(self, x: np.ndarray) -> None:
        """check_coordinate_domain (orbit_timestep_gorilla.f90:278-358), in place on x[n,3] (host logic)."""
        s = self.mesh.scalars
        per = 2.0 * math.pi / s["n_field_periods"]
        x = x.reshape(-1, 3)

        def modulo(a, p):   # Fortran MODULO for reals as gfortran expands it: fmod (exact) + sign fix
            r = np.fmod(a, p)
            r = np.where((r != 0.0) & ((r < 0.0) != (p < 0.0)), r + p, r)
            return np.where(r == 0.0, math.copysign(0.0, p), r)
        if s["coord_system"] == 1:
            if self.settings.boole_periodic_relocation:
                x[:, 1] = modulo(x[:, 1], per)
            elif np.any((x[:, 1] < 0.0) | (x[:, 1] > per)):
                raise GorillaError(4, "Particle coordinate phi outside [0, 2 pi/n_field_periods]")
        else:
            if np.any((x[:, 0] < s["sfc_s_min"]) | (x[:, 0] > 1.0)):
                raise GorillaError(4, "Particle flux coordinate s outside [sfc_s_min, 1]")
            if self.settings.boole_periodic_relocation:
                x[:, 1] = modulo(x[:, 1], 2.0 * math.pi)
                x[:, 2] = modulo(x[:, 2], per)
            elif np.any((x[:, 1] < 0) | (x[:, 1] > 2 * math.pi) | (x[:, 2] < 0) | (x[:, 2] > per)):
                raise GorillaError(4, "Particle coordinate theta/phi outside the domain")

    def find_tetra(self, x, vpar, vperp, sign_t_step: int = 1):
        """find_tetra (find_tetra_mod.f90:283-600) for a batch; returns (ind_tetr, iface); x may be updated."""
        n = x.shape[0]
        _require(x, np.float64, (n, 3), "x"); _require(vpar, np.float64, (n,), "vpar"); _require(vperp, np.float64, (n,), "vperp")
        ind, ifc = np.empty(n, np.int32), np.empty(n, np.int32)
        _check(load_library().gorilla_b200_find_tetra(self._h, n, _ptr(x), _ptr(vpar), _ptr(vperp), _ptr(ind),
                                                      _ptr(ifc), int(sign_t_step)))
        return ind, ifc

    @staticmethod
    def _require_state(n, x, vpar, vperp, boole_initialized, ind_tetr, iface, t_remain_out=None, n_pushes=None):
        _require(x, np.float64, (n, 3), "x")
        _require(vpar, np.float64, (n,), "vpar")
        _require(vperp, np.float64, (n,), "vperp")
        _require(boole_initialized, np.int32, (n,), "boole_initialized")
        _require(ind_tetr, np.int32, (n,), "ind_tetr")
        _require(iface, np.int32, (n,), "iface")
        _require(t_remain_out, np.float64, (n,), "t_remain_out", allow_none=True)
        _require(n_pushes, np.int64, (n,), "n_pushes", allow_none=True)

    def orbit_timestep_gorilla(self, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface,
                               t_remain_out=None, n_pushes=None, trace_cap: int = 0, optional_quantities=None):
        """Batched orbit_timestep_gorilla; all arrays are updated in place (numpy, C-contiguous):
        x [n,3] f64, vpar/vperp [n] f64, boole_initialized/ind_tetr/iface [n] i32.
        optional_quantities: [n,4] f64 output (t_hamiltonian, gyrophase, vpar_int, vpar2_int summed over the pushes of
        the time step; pusher_tetra_poly's optional_quantities argument, pusher_tetra_poly.f90:204,662-667).
        Returns (trace_ind_tetr, trace_iface) when trace_cap > 0, else None."""
        lib = load_library()
        n = x.shape[0]
        self._require_state(n, x, vpar, vperp, boole_initialized, ind_tetr, iface, t_remain_out, n_pushes)
        if optional_quantities is not None:
            oq = optional_quantities
            if oq.dtype != np.float64 or not oq.flags.c_contiguous or oq.shape != (n, 4):
                raise TypeError("optional_quantities must be a C-contiguous float64 [n,4] array")
            if trace_cap > 0:
                tt, tf = np.zeros((n, trace_cap), np.int32), np.zeros((n, trace_cap), np.int32)
                _check(lib.gorilla_b200_debug_orbit_timestep_trace_optional(
                    self._h, n, _ptr(x), _ptr(vpar), _ptr(vperp), float(t_step), _ptr(boole_initialized), _ptr(ind_tetr),
                    _ptr(iface), _ptr(t_remain_out), _ptr(n_pushes), trace_cap, _ptr(tt), _ptr(tf), _ptr(oq)))
                return tt, tf
            _check(lib.gorilla_b200_orbit_timestep_optional(self._h, n, _ptr(x), _ptr(vpar), _ptr(vperp), float(t_step),
                                                            _ptr(boole_initialized), _ptr(ind_tetr), _ptr(iface),
                                                            _ptr(t_remain_out), _ptr(n_pushes), _ptr(oq)))
            return None
        if trace_cap > 0:
            tt, tf = np.zeros((n, trace_cap), np.int32), np.zeros((n, trace_cap), np.int32)
            _check(lib.gorilla_b200_orbit_timestep_trace(self._h, n, _ptr(x), _ptr(vpar), _ptr(vperp), float(t_step),
                                                         _ptr(boole_initialized), _ptr(ind_tetr), _ptr(iface),
                                                         _ptr(t_remain_out), _ptr(n_pushes), trace_cap, _ptr(tt),
                                                         _ptr(tf)))
            return tt, tf
        _check(lib.gorilla_b200_orbit_timestep(self._h, n, _ptr(x), _ptr(vpar), _ptr(vperp), float(t_step),
                                               _ptr(boole_initialized), _ptr(ind_tetr), _ptr(iface),
                                               _ptr(t_remain_out), _ptr(n_pushes)))
        return None

    def orbit_timestep_gorilla_dev(self, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface,
                                   t_remain_out=None, n_pushes=None, stream=None):
        """Same on torch CUDA tensors already resident in HBM (no copies, no sync)."""
        def dp(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        _check(load_library().gorilla_b200_orbit_timestep_dev(
            self._h, x.shape[0], dp(x), dp(vpar), dp(vperp), float(t_step), dp(boole_initialized), dp(ind_tetr),
            dp(iface), dp(t_remain_out), dp(n_pushes), C.c_void_p(stream or 0)))

    def orbit_timestep_gorilla_optional_dev(self, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface,
                                            optional_quantities, t_remain_out=None, n_pushes=None, stream=None):
        """orbit_timestep_gorilla_dev plus the optional quantities ([n,4] f64 CUDA tensor)."""
        def dp(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        _check(load_library().gorilla_b200_orbit_timestep_optional_dev(
            self._h, x.shape[0], dp(x), dp(vpar), dp(vperp), float(t_step), dp(boole_initialized), dp(ind_tetr),
            dp(iface), dp(t_remain_out), dp(n_pushes), dp(optional_quantities), C.c_void_p(stream or 0)))

    def orbit_timestep_gorilla_events(self, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface, par_adiab_inv,
                                      counter_vpar_0, counter_phi_0, event_cap: int, *, boole_poincare_phi_0=True,
                                      n_skip_phi_0=1, boole_poincare_vpar_0=True, boole_J_par=True, n_skip_vpar_0=1,
                                      t_remain_out=None, n_pushes=None, boole_full_orbit=False, n_skip_full_orbit=1):
        """orbit_timestep_gorilla with the event capture of gorilla_plot_orbit_integration (gorilla_plot_mod.f90:553-638):
        toroidal mappings, banana tips / J_par and (boole_full_orbit) the orbit point, p_phi and E_tot after every
        n_skip_full_orbit-th push.  par_adiab_inv [n] f64, counter_vpar_0 / counter_phi_0 [n] i32 carry the
        per-particle state between calls.  Returns (events, n_events): a structured array (EVENT_DTYPE) sorted by
        (particle, push, kind) holding min(n_events, event_cap) records."""
        n = x.shape[0]
        self._require_state(n, x, vpar, vperp, boole_initialized, ind_tetr, iface, t_remain_out, n_pushes)
        _require(par_adiab_inv, np.float64, (n,), "par_adiab_inv")
        _require(counter_vpar_0, np.int32, (n,), "counter_vpar_0")
        _require(counter_phi_0, np.int32, (n,), "counter_phi_0")
        cfg = _EventSettings(int(boole_poincare_phi_0), int(n_skip_phi_0), int(boole_poincare_vpar_0), int(boole_J_par),
                             int(n_skip_vpar_0), int(boole_full_orbit), int(n_skip_full_orbit), 0)
        ev = np.zeros(max(event_cap, 1), EVENT_DTYPE)
        nev = C.c_int64(0)
        _check(load_library().gorilla_b200_orbit_timestep_events(
            self._h, n, _ptr(x), _ptr(vpar), _ptr(vperp), float(t_step), _ptr(boole_initialized), _ptr(ind_tetr),
            _ptr(iface), _ptr(t_remain_out), _ptr(n_pushes), C.byref(cfg), _ptr(par_adiab_inv), _ptr(counter_vpar_0),
            _ptr(counter_phi_0), _ptr(ev), int(event_cap), C.byref(nev)))
        ev = ev[:min(nev.value, event_cap)]
        return ev[np.lexsort((ev["kind"], ev["push"], ev["particle"]))], nev.value

    # ---- diagnostics ---------------------------------------------------------------------------
    def invariants(self, x, vpar, vperp, ind_tetr):
        """(energy_tot_func, p_phi_func, perpinv) per particle (supporting_functions_mod.f90:279-408)."""
        n = x.shape[0]
        _require(x, np.float64, (n, 3), "x"); _require(vpar, np.float64, (n,), "vpar"); _require(vperp, np.float64, (n,), "vperp")
        _require(ind_tetr, np.int32, (n,), "ind_tetr")
        e, p, mu = np.empty(n), np.empty(n), np.empty(n)
        _check(load_library().gorilla_b200_invariants(self._h, n, _ptr(x), _ptr(vpar), _ptr(vperp), _ptr(ind_tetr),
                                                      _ptr(e), _ptr(p), _ptr(mu)))
        return e, p, mu

    def invariants_dev(self, x, vpar, vperp, ind_tetr, energy, p_phi, perpinv, stream=None):
        def dp(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        _check(load_library().gorilla_b200_invariants_dev(self._h, x.shape[0], dp(x), dp(vpar), dp(vperp),
                                                          dp(ind_tetr), dp(energy), dp(p_phi), dp(perpinv),
                                                          C.c_void_p(stream or 0)))

    def counters(self) -> Counters:
        c = _Counters()
        _check(load_library().gorilla_b200_get_counters(self._h, C.byref(c)))
        return Counters(c.n_particles, c.n_pushes, c.n_lost, c.n_finished, tuple(c.n_fallback), c.n_domain_errors,
                        c.kernel_ms, c.find_ms, c.n_adaptive, c.n_lost_inner, c.n_failed)

    def sort_permutation_dev(self, ind_tetr, perm, stream=None):
        _check(load_library().gorilla_b200_sort_permutation_dev(self._h, ind_tetr.shape[0],
                                                                C.c_void_p(ind_tetr.data_ptr()),
                                                                C.c_void_p(perm.data_ptr()), C.c_void_p(stream or 0)))

    def resort_dev(self, x, vpar, vperp, boole_initialized, ind_tetr, iface, extra=(), perm_out=None, stream=None):
        """Re-sort a resident batch (torch CUDA tensors) by tetrahedron index in place (gorilla_b200_resort_dev); `extra`:
        further float64 [n] tensors permuted alongside."""
        def dp(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        ex = (C.c_void_p * max(1, len(extra)))(*[t.data_ptr() for t in extra])
        _check(load_library().gorilla_b200_resort_dev(self._h, x.shape[0], dp(x), dp(vpar), dp(vperp), dp(boole_initialized),
                                                      dp(ind_tetr), dp(iface), len(extra), ex, dp(perm_out),
                                                      C.c_void_p(stream or 0)))

    def set_host_resort(self, on: bool):
        """Host-pointer orbit_timestep_gorilla: sort each uploaded batch by tetrahedron before the push (caller's order is
        restored before the download)."""
        _check(load_library().gorilla_b200_set_host_resort(self._h, int(on)))

    # ---- multi-GPU + diagnostics reduction -----------------------------------------------------
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        """Join the NCCL communicator of the job (one handle per GPU); unique_id from comm_unique_id() on rank 0."""
        assert len(unique_id) == COMM_ID_BYTES
        _check(load_library().gorilla_b200_comm_init(self._h, C.c_char_p(unique_id), int(rank), int(nranks)))

    def comm_free(self):
        _check(load_library().gorilla_b200_comm_free(self._h))

    def comm_allreduce_f64(self, buf, op: str = "sum", stream=None):
        """In-place all-reduce of a small float64 CUDA tensor over the handle's communicator (no-op on one GPU)."""
        _check(load_library().gorilla_b200_comm_allreduce_f64(self._h, C.c_void_p(buf.data_ptr()), buf.numel(),
                                                              {"sum": 0, "max": 1, "min": 2}[op], C.c_void_p(stream or 0)))

    def diag_reset(self, stream=None):
        _check(load_library().gorilla_b200_diag_reset(self._h, C.c_void_p(stream or 0)))

    def diag_reduce_dev(self, x, vpar, vperp, ind_tetr, energy_ref=None, p_phi_ref=None, perpinv_ref=None,
                        stream=None) -> Diag:
        """gorilla_b200_diag_reduce_dev: counters since diag_reset + max/rms drift of E, perpinv, p_phi against the given
        reference values, reduced over all ranks of the communicator (collective)."""
        def dp(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        d = _Diag()
        _check(load_library().gorilla_b200_diag_reduce_dev(self._h, x.shape[0], dp(x), dp(vpar), dp(vperp), dp(ind_tetr),
                                                           dp(energy_ref), dp(p_phi_ref), dp(perpinv_ref), C.byref(d),
                                                           C.c_void_p(stream or 0)))
        return Diag(d.n_particles, d.n_pushes, d.n_lost, d.n_lost_outer, d.n_lost_inner, d.n_failed, d.n_finished,
                    tuple(d.n_fallback), d.n_adaptive, d.n_sampled, d.max_delta_energy, d.rms_delta_energy,
                    d.max_delta_perpinv, d.rms_delta_perpinv, d.max_delta_p_phi, d.rms_delta_p_phi, d.nranks)

    def set_prefetch(self, mode: int):
        """Neighbour-record prefetch of the push kernels: 1 on, 0 off, -1 auto (gorilla_b200_set_prefetch)."""
        _check(load_library().gorilla_b200_set_prefetch(self._h, int(mode)))

    def set_gather(self, mode: int):
        """Record gather of the order-2 and RK4 kernels: 0 vector loads, 1 per-lane bulk copies (TMA) one push ahead,
        2 warp-cooperative cp.async copies one push ahead, -1 auto."""
        _check(load_library().gorilla_b200_set_gather(self._h, int(mode)))

    def get_gather(self) -> int:
        """The gather mode in effect (what -1 resolved to)."""
        m = C.c_int32(0)
        _check(load_library().gorilla_b200_get_gather(self._h, C.byref(m)))
        return int(m.value)

    def diag_reduce(self, x, vpar, vperp, ind_tetr, energy_ref=None, p_phi_ref=None, perpinv_ref=None) -> Diag:
        """gorilla_b200_diag_reduce: the same reduction from numpy arrays on the host (collective over the communicator)."""
        n = x.shape[0]
        _require(x, np.float64, (n, 3), "x"); _require(vpar, np.float64, (n,), "vpar"); _require(vperp, np.float64, (n,), "vperp")
        _require(ind_tetr, np.int32, (n,), "ind_tetr")
        for a, nm in ((energy_ref, "energy_ref"), (p_phi_ref, "p_phi_ref"), (perpinv_ref, "perpinv_ref")):
            _require(a, np.float64, (n,), nm, allow_none=True)
        d = _Diag()
        _check(load_library().gorilla_b200_diag_reduce(self._h, n, _ptr(x), _ptr(vpar), _ptr(vperp), _ptr(ind_tetr),
                                                       _ptr(energy_ref), _ptr(p_phi_ref), _ptr(perpinv_ref), C.byref(d)))
        return Diag(d.n_particles, d.n_pushes, d.n_lost, d.n_lost_outer, d.n_lost_inner, d.n_failed, d.n_finished,
                    tuple(d.n_fallback), d.n_adaptive, d.n_sampled, d.max_delta_energy, d.rms_delta_energy,
                    d.max_delta_perpinv, d.rms_delta_perpinv, d.max_delta_p_phi, d.rms_delta_p_phi, d.nranks)

    def set_launch_config(self, ctas_per_sm: int = 0, threads_per_cta: int = 0):
        _check(load_library().gorilla_b200_set_launch_config(self._h, ctas_per_sm, threads_per_cta))

    def _debug_force_full(self, on: bool):
        load_library().gorilla_b200_debug_force_full(self._h, int(on))

    def _debug_use_group(self, mode: int):
        """orders 3/4: 0 = 4-warp CTAs, 1 = lock-step solver kernel, 2 = lock-step kernel with the solves re-binned by mode"""
        load_library().gorilla_b200_debug_use_group(self._h, int(mode))

    def _debug_find_bins(self, on: bool):
        load_library().gorilla_b200_debug_find_bins(self._h, int(on))


def initialize_gorilla(grid: TetraGridSettings, settings: GorillaSettings) -> Gorilla:
    """initialize_gorilla (orbit_timestep_gorilla.f90:151-274): build the mesh on the host, upload it."""
    return Gorilla(build_mesh(grid, settings), settings)
