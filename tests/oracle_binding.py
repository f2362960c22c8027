"""ctypes binding of the CPU oracle (oracle/libgorilla_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent.parent / "oracle"
ORACLE_LIB = ORACLE_DIR / "libgorilla_oracle.so"


class GorMesh(C.Structure):
    _fields_ = [
        ("ntetr", C.c_int64), ("tetra_physics", C.POINTER(C.c_double)), ("tetra_grid", C.POINTER(C.c_int32)),
        ("cm_over_e", C.c_double), ("particle_mass", C.c_double), ("particle_charge", C.c_double),
        ("sign_sqg", C.c_int32), ("coord_system", C.c_int32), ("n_field_periods", C.c_int32), ("grid_kind", C.c_int32),
        ("grid_size", C.c_int32 * 3),
        ("Rmin", C.c_double), ("Rmax", C.c_double), ("Zmin", C.c_double), ("Zmax", C.c_double),
        ("sfc_s_min", C.c_double),
        ("ipusher", C.c_int32), ("poly_order", C.c_int32), ("boole_guess", C.c_int32),
        ("boole_strong_electric_field", C.c_int32), ("boole_periodic_relocation", C.c_int32),
        ("boole_dt_dtau", C.c_int32), ("i_time_tracing_option", C.c_int32),
        ("boole_time_hamiltonian", C.c_int32), ("boole_gyrophase", C.c_int32), ("boole_vpar_int", C.c_int32),
        ("boole_vpar2_int", C.c_int32), ("boole_adaptive_time_steps", C.c_int32), ("max_n_intermediate_steps", C.c_int32),
        ("desired_delta_energy", C.c_double), ("tetra_skew_coord", C.POINTER(C.c_double)),
        ("handover_processing_kind", C.c_int32),
        ("i_precomp", C.c_int32), ("boole_newton_precalc", C.c_int32), ("tetra_physics_poly4", C.POINTER(C.c_double)),
        ("boole_pusher_ode45", C.c_int32), ("pad_ode45", C.c_int32), ("rel_err_ode45", C.c_double),
    ]


class GorTrace(C.Structure):
    _fields_ = [("n_pushes", C.c_int64), ("cap", C.c_int64), ("ind_tetr", C.POINTER(C.c_int32)),
                ("iface", C.POINTER(C.c_int32)), ("n_fallback", C.c_int64 * 4), ("n_solver_iters", C.c_int64),
                ("n_solver_calls", C.c_int64), ("optional_quantities", C.c_double * 4), ("n_adaptive", C.c_int64)]


class GorEvent(C.Structure):
    _fields_ = [("particle", C.c_int64), ("kind", C.c_int32), ("counter", C.c_int32), ("push", C.c_int64),
                ("x", C.c_double * 3), ("value", C.c_double * 2)]


class GorEventSettings(C.Structure):
    _fields_ = [("boole_poincare_phi_0", C.c_int32), ("n_skip_phi_0", C.c_int32), ("boole_poincare_vpar_0", C.c_int32),
                ("boole_J_par", C.c_int32), ("n_skip_vpar_0", C.c_int32), ("boole_full_orbit", C.c_int32),
                ("n_skip_full_orbit", C.c_int32), ("reserved", C.c_int32)]


EVENT_DTYPE = np.dtype([("particle", np.int64), ("kind", np.int32), ("counter", np.int32), ("push", np.int64),
                        ("x", np.float64, 3), ("value", np.float64, 2), ("t", np.float64)])


def build_oracle(force: bool = False) -> Path:
    if force or not ORACLE_LIB.exists() or ORACLE_LIB.stat().st_mtime < max(
            (ORACLE_DIR / "gorilla_oracle.c").stat().st_mtime, (ORACLE_DIR / "gorilla_oracle.h").stat().st_mtime):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-B"], check=True, capture_output=True)
    return ORACLE_LIB


_lib = None


def load_oracle():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(str(ORACLE_LIB))
        d, i32p, dp = C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_double)
        L.gor_orbit_timestep.argtypes = [C.POINTER(GorMesh), dp, dp, dp, d, i32p, i32p, i32p, dp, C.POINTER(GorTrace)]
        L.gor_orbit_timestep_batch.argtypes = [C.POINTER(GorMesh), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, d,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.gor_orbit_timestep_batch.restype = C.c_int64
        L.gor_orbit_timestep_batch_opt.argtypes = [C.POINTER(GorMesh), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, d,
                                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_int]
        L.gor_orbit_timestep_batch_opt.restype = C.c_int64
        L.gor_orbit_timestep_events.argtypes = [C.POINTER(GorMesh), dp, dp, dp, d, i32p, i32p, i32p, dp, C.POINTER(GorTrace),
                                                C.POINTER(GorEventSettings), dp, i32p, i32p, C.c_int64, C.c_void_p,
                                                C.c_int64, C.POINTER(C.c_int64)]
        L.gor_make_precomp_poly4.argtypes = [C.POINTER(GorMesh), dp]
        L.gor_make_precomp_poly4.restype = None
        L.gor_find_tetra.argtypes = [C.POINTER(GorMesh), dp, d, d, i32p, i32p, C.c_int]
        L.gor_find_tetra.restype = None
        L.gor_check_coordinate_domain.argtypes = [C.POINTER(GorMesh), dp]
        L.gor_energy_tot.argtypes = [C.POINTER(GorMesh), dp, d, C.c_int32]
        L.gor_energy_tot.restype = d
        L.gor_p_phi.argtypes = [C.POINTER(GorMesh), d, dp, C.c_int32]
        L.gor_p_phi.restype = d
        L.gor_bmod.argtypes = [C.POINTER(GorMesh), dp, C.c_int32]
        L.gor_bmod.restype = d
        for name, nargs in (("gor_quadratic_solver1", 3), ("gor_quadratic_solver2", 3), ("gor_cubic_solver", 4)):
            getattr(L, name).argtypes = [d] * nargs
            getattr(L, name).restype = d
        L.gor_quartic_solver.argtypes = [C.c_int] + [d] * 5
        L.gor_quartic_solver.restype = d
        L.gor_quadratic_roots.argtypes = [d, d, C.POINTER(C.c_int), dp]
        L.gor_cubic_roots.argtypes = [d, d, d, C.POINTER(C.c_int), dp]
        L.gor_quartic_roots.argtypes = [d, d, d, d, C.POINTER(C.c_int), dp]
        L.gor_cmplx_roots_gen.argtypes = [C.c_int, dp, dp]
        L.gor_frac_jump_phase.argtypes = [C.c_int, dp]
        _lib = L
    return _lib


class OracleMesh:
    """The oracle's view of a gorilla_b200.Mesh + settings (same AoS arrays, no copies)."""

    def __init__(self, mesh, settings):
        self.mesh = mesh  # keeps the arrays alive
        s = mesh.scalars
        m = GorMesh()
        m.ntetr = mesh.ntetr
        m.tetra_physics = mesh.tetra_physics.ctypes.data_as(C.POINTER(C.c_double))
        m.tetra_grid = mesh.tetra_grid.ctypes.data_as(C.POINTER(C.c_int32))
        for k in ("cm_over_e", "particle_mass", "particle_charge", "Rmin", "Rmax", "Zmin", "Zmax", "sfc_s_min"):
            setattr(m, k, float(s.get(k, 0.0)))
        for k in ("sign_sqg", "coord_system", "n_field_periods", "grid_kind"):
            setattr(m, k, int(s[k]))
        for i in range(3):
            m.grid_size[i] = int(s["grid_size"][i])
        m.ipusher = settings.ipusher
        m.poly_order = settings.poly_order
        m.boole_guess = int(settings.boole_guess)
        m.boole_strong_electric_field = int(settings.boole_strong_electric_field)
        m.boole_periodic_relocation = int(settings.boole_periodic_relocation)
        m.boole_dt_dtau = int(settings.boole_dt_dtau)
        m.i_time_tracing_option = int(settings.i_time_tracing_option)
        m.boole_time_hamiltonian = int(settings.boole_time_Hamiltonian)
        m.boole_gyrophase = int(settings.boole_gyrophase)
        m.boole_vpar_int = int(settings.boole_vpar_int)
        m.boole_vpar2_int = int(settings.boole_vpar2_int)
        m.boole_adaptive_time_steps = int(settings.boole_adaptive_time_steps)
        m.max_n_intermediate_steps = int(settings.max_n_intermediate_steps)
        m.desired_delta_energy = float(settings.desired_delta_energy)
        m.handover_processing_kind = int(settings.handover_processing_kind)
        if settings.handover_processing_kind == 2:
            m.tetra_skew_coord = mesh.tetra_skew_coord.ctypes.data_as(C.POINTER(C.c_double))
        m.boole_pusher_ode45 = int(getattr(settings, "boole_pusher_ode45", False))
        m.rel_err_ode45 = float(getattr(settings, "rel_err_ode45", 1.0e-8))
        m.i_precomp = int(getattr(settings, "i_precomp", 0))
        m.boole_newton_precalc = int(getattr(settings, "boole_newton_precalc", False))
        self.L = load_oracle()
        self.poly4 = None
        if m.i_precomp != 0 or m.boole_newton_precalc:
            # make_precomp_poly4 (tetra_physics_poly_precomp_mod.f90:160-476): [ntetr][544], built by the oracle itself
            self.poly4 = np.empty((mesh.ntetr, 544))
            self.c = m
            self.L.gor_make_precomp_poly4(C.byref(m), self.poly4.ctypes.data_as(C.POINTER(C.c_double)))
            m.tetra_physics_poly4 = self.poly4.ctypes.data_as(C.POINTER(C.c_double))
        self.c = m

    def orbit_timestep_batch(self, x, vpar, vperp, t_step, binit, ind_tetr, iface, t_remain_out=None, n_pushes=None,
                             nthreads: int = 0, optional_quantities=None) -> int:
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        return int(self.L.gor_orbit_timestep_batch_opt(C.byref(self.c), x.shape[0], p(x), p(vpar), p(vperp),
                                                       float(t_step), p(binit), p(ind_tetr), p(iface), p(t_remain_out),
                                                       p(n_pushes), p(optional_quantities), int(nthreads)))

    def orbit_timestep_trace(self, x, vpar, vperp, t_step, binit, ind_tetr, iface, trace_cap: int):
        """Per-particle call recording the visited (ind_tetr, iface) sequence; arrays updated in place.
        Returns dict(trace_tetr, trace_face, n_pushes, t_remain, fallback, solver_iters, solver_calls)."""
        n = x.shape[0]
        tt, tf = np.zeros((n, max(trace_cap, 1)), np.int32), np.zeros((n, max(trace_cap, 1)), np.int32)
        npush, tro = np.zeros(n, np.int64), np.zeros(n)
        fb = np.zeros(4, np.int64)
        optq = np.zeros((n, 4))
        iters = calls = nadapt = 0
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        for i in range(n):
            tr = GorTrace()
            tr.cap = trace_cap
            tr.ind_tetr = tt[i].ctypes.data_as(ip)
            tr.iface = tf[i].ctypes.data_as(ip)
            t_out = C.c_double(0.0)
            rc = self.L.gor_orbit_timestep(
                C.byref(self.c), x[i].ctypes.data_as(dp), vpar[i:i + 1].ctypes.data_as(dp),
                vperp[i:i + 1].ctypes.data_as(dp), float(t_step), binit[i:i + 1].ctypes.data_as(ip),
                ind_tetr[i:i + 1].ctypes.data_as(ip), iface[i:i + 1].ctypes.data_as(ip), C.byref(t_out), C.byref(tr))
            assert rc == 0, rc
            npush[i], tro[i] = tr.n_pushes, t_out.value
            optq[i] = tr.optional_quantities[:]
            nadapt += tr.n_adaptive
            fb += np.array(tr.n_fallback[:], np.int64)
            iters += tr.n_solver_iters
            calls += tr.n_solver_calls
        return dict(trace_tetr=tt, trace_face=tf, n_pushes=npush, t_remain=tro, fallback=fb, solver_iters=iters,
                    solver_calls=calls, optional_quantities=optq, n_adaptive=nadapt)

    def orbit_timestep_events(self, x, vpar, vperp, t_step, binit, ind_tetr, iface, par_adiab_inv, counter_vpar_0,
                              counter_phi_0, cap, poincare_phi_0=True, n_skip_phi_0=1, poincare_vpar_0=True, J_par=True,
                              n_skip_vpar_0=1, full_orbit=False, n_skip_full_orbit=1):
        """Per-particle orbit_timestep with event capture; returns (events[:min(n_events, cap)], n_events, n_pushes)."""
        n = x.shape[0]
        cfg = GorEventSettings(int(poincare_phi_0), n_skip_phi_0, int(poincare_vpar_0), int(J_par), n_skip_vpar_0,
                               int(full_orbit), n_skip_full_orbit, 0)
        ev = np.zeros(cap, EVENT_DTYPE)
        nev = C.c_int64(0)
        npush = np.zeros(n, np.int64)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        for i in range(n):
            tr = GorTrace()
            t_out = C.c_double(0.0)
            rc = self.L.gor_orbit_timestep_events(
                C.byref(self.c), x[i].ctypes.data_as(dp), vpar[i:i + 1].ctypes.data_as(dp),
                vperp[i:i + 1].ctypes.data_as(dp), float(t_step), binit[i:i + 1].ctypes.data_as(ip),
                ind_tetr[i:i + 1].ctypes.data_as(ip), iface[i:i + 1].ctypes.data_as(ip), C.byref(t_out), C.byref(tr),
                C.byref(cfg), par_adiab_inv[i:i + 1].ctypes.data_as(dp), counter_vpar_0[i:i + 1].ctypes.data_as(ip),
                counter_phi_0[i:i + 1].ctypes.data_as(ip), i, ev.ctypes.data_as(C.c_void_p), cap, C.byref(nev))
            assert rc == 0, rc
            npush[i] = tr.n_pushes
        return ev[:min(nev.value, cap)], nev.value, npush

    def find_tetra(self, x, vpar, vperp, sign_t_step=1):
        n = x.shape[0]
        ind, ifc = np.empty(n, np.int32), np.empty(n, np.int32)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        for i in range(n):
            self.L.gor_find_tetra(C.byref(self.c), x[i].ctypes.data_as(dp), float(vpar[i]), float(vperp[i]),
                                  ind[i:i + 1].ctypes.data_as(ip), ifc[i:i + 1].ctypes.data_as(ip), sign_t_step)
        return ind, ifc

    def invariants(self, x, vpar, vperp, ind_tetr):
        n = x.shape[0]
        e, p, mu = np.full(n, np.nan), np.full(n, np.nan), np.full(n, np.nan)
        dp = C.POINTER(C.c_double)
        tp = self.mesh.tetra_physics
        for i in range(n):
            it = int(ind_tetr[i])
            if it < 1:
                continue
            z = np.empty(4)
            z[:3] = x[i] - tp[it - 1, 0:3]
            z[3] = vpar[i]
            bmod = self.L.gor_bmod(C.byref(self.c), z.ctypes.data_as(dp), it)
            mu[i] = -0.5 * (vperp[i] * vperp[i]) / bmod
            e[i] = self.L.gor_energy_tot(C.byref(self.c), z.ctypes.data_as(dp), mu[i], it)
            p[i] = self.L.gor_p_phi(C.byref(self.c), float(vpar[i]), z.ctypes.data_as(dp), it)
        return e, p, mu
