"""ctypes binding of tests/_build/libhost_mirror.so (host compile of the device headers; test-only)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "host_mirror" / "host_mirror.cpp"
LIB = HERE / "_build" / "libhost_mirror.so"
CSRC = HERE.parent / "gorilla_b200" / "csrc"


def build_host_mirror(force=False) -> Path:
    LIB.parent.mkdir(exist_ok=True)
    deps = [SRC] + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp"))
    if force or not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(LIB),
                        str(SRC)], check=True, capture_output=True)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        build_host_mirror()
        L = C.CDLL(str(LIB))
        d, vp = C.c_double, C.c_void_p
        L.hm_create.restype = vp
        L.hm_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, d, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_int, d]
        L.hm_free.argtypes = [vp]
        L.hm_has_phi.argtypes = [vp]
        L.hm_orbit_timestep.restype = C.c_int64
        L.hm_orbit_timestep.argtypes = [vp, C.c_int64, vp, vp, vp, d, vp, vp, vp, vp, vp, C.c_int32, vp, vp, C.c_int, vp, vp]
        L.hm_orbit_timestep_events.restype = C.c_int64
        L.hm_orbit_timestep_events.argtypes = [vp, C.c_int64, vp, vp, vp, d, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int,
                                               vp, vp, vp, vp, C.c_int64, vp, C.c_int, C.c_int]
        L.hm_hypot.restype = d
        L.hm_hypot.argtypes = [d, d]
        L.hm_csqrt.argtypes = [d, d, vp]
        L.hm_cdiv.argtypes = [d, d, d, d, vp]
        L.hm_cmul.argtypes = [d, d, d, d, vp]
        L.hm_rmul.argtypes = [d, d, d, vp]
        L.hm_frac_jump_phase.argtypes = [C.c_int, vp]
        L.hm_cmplx_roots_gen.argtypes = [C.c_int, vp, vp]
        for name, nargs in (("hm_quadratic_solver1", 3), ("hm_quadratic_solver2", 3), ("hm_cubic_solver", 4)):
            getattr(L, name).argtypes = [d] * nargs
            getattr(L, name).restype = d
        L.hm_quartic_solver.argtypes = [C.c_int] + [d] * 5
        L.hm_quartic_solver.restype = d
        _lib = L
    return _lib


EVENT_DTYPE = np.dtype([("particle", np.int64), ("kind", np.int32), ("counter", np.int32), ("push", np.int64),
                        ("x", np.float64, 3), ("value", np.float64, 2), ("t", np.float64)])


def oq_mask_of(settings) -> int:
    return (int(bool(settings.boole_time_Hamiltonian)) | 2 * int(bool(settings.boole_gyrophase))
            | 4 * int(bool(settings.boole_vpar_int)) | 8 * int(bool(settings.boole_vpar2_int)))


class HostMirror:
    def __init__(self, mesh, settings):
        self.L = load()
        self.mesh = mesh
        self._desc = mesh.desc()
        self.h = self.L.hm_create(C.byref(self._desc), settings.poly_order, int(settings.boole_guess),
                                  int(settings.boole_periodic_relocation), int(settings.ipusher),
                                  int(settings.boole_strong_electric_field), int(settings.i_time_tracing_option),
                                  oq_mask_of(settings), int(settings.boole_adaptive_time_steps),
                                  float(settings.desired_delta_energy), int(settings.max_n_intermediate_steps),
                                  int(settings.handover_processing_kind), int(settings.i_precomp),
                                  int(settings.boole_newton_precalc), int(settings.boole_pusher_ode45),
                                  float(settings.rel_err_ode45))
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hm_free(self.h)
            self.h = None

    def orbit_timestep(self, x, vpar, vperp, t_step, binit, ind_tetr, iface, trace_cap=0, force_full=False,
                       optional=False):
        n = x.shape[0]
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        tt, tf = np.zeros((n, max(trace_cap, 1)), np.int32), np.zeros((n, max(trace_cap, 1)), np.int32)
        npush, tro, fb = np.zeros(n, np.int64), np.zeros(n), np.zeros(5, np.int64)
        optq = np.zeros((n, 4)) if optional else None
        dom = self.L.hm_orbit_timestep(self.h, n, p(x), p(vpar), p(vperp), float(t_step), p(binit), p(ind_tetr),
                                       p(iface), p(tro), p(npush), trace_cap, p(tt), p(tf), int(force_full), p(fb), p(optq))
        return dict(trace_tetr=tt, trace_face=tf, n_pushes=npush, t_remain=tro, fallback=fb[:4], domain_errors=dom,
                    optional_quantities=optq, n_adaptive=int(fb[4]))

    def orbit_timestep_events(self, x, vpar, vperp, t_step, binit, ind_tetr, iface, par_adiab_inv, counter_vpar_0,
                              counter_phi_0, cap, poincare_phi_0=True, n_skip_phi_0=1, poincare_vpar_0=True, J_par=True,
                              n_skip_vpar_0=1, force_full=False, full_orbit=False, n_skip_full_orbit=1):
        n = x.shape[0]
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        ev = np.zeros(max(cap, 1), EVENT_DTYPE)
        nev = np.zeros(1, np.int64)
        npush, tro = np.zeros(n, np.int64), np.zeros(n)
        flags = int(poincare_phi_0) | 2 * int(poincare_vpar_0) | 4 * int(J_par) | 8 * int(full_orbit)
        rc = self.L.hm_orbit_timestep_events(self.h, n, p(x), p(vpar), p(vperp), float(t_step), p(binit), p(ind_tetr),
                                             p(iface), p(tro), p(npush), flags, n_skip_phi_0, n_skip_vpar_0,
                                             p(par_adiab_inv), p(counter_vpar_0), p(counter_phi_0), p(ev), cap, p(nev),
                                             int(force_full), int(n_skip_full_orbit))
        assert rc == 0
        return ev[:min(int(nev[0]), cap)], int(nev[0]), npush
