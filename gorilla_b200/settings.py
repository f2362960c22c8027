"""Settings of the hot path: the namelists GORILLANML and TETRA_GRID_NML.

Reference: SRC/gorilla_settings_mod.f90:94-150 (namelist + consistency checks),
SRC/tetra_grid_settings_mod.f90:70-105, blueprint values INPUT/gorilla.inp, INPUT/tetra_grid.inp.
Only the entries the hot path reads are kept; unknown namelist keys are ignored on load.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, fields
from pathlib import Path


@dataclass
class GorillaSettings:
    eps_Phi: float = 0.0
    coord_system: int = 2
    ispecies: int = 2
    boole_periodic_relocation: bool = True
    ipusher: int = 2
    boole_pusher_ode45: bool = False
    rel_err_ode45: float = 1.0e-8              # INPUT/gorilla.inp:36
    boole_dt_dtau: bool = True
    boole_newton_precalc: bool = False
    poly_order: int = 2
    i_precomp: int = 0
    boole_guess: bool = True
    i_time_tracing_option: int = 1
    handover_processing_kind: int = 1
    boole_adaptive_time_steps: bool = False
    desired_delta_energy: float = 1.0e-10      # INPUT/gorilla.inp:147
    max_n_intermediate_steps: int = 10000      # INPUT/gorilla.inp:151
    boole_strong_electric_field: bool = False
    boole_grid_for_find_tetra: bool = False
    # optional quantities of pusher_tetra_poly (gorilla_settings_mod.f90:51-55, namelist :100)
    boole_time_Hamiltonian: bool = False
    boole_gyrophase: bool = False
    boole_vpar_int: bool = False
    boole_vpar2_int: bool = False
    # analytical helical perturbation of A_phi at the vertices of a grid_kind 2 mesh (INPUT/gorilla.inp:118-125,
    # tetra_physics_mod.f90:1158-1161); read by build_mesh
    boole_helical_pert: bool = False
    helical_pert_eps_Aphi: float = 1.0e-1
    helical_pert_m_fourier: int = 2
    helical_pert_n_fourier: int = 2
    # random noise on the vertex potentials of a built mesh (INPUT/gorilla.inp:84-109, tetra_physics_mod.f90:400-444); the
    # stream is the library's own (noise_seed; 0 = fixed default), not gfortran's random_number
    boole_axi_noise_vector_pot: bool = False
    axi_noise_eps_A: float = 1.0e-1
    boole_axi_noise_elec_pot: bool = False
    axi_noise_eps_Phi: float = 3.0e-1
    boole_non_axi_noise_vector_pot: bool = False
    non_axi_noise_eps_A: float = 1.0e-4
    noise_seed: int = 0


@dataclass
class TetraGridSettings:
    grid_kind: int = 3
    n1: int = 100
    n2: int = 40
    n3: int = 40
    boole_n_field_periods: bool = True
    n_field_periods_manual: int = 1
    i_radial_spacing: int = 0
    theta_geom_flux: int = 1          # grid_kind 2: poloidal grid equidistant in 1 the flux angle | 2 the geometrical angle
    sfc_s_min: float = 0.1
    theta0_at_xpoint: float = 1.0   # logical in the namelist (.true. = theta = 0 on the axis -> X-point ray)
    R0_analytic_circ: float = 0.0
    a_analytic_circ: float = 0.0
    B0_analytic_circ: float = 0.0
    q0_analytic_circ: float = 1.0
    q1_analytic_circ: float = 0.0
    g_file_filename: str = ""
    convex_wall_filename: str = ""
    netcdf_filename: str = ""
    knots_SOLEDGE3X_EIRENE_filename: str = ""
    triangles_SOLEDGE3X_EIRENE_filename: str = ""
    # not a namelist entry: the optional argument bmod_multiplier of initialize_gorilla (orbit_timestep_gorilla.f90:151)
    bmod_multiplier: float = 1.0
    # field_divB0.inp lines 11-12: moving-average filter windows of the psi(R, Z) table (bdivfree.f90:1144-1164)
    nwindow_r: int = 0
    nwindow_z: int = 0


_ASSIGN = re.compile(r"^\s*([A-Za-z_][A-Za-z0-9_]*)\s*=\s*(.*?)\s*,?\s*$")


def _parse_value(txt: str):
    t = txt.strip().rstrip(",").strip()
    if t.lower() in (".true.", "t", ".t."):
        return True
    if t.lower() in (".false.", "f", ".f."):
        return False
    if (t.startswith("'") and t.endswith("'")) or (t.startswith('"') and t.endswith('"')):
        return t[1:-1]
    try:
        return int(t)
    except ValueError:
        pass
    return float(t.lower().replace("d", "e"))


def parse_namelist(path: str | Path, group: str) -> dict:
    """Minimal Fortran namelist reader (one `key = value ,` per line, `!` comments)."""
    out: dict = {}
    inside = False
    for raw in Path(path).read_text().splitlines():
        line = raw.split("!")[0].strip()
        if not line:
            continue
        if line.lower().startswith("&" + group.lower()):
            inside = True
            continue
        if inside and line.startswith("/"):
            break
        if inside:
            m = _ASSIGN.match(line)
            if m:
                out[m.group(1).lower()] = _parse_value(m.group(2))
    return out


def _fill(cls, values: dict):
    obj = cls()
    lut = {f.name.lower(): f for f in fields(cls)}
    for k, v in values.items():
        f = lut.get(k)
        if f is None:
            continue
        cur = getattr(obj, f.name)
        if isinstance(cur, bool):
            v = bool(v)
        elif isinstance(cur, int) and not isinstance(v, bool):
            v = int(v)
        elif isinstance(cur, float):
            v = float(v)
        setattr(obj, f.name, v)
    return obj


def load_gorilla_inp(path: str | Path = "gorilla.inp") -> GorillaSettings:
    """load_gorilla_inp (gorilla_settings_mod.f90:111-150)."""
    return _fill(GorillaSettings, parse_namelist(path, "GORILLANML"))


def load_tetra_grid_inp(path: str | Path = "tetra_grid.inp") -> TetraGridSettings:
    """load_tetra_grid_inp (tetra_grid_settings_mod.f90:81-105)."""
    return _fill(TetraGridSettings, parse_namelist(path, "TETRA_GRID_NML"))
