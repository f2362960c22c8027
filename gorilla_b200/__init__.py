"""gorilla_b200 -- B200-native guiding-centre orbit pusher behind GORILLA's orbit_timestep_gorilla API.

The package is a thin host-side mirror of the reference's Fortran interface
(SRC/orbit_timestep_gorilla.f90:10: initialize_gorilla, orbit_timestep_gorilla, check_coordinate_domain)
over the C ABI in include/gorilla_b200.h.  All compute runs in hand-written CUDA kernels
(gorilla_b200/csrc); there is no CPU fall-back and importing the compute API fails loudly when
libgorilla_b200.so has not been built (python -m gorilla_b200.build).
"""
from .settings import GorillaSettings, TetraGridSettings, load_gorilla_inp, load_tetra_grid_inp  # noqa: F401
from .api import (  # noqa: F401
    Gorilla,
    GorillaError,
    Mesh,
    build_mesh,
    load_mesh,
    initialize_gorilla,
    launch_count,
)

__all__ = [
    "Gorilla", "GorillaError", "Mesh", "build_mesh", "load_mesh", "initialize_gorilla", "launch_count",
    "GorillaSettings", "TetraGridSettings", "load_gorilla_inp", "load_tetra_grid_inp",
]
