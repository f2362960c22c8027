"""Shared fixtures.  `-m "not gpu"` runs everything that does not need a device (oracle, host logic,
host compile of the device headers, ABI surface); `-m gpu` runs the CUDA parity tests through the C ABI."""
from __future__ import annotations

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def product_lib():
    """libgorilla_b200.so -- built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from gorilla_b200 import api, build
    import torch
    if not (api._LIB_PATH.exists() and torch.cuda.is_available()):
        build.build()  # no-op when up to date; on the GPU box the prebuilt library travels with the snapshot
    return api.load_library()


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle_binding
    return oracle_binding.load_oracle()


@pytest.fixture(scope="session")
def host_mirror_lib():
    import host_mirror_binding
    return host_mirror_binding.load()


@pytest.fixture(scope="session")
def small_mesh(product_lib):
    """Analytic circular tokamak (EXAMPLES/example_8 geometry) on a 20x20x20 grid: 48 000 tetrahedra."""
    import workloads
    from gorilla_b200 import build_mesh
    grid, settings = workloads.analytic_tokamak(20, 20, 20)
    return build_mesh(grid, settings), grid, settings


@pytest.fixture(scope="session")
def small_mesh_phi(product_lib):
    """Same geometry with an electrostatic potential (eps_Phi != 0) so the Phi sub-record is live."""
    import workloads
    from gorilla_b200 import build_mesh
    grid, settings = workloads.analytic_tokamak(16, 16, 16)
    settings.eps_Phi = -1.0e-7
    return build_mesh(grid, settings), grid, settings


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
