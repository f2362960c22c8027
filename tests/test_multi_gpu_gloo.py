"""N > 1 host logic on CPU (gloo, world_size 2): particles shard across ranks with the mesh replicated, no
data-path collective, counters reduced at the end -- the structure bench.py uses with NCCL.  The per-rank
"device" here is the oracle (test infrastructure), so this checks the sharding/reduction logic only."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import workloads
    from gorilla_b200 import build_mesh
    from oracle_binding import OracleMesh
    grid, settings = workloads.analytic_tokamak(10, 10, 10)
    settings.poly_order = 2
    mesh = build_mesh(grid, settings)              # replicated on every rank
    om = OracleMesh(mesh, settings)
    x, vpar, vperp = workloads.particles_cyl(n_total, 7, rmin_frac=0.5, rmax_frac=0.98, energy_ev=3e4)
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world   # contiguous shard [r N/G, (r+1) N/G)
    xs, vs, ws = x[lo:hi].copy(), vpar[lo:hi].copy(), vperp[lo:hi].copy()
    st = workloads.fresh_state(hi - lo)
    pushes = om.orbit_timestep_batch(xs, vs, ws, 1e-4, *st, nthreads=1)
    red = torch.tensor([pushes, int((st[1] == -1).sum()), hi - lo], dtype=torch.int64)
    dist.all_reduce(red, op=dist.ReduceOp.SUM)     # the only collective
    np.save(Path(out_dir) / f"x_{rank}.npy", xs)
    if rank == 0:
        np.save(Path(out_dir) / "reduced.npy", red.numpy())
    dist.destroy_process_group()


def test_sharded_run_equals_single_process(tmp_path, product_lib, oracle_lib):
    import workloads
    from gorilla_b200 import build_mesh
    from oracle_binding import OracleMesh
    n = 64
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    red = np.load(tmp_path / "reduced.npy")
    xs = np.concatenate([np.load(tmp_path / f"x_{r}.npy") for r in range(2)])
    grid, settings = workloads.analytic_tokamak(10, 10, 10)
    settings.poly_order = 2
    om = OracleMesh(build_mesh(grid, settings), settings)
    x, vpar, vperp = workloads.particles_cyl(n, 7, rmin_frac=0.5, rmax_frac=0.98, energy_ev=3e4)
    st = workloads.fresh_state(n)
    pushes = om.orbit_timestep_batch(x, vpar, vperp, 1e-4, *st, nthreads=1)
    assert red[0] == pushes and red[1] == int((st[1] == -1).sum()) and red[2] == n
    assert np.array_equal(xs, x)      # particles are independent: sharding does not change any orbit
    assert red[1] > 0                 # some losses, so the loss counter reduction is exercised
