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
    from gorilla_b200.api import shard_range
    lo, cnt = shard_range(n_total, rank, world)     # gorilla_b200_shard_range: contiguous shard [r N/G, (r+1) N/G)
    hi = lo + cnt
    assert (lo, hi) == (rank * n_total // world, (rank + 1) * n_total // world)
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


# ---------------------------------------------------------------------------------------------- two GPUs, NCCL, through the C ABI
def _gpu_worker(rank, world, uid_bytes, n_total, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import workloads
    from gorilla_b200 import Gorilla, build_mesh
    from gorilla_b200.api import shard_range
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    grid, settings = workloads.analytic_tokamak(10, 10, 10)
    settings.poly_order = 2
    torch.zeros(1, device=dev)                    # CUDA context on this rank's device: the handle binds to the current device
    g = Gorilla(build_mesh(grid, settings), settings)
    g.comm_init(uid_bytes, rank, world)            # the library's own NCCL communicator; no torch.distributed here at all
    x, vpar, vperp = workloads.particles_cyl(n_total, 7, rmin_frac=0.5, rmax_frac=0.98, energy_ev=3e4)
    lo, cnt = shard_range(n_total, rank, world)
    hi = lo + cnt
    st = workloads.fresh_state(hi - lo)
    xd, vd, wd, bd, it, fd = [torch.from_numpy(np.ascontiguousarray(a)).to(dev)
                              for a in (x[lo:hi], vpar[lo:hi], vperp[lo:hi], *st)]
    g.orbit_timestep_gorilla_dev(xd, vd, wd, 0.0, bd, it, fd)
    e0, p0, m0 = (torch.empty(hi - lo, dtype=torch.float64, device=dev) for _ in range(3))
    g.invariants_dev(xd, vd, wd, it, e0, p0, m0)
    g.diag_reset()
    g.orbit_timestep_gorilla_dev(xd, vd, wd, 1e-4, bd, it, fd)
    d = g.diag_reduce_dev(xd, vd, wd, it, e0, p0, m0)       # device reduction + ncclAllReduce inside the library
    buf = torch.tensor([float(rank + 1), -float(rank)], dtype=torch.float64, device=dev)
    g.comm_allreduce_f64(buf, "sum")
    torch.cuda.synchronize()
    np.save(Path(out_dir) / f"gx_{rank}.npy", xd.cpu().numpy())
    np.save(Path(out_dir) / f"gdiag_{rank}.npy", np.array([d.nranks, d.n_particles, d.n_pushes, d.n_lost, d.n_sampled,
                                                           d.max_delta_energy, *buf.tolist()], dtype=np.float64))
    g.comm_free()
    g.close()


@pytest.mark.gpu
def test_two_ranks_reduce_through_the_library_communicator(tmp_path, product_lib, oracle_lib):
    """a 2-rank run whose only collective is the C symbol (gorilla_b200_comm_init + _diag_reduce_dev)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import workloads
    from gorilla_b200 import build_mesh
    from gorilla_b200.api import comm_unique_id
    from oracle_binding import OracleMesh
    n = 4000
    uid = comm_unique_id()
    mp.spawn(_gpu_worker, args=(2, uid, n, str(tmp_path)), nprocs=2, join=True)
    grid, settings = workloads.analytic_tokamak(10, 10, 10)
    settings.poly_order = 2
    om = OracleMesh(build_mesh(grid, settings), settings)
    x, vpar, vperp = workloads.particles_cyl(n, 7, rmin_frac=0.5, rmax_frac=0.98, energy_ev=3e4)
    st = workloads.fresh_state(n)
    pushes = om.orbit_timestep_batch(x, vpar, vperp, 1e-4, *st)
    xs = np.concatenate([np.load(tmp_path / f"gx_{r}.npy") for r in range(2)])
    assert np.array_equal(xs, x)
    d0, d1 = np.load(tmp_path / "gdiag_0.npy"), np.load(tmp_path / "gdiag_1.npy")
    assert np.array_equal(d0, d1)                                   # every rank holds the reduced record
    lost = int((st[1] == -1).sum())
    assert d0[0] == 2 and d0[1] == n and d0[2] == pushes and d0[3] == lost and d0[4] == n - lost
    assert d0[6] == 3.0 and d0[7] == -1.0
