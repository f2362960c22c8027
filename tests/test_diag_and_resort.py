"""What sits around the push kernels behind the C ABI (round 2): particle re-sorting by tetrahedron index in the library,
the diagnostics reduction (counters with the inner / outer / failed loss split, max and rms drift of energy, magnetic moment
and toroidal momentum -- supporting_functions_mod.f90:279-301,377-408, gorilla_plot_mod.f90:290-294,550,603), the NCCL
communicator entry points, and calls on several streams of one handle."""
import ctypes as C

import numpy as np
import pytest

import workloads
from gorilla_b200 import Gorilla, api, build_mesh
from oracle_binding import OracleMesh


def test_shard_range_is_the_contiguous_partition(product_lib):
    """[r N/G, (r+1) N/G) (BASELINE config 5): shards tile [0, N) without gap or overlap, sizes differ by <= 1."""
    for n in (0, 1, 7, 64, 10_000_000, 10_000_019):
        for g in (1, 2, 3, 4, 8):
            pos = 0
            sizes = []
            for r in range(g):
                first, count = api.shard_range(n, r, g)
                assert first == pos == r * n // g
                pos += count
                sizes.append(count)
            assert pos == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(api.GorillaError):
        api.shard_range(10, 2, 2)


def test_wrappers_refuse_wrong_dtypes(product_lib, small_mesh):
    """an int32 n_pushes, a float32 array or a strided view must be a TypeError, not a heap overflow (checked
    before the library is entered, so no device is needed)."""
    mesh, _, settings = small_mesh
    g = Gorilla.__new__(Gorilla)   # no device: only the argument checks are exercised
    g.mesh, g.settings, g._h = mesh, settings, None
    n = 8
    x, vpar, vperp = workloads.particles_cyl(n, 1)
    b, i, f = workloads.fresh_state(n)
    with pytest.raises(TypeError):
        g.orbit_timestep_gorilla(x, vpar, vperp, 1e-6, b, i, f, n_pushes=np.zeros(n, np.int32))
    with pytest.raises(TypeError):
        g.orbit_timestep_gorilla(x.astype(np.float32), vpar, vperp, 1e-6, b, i, f)
    with pytest.raises(TypeError):
        g.orbit_timestep_gorilla(x, vpar[::1][:4], vperp, 1e-6, b, i, f)
    with pytest.raises(TypeError):
        g.orbit_timestep_gorilla(np.asfortranarray(x), vpar, vperp, 1e-6, b, i, f)
    with pytest.raises(TypeError):
        g.find_tetra(x, vpar.astype(np.float32), vperp)
    with pytest.raises(TypeError):
        g.invariants(x, vpar, vperp, i.astype(np.int64))
    with pytest.raises(TypeError):
        g.orbit_timestep_gorilla_events(x, vpar, vperp, 1e-6, b, i, f, np.zeros(n), np.zeros(n, np.int64),
                                        np.zeros(n, np.int32), 16)


def test_mesh_views_outlive_the_mesh_object(product_lib):
    """the numpy views of a library-built mesh keep the C-side mesh alive."""
    import gc
    grid, settings = workloads.analytic_tokamak(6, 6, 6)
    mesh = build_mesh(grid, settings)
    tp, ref = mesh.tetra_physics, mesh.tetra_physics.copy()
    row = tp[5:7]
    del mesh
    gc.collect()
    junk = [np.random.rand(1 << 16) for _ in range(8)]   # churn the allocator
    assert np.array_equal(tp, ref) and np.array_equal(row, ref[5:7]) and len(junk) == 8


# ---------------------------------------------------------------------------------------------------- GPU
def _dev_state(n, seed, dev, **kw):
    import torch
    x, vpar, vperp = workloads.particles_cyl(n, seed, **kw)
    b, i, f = workloads.fresh_state(n)
    return [torch.from_numpy(a).to(dev) for a in (x, vpar, vperp, b, i, f)], (x, vpar, vperp)


@pytest.mark.gpu
def test_resort_dev_sorts_all_arrays_consistently(small_mesh, cuda_device):
    import torch
    mesh, _, settings = small_mesh
    g = Gorilla(mesh, settings)
    n = 20000
    (xd, vd, wd, bd, it, fd), _ = _dev_state(n, 11, cuda_device, rmax_frac=0.98, energy_ev=3e4)
    g.orbit_timestep_gorilla_dev(xd, vd, wd, 2e-5, bd, it, fd)     # localise + move; some particles get lost
    before = [t.clone() for t in (xd, vd, wd, bd, it, fd)]
    tag = torch.arange(n, dtype=torch.float64, device=cuda_device)   # an extra per-particle array that must follow
    perm = torch.empty(n, dtype=torch.int64, device=cuda_device)
    g.resort_dev(xd, vd, wd, bd, it, fd, extra=(tag,), perm_out=perm)
    torch.cuda.synchronize()
    p = perm.cpu().numpy()
    assert np.array_equal(np.sort(p), np.arange(n))                    # a permutation
    key = it.cpu().numpy().astype(np.int64)
    key[key < 1] = 2 ** 32
    assert np.all(np.diff(key) >= 0) and (key == 2 ** 32).sum() > 0     # sorted by tetrahedron, lost particles last
    for new, old in zip((xd, vd, wd, bd, it, fd), before):
        assert torch.equal(new, old[perm])                             # new[i] = old[perm[i]] for every array
    assert np.array_equal(tag.cpu().numpy(), p.astype(np.float64))
    # pushing the sorted batch gives the same particles the same orbits
    ref = [t.clone() for t in before]
    g.orbit_timestep_gorilla_dev(*ref[:3], 2e-5, *ref[3:])
    g.orbit_timestep_gorilla_dev(xd, vd, wd, 2e-5, bd, it, fd)
    torch.cuda.synchronize()
    for new, old in zip((xd, vd, wd, bd, it, fd), ref):
        assert torch.equal(new, old[perm])
    g.close()


@pytest.mark.gpu
def test_host_resort_returns_identical_results_in_caller_order(small_mesh, cuda_device):
    mesh, _, settings = small_mesh
    out = []
    for on in (False, True):
        g = Gorilla(mesh, settings)
        g.set_host_resort(on)
        n = 12000
        x, vpar, vperp = workloads.particles_cyl(n, 4, rmax_frac=0.97, energy_ev=2e4)
        b, i, f = workloads.fresh_state(n)
        tro, npu = np.zeros(n), np.zeros(n, np.int64)
        for _ in range(2):     # first call localises (unsorted: all ind_tetr = -1), second starts from ind_tetr / iface
            g.orbit_timestep_gorilla(x, vpar, vperp, 1.5e-5, b, i, f, t_remain_out=tro, n_pushes=npu)
        out.append((x, vpar, vperp, b, i, f, tro, npu))
        g.close()
    assert all(np.array_equal(p, q) for p, q in zip(out[0], out[1]))
    assert (out[0][4] == -1).sum() > 0


@pytest.mark.gpu
def test_diag_reduce_matches_numpy_and_splits_losses(cuda_device, product_lib):
    """Counters accumulate over calls since diag_reset; drift statistics equal what numpy forms from the per-particle
    invariants; flux-coordinate losses split into inner (s = sfc_s_min) and outer (s = 1) boundary."""
    import torch
    from pathlib import Path
    nc = Path(__file__).resolve().parent.parent / "data" / "equilibria" / "netcdf_file_for_test.nc"
    grid, settings = workloads.vmec_qi(nc, n1=24, n2=10, n3=16)
    mesh = build_mesh(grid, settings)
    g = Gorilla(mesh, settings)
    n = 6000
    rng = np.random.Generator(np.random.PCG64(5))
    x, vpar, vperp = workloads.particles_vmec_alpha(n, 3)
    x[:, 0] = np.where(rng.random(n) < 0.5, 0.14, 0.93)      # near both boundaries so that both kinds of loss occur
    b, i, f = workloads.fresh_state(n)
    dev = cuda_device
    xd, vd, wd, bd, it, fd = [torch.from_numpy(a).to(dev) for a in (x, vpar, vperp, b, i, f)]
    g.orbit_timestep_gorilla_dev(xd, vd, wd, 0.0, bd, it, fd)
    e0, p0, m0 = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
    g.invariants_dev(xd, vd, wd, it, e0, p0, m0)
    g.diag_reset()
    pushes = 0
    for _ in range(3):
        g.orbit_timestep_gorilla_dev(xd, vd, wd, 2e-5, bd, it, fd)
        c = g.counters()
        pushes += c.n_pushes
        assert c.n_lost == int((it == -1).sum())     # per call: ind_tetr == -1 after the call (lost in it or before it)
    d = g.diag_reduce_dev(xd, vd, wd, it, e0, p0, m0)
    ind = it.cpu().numpy()
    n_lost_now = int((ind == -1).sum())
    assert d.nranks == 1 and d.n_particles == n and d.n_pushes == pushes
    # accumulated since the reset: every particle is counted once, in the call it was lost in
    assert d.n_lost == n_lost_now > 0
    assert d.n_lost == d.n_lost_outer + d.n_lost_inner + d.n_failed
    xs = xd.cpu().numpy()
    gone = ind == -1
    # exit points of particles that left through a boundary face sit on that boundary; the others were removed by the pusher
    on_inner, on_outer = gone & (np.abs(xs[:, 0] - 0.1) < 1e-6), gone & (np.abs(xs[:, 0] - 1.0) < 1e-6)
    assert d.n_lost_inner == int(on_inner.sum()) > 0 and d.n_lost_outer == int(on_outer.sum()) > 0
    assert d.n_failed == int((gone & ~on_inner & ~on_outer).sum())
    # drift statistics against numpy
    e1, p1, m1 = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
    g.invariants_dev(xd, vd, wd, it, e1, p1, m1)
    alive = ~gone
    for now, ref, mx, rms in ((e1, e0, d.max_delta_energy, d.rms_delta_energy), (m1, m0, d.max_delta_perpinv, d.rms_delta_perpinv),
                              (p1, p0, d.max_delta_p_phi, d.rms_delta_p_phi)):
        dd = np.abs(now.cpu().numpy()[alive] / ref.cpu().numpy()[alive] - 1.0)
        assert mx == dd.max()
        assert abs(rms - np.sqrt((dd ** 2).sum() / alive.sum())) <= 1e-12 * max(rms, 1e-300) + 1e-30
    assert d.n_sampled == int(alive.sum())
    assert d.max_delta_energy < 0.1 and d.max_delta_perpinv < 1e-12      # order 2 on a coarse mesh; mu round-trips through vperp
    # the host-pointer variant gives the same numbers
    dh = g.diag_reduce(xd.cpu().numpy(), vd.cpu().numpy(), wd.cpu().numpy(), ind, e0.cpu().numpy(), p0.cpu().numpy(),
                       m0.cpu().numpy())
    import dataclasses as _dc
    for f in _dc.fields(d):
        a, b = getattr(d, f.name), getattr(dh, f.name)
        if f.name.startswith("rms_"):       # sums of squares are accumulated with atomics: order, hence last bits, may differ
            assert abs(a - b) <= 1e-12 * max(abs(a), 1e-300)
        else:
            assert a == b, f.name
    # a reset really resets
    g.diag_reset()
    assert g.diag_reduce_dev(xd, vd, wd, it).n_pushes == 0
    g.close()


@pytest.mark.gpu
def test_single_rank_communicator(small_mesh, cuda_device):
    """gorilla_b200_comm_unique_id / _comm_init / _diag_reduce_dev / _comm_allreduce_f64 / _comm_free through NCCL with one
    rank (the 2- and 8-rank runs are bench.py under torchrun; a second rank needs a second GPU)."""
    import torch
    mesh, _, settings = small_mesh
    g = Gorilla(mesh, settings)
    uid = api.comm_unique_id()
    assert len(uid) == api.COMM_ID_BYTES and any(uid)
    g.comm_init(uid, 0, 1)
    with pytest.raises(api.GorillaError):
        g.comm_init(uid, 0, 1)            # a handle has one communicator
    (xd, vd, wd, bd, it, fd), _ = _dev_state(3000, 2, cuda_device)
    g.orbit_timestep_gorilla_dev(xd, vd, wd, 1e-5, bd, it, fd)
    c = g.counters()
    d = g.diag_reduce_dev(xd, vd, wd, it)
    assert d.nranks == 1 and d.n_pushes == c.n_pushes > 0 and d.n_particles == 3000
    buf = torch.tensor([1.5, -2.0, 7.0], dtype=torch.float64, device=cuda_device)
    for op in ("sum", "max", "min"):
        g.comm_allreduce_f64(buf, op)
    torch.cuda.synchronize()
    assert buf.tolist() == [1.5, -2.0, 7.0]
    g.comm_free()
    g.comm_free()                          # idempotent
    g.close()


@pytest.mark.gpu
def test_calls_on_several_streams_of_one_handle(small_mesh, cuda_device):
    """every call has its own counter block / work-queue cursor, so batches issued back to back on different
    streams of one handle are all pushed exactly once."""
    import torch
    mesh, _, settings = small_mesh
    g = Gorilla(mesh, settings)
    nb, n = 12, 4000       # more batches in flight than the ring has slots
    streams = [torch.cuda.Stream(device=cuda_device) for _ in range(4)]
    batches, refs = [], []
    for k in range(nb):
        (xd, vd, wd, bd, it, fd), host = _dev_state(n, 100 + k, cuda_device)
        batches.append((xd, vd, wd, bd, it, fd))
        refs.append(host)
    torch.cuda.synchronize()
    npd = [torch.zeros(n, dtype=torch.int64, device=cuda_device) for _ in range(nb)]
    for k, bt in enumerate(batches):
        s = streams[k % len(streams)]
        with torch.cuda.stream(s):
            g.orbit_timestep_gorilla_dev(*bt[:3], 1e-5, *bt[3:], n_pushes=npd[k], stream=s.cuda_stream)
    torch.cuda.synchronize()
    g2 = Gorilla(mesh, settings)
    for k in range(nb):
        x, vpar, vperp = (a.copy() for a in refs[k])
        b, i, f = workloads.fresh_state(n)
        npu = np.zeros(n, np.int64)
        g2.orbit_timestep_gorilla(x, vpar, vperp, 1e-5, b, i, f, n_pushes=npu)
        assert np.array_equal(batches[k][0].cpu().numpy(), x) and np.array_equal(batches[k][4].cpu().numpy(), i)
        assert np.array_equal(npd[k].cpu().numpy(), npu) and npu.min() > 0
    g.close()
    g2.close()


@pytest.mark.gpu
def test_periodic_relocation_many_periods_away(cuda_device, product_lib):
    """MODULO for reals is fmod + sign fix in gfortran (exact); start angles several periods away relocate to the
    same bits in the oracle, the host mirror of check_coordinate_domain and the device."""
    grid, settings = workloads.analytic_tokamak(10, 10, 10)
    settings.boole_periodic_relocation = True
    settings.poly_order = 2
    mesh = build_mesh(grid, settings)
    n = 512
    x, vpar, vperp = workloads.particles_cyl(n, 9)
    rng = np.random.Generator(np.random.PCG64(1))
    x[:, 1] += 2 * np.pi * rng.integers(-40, 40, n)
    want = np.fmod(x[:, 1], 2 * np.pi * 1.0)
    want = np.where(want < 0, want + 2 * np.pi, want)
    g, om = Gorilla(mesh, settings), OracleMesh(mesh, settings)
    xa, xb, xc = x.copy(), x.copy(), x.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    om.orbit_timestep_batch(xa, vpar.copy(), vperp.copy(), 0.0, *sa, nthreads=1)
    g.orbit_timestep_gorilla(xb, vpar.copy(), vperp.copy(), 0.0, *sb)
    g.check_coordinate_domain(xc)
    assert np.array_equal(xa[:, 1], want) and np.array_equal(xb[:, 1], want) and np.array_equal(xc[:, 1], want)
    assert np.array_equal(sa[1], sb[1]) and (sb[1] > 0).all()
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pusher", ["poly2", "rk4"])
def test_gather_and_prefetch_modes_give_identical_results(small_mesh, small_mesh_phi, cuda_device, pusher):
    """gorilla_b200_set_gather (vector loads, per-lane bulk copies or warp-cooperative copies one push ahead) and gorilla_b200_set_prefetch only change
    how a record reaches the lane: every particle, trace and counter is identical -- including lanes whose predicted exit
    face was wrong (fall-back pushes), lost particles and refills."""
    import dataclasses
    for mesh, _, settings in (small_mesh, small_mesh_phi):
        st = dataclasses.replace(settings, ipusher=1) if pusher == "rk4" else dataclasses.replace(settings, ipusher=2, poly_order=2)
        out = []
        for gather, prefetch in ((0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (2, 1)):
            g = Gorilla(mesh, st)
            g.set_gather(gather)
            g.set_prefetch(prefetch)
            n = 6000
            x, vpar, vperp = workloads.particles_cyl(n, 21, rmax_frac=0.98, energy_ev=2e4)
            b, i, f = workloads.fresh_state(n)
            tro, npu = np.zeros(n), np.zeros(n, np.int64)
            traces = []
            for _ in range(2):
                traces.append(g.orbit_timestep_gorilla(x, vpar, vperp, 1.5e-5, b, i, f, t_remain_out=tro, n_pushes=npu, trace_cap=48))
            c = g.counters()
            out.append((x, vpar, vperp, b, i, f, tro, npu, traces[0][0], traces[1][0], traces[1][1],
                        np.array([c.n_pushes, c.n_lost, c.n_finished, *c.n_fallback])))
            g.close()
        for other in out[1:]:
            assert all(np.array_equal(p, q) for p, q in zip(out[0], other))
        assert (out[0][4] == -1).sum() > 0 and out[0][11][0] > 0
