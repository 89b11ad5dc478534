"""CPU tests: host logic of the drop-in layer, the C-ABI surface, and the N>1 sharding path
(world_size 2 and 3, gloo) — no kernel launches."""
import ctypes
import functools
import os
import re

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import reference_vectors as rv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ C-ABI surface
def test_library_exports_every_declared_symbol():
    from dynlab_b200 import _native

    header = open(os.path.join(ROOT, "include", "dynlab_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dlb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 14
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dynlab_b200.h but not exported"
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    assert _native.lib.dlb_abi_version() == 2


def test_flow_registry_matches_abi():
    from dynlab_b200 import _native, flows

    assert len(flows.FLOW_REGISTRY) == 15
    for name, (fid, dim, pnames) in flows.FLOW_REGISTRY.items():
        assert _native.lib.dlb_flow_name(fid).decode() == name
        assert _native.lib.dlb_flow_dim(fid) == dim
        assert _native.lib.dlb_flow_nparams(fid) == len(pnames)
    assert _native.lib.dlb_flow_dim(99) < 0
    assert _native.lib.dlb_workspace_bytes(8192, 4096) >= 8 * 4 * (8192 + 4096)


def test_abi_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call."""
    from dynlab_b200 import _native as nat

    p = nat.darray([0.1, 0.2])
    x = np.linspace(0, 1, 4)
    rc = nat.lib.dlb_flowmap_rk45(0, p, 2, nat.dptr(x), 4, nat.dptr(x), 0, 4, 0.0, 1.0, 1e-6, 1e-6,
                                  np.inf, 0.0, 8, None, None, 8, 8, 1 << 20, None)
    assert rc == nat.DLB_ERR_ARG and b"parameter count" in nat.lib.dlb_last_error()
    with pytest.raises(ValueError):
        nat.check(rc)
    rc = nat.lib.dlb_ftle_stencil(8, 4, 0, nat.dptr(x), 4, nat.dptr(x), 4, 3, 1.0, 0, 4, 8, None, 0,
                                  8, 1 << 20, None)
    assert rc == nat.DLB_ERR_ARG and b"edge_order" in nat.lib.dlb_last_error()
    rc = nat.lib.dlb_ftle_stencil(8, 2, 0, nat.dptr(x), 2, nat.dptr(x), 2, 2, 1.0, 0, 2, 8, None, 0,
                                  8, 1 << 20, None)
    assert rc == nat.DLB_ERR_TOO_SMALL


def test_peer_abi_argument_errors_without_gpu():
    """The multi-GPU helpers validate their arguments before any CUDA call."""
    import ctypes as C

    from dynlab_b200 import _native as nat

    lib = nat.lib
    assert lib.dlb_device_count(None) == nat.DLB_ERR_ARG
    assert lib.dlb_peer_alloc(16, None) == nat.DLB_ERR_ARG
    assert lib.dlb_ipc_export(None, None) == nat.DLB_ERR_ARG
    assert lib.dlb_ipc_open(None, None) == nat.DLB_ERR_ARG
    assert lib.dlb_peer_copy(None, None, 8, None) == nat.DLB_ERR_ARG
    assert lib.dlb_peer_copy(None, None, 0, None) == nat.DLB_OK  # nothing to copy
    dst = (C.c_void_p * 1)(None)
    assert lib.dlb_gather_rows(None, 8, dst, 1, dst, 1, None) == nat.DLB_ERR_ARG
    assert lib.dlb_gather_rows(None, 0, dst, nat.MAX_PEERS + 1, dst, 1, None) == nat.DLB_ERR_ARG
    assert lib.dlb_gather_rows(None, 0, dst, 1, dst, 1, None) == nat.DLB_OK  # no rows, no flags
    assert lib.dlb_halo_exchange(None, 8, 0, 0, None, None, None, None, 1, None) == nat.DLB_ERR_ARG
    assert lib.dlb_wait_flags(None, 0, 1, 1.0, None, None) == nat.DLB_OK
    assert lib.dlb_wait_flags(None, 1, 1, 1.0, None, None) == nat.DLB_ERR_ARG
    assert lib.dlb_wait_flags(8, 33, 1, 1.0, None, None) == nat.DLB_ERR_ARG
    assert b"dlb_wait_flags" in lib.dlb_last_error()
    x = np.linspace(0, 1, 4)
    rc = lib.dlb_ftle_stencil_gather(None, 4, 0, nat.dptr(x), 4, nat.dptr(x), 4, 2, 1.0, 0, 4, None,
                                     None, 0, None, 0, 1, 8, 1 << 20, None)
    assert rc == nat.DLB_ERR_ARG
    p = nat.darray([0.1, 0.2])
    rc = lib.dlb_flowmap_rk45_resident(0, p, 2, 8, 4, 8, 0, 4, 0.0, 1.0, 1e-6, 1e-6, np.inf, 0.0, 8,
                                       None, None, 8, 0, 4, 8, 1 << 20, None)
    assert rc == nat.DLB_ERR_ARG and b"parameter count" in lib.dlb_last_error()
    rc = lib.dlb_flowmap_rk45_resident(0, p, 2, 8, 4, 8, 2, 4, 0.0, 1.0, 1e-6, 1e-6, np.inf, 0.0, 8,
                                       None, None, 8, 1, 4, 8, 1 << 20, None)  # counted rows 1..4 of 2..5
    assert rc == nat.DLB_ERR_ARG and b"counted rows" in lib.dlb_last_error()


# ------------------------------------------------------------------ flows / resolver / utils
def test_host_flows_reference_vectors():
    from dynlab_b200 import flows

    x, y = np.meshgrid([0, 1, 2], [0, 1])
    for name, (ue, ve) in rv.FLOWS_2D.items():
        u, v = getattr(flows, name)(0, (x, y))
        assert np.allclose(ue, u) and np.allclose(ve, v), name
    X = np.meshgrid([0, 1], [0, 1], [0, 1])
    for name, exp in rv.FLOWS_3D.items():
        for e, o in zip(exp, getattr(flows, name)(0, X)):
            assert np.allclose(e, o), name


def test_host_flows_golden(golden):
    from dynlab_b200 import flows

    g = golden("flows")
    for n, (_, dim, _) in flows.FLOW_REGISTRY.items():
        P = (g["pts3"] if dim == 3 else g["pts2"]) * (2000.0 if n == "bickley_jet" else 1.0)
        for k, t in enumerate(g["ts"]):
            out = np.stack(np.broadcast_arrays(*getattr(flows, n)(t, tuple(P))))
            assert np.array_equal(out, g[n][k]), n


def test_resolver():
    from dynlab_b200 import flows
    from dynlab_b200._resolve import UnsupportedFlowError, resolve_flow

    assert resolve_flow(flows.double_gyre) == ("double_gyre", 0, 2, [0.1, 0.2 * np.pi, 0.25])
    p = functools.partial(functools.partial(flows.double_gyre, A=0.3), e=0.5)
    assert resolve_flow(p)[3] == [0.3, 0.2 * np.pi, 0.5]
    assert resolve_flow(flows.FlowSpec("lorenz", rho=30.0)) == ("lorenz", 14, 3, [10.0, 30.0, 8.0 / 3.0])
    with pytest.raises(UnsupportedFlowError):
        resolve_flow(lambda t, Y: (Y[1], -Y[0]))
    with pytest.raises(UnsupportedFlowError):
        resolve_flow(functools.partial(flows.double_gyre, 0.0))
    with pytest.raises(TypeError):
        resolve_flow(functools.partial(flows.double_gyre, B=1.0))
    with pytest.raises(TypeError):
        flows.FlowSpec("double_gyre", B=1.0)

    # a function living in a module called dynlab.flows with the reference's signature
    ns = {}
    exec("def pendulum(t, Y, A=0.25, gamma=0.0, omega=1.0):\n    return Y[1], 0*Y[0]\n", ns)
    f = ns["pendulum"]
    f.__module__ = "dynlab.flows"
    assert resolve_flow(f) == ("pendulum", 8, 2, [0.25, 0.0, 1.0])


def test_force_eigenvectors_reference_vectors():
    from dynlab_b200.utils import force_eigenvectors2D

    for a, e in rv.FORCE_EIG:
        assert (np.array([[e]]) == force_eigenvectors2D(np.array([[a]]))).all()


def test_integrator_kwargs():
    from dynlab_b200.utils import parse_integrator_kwargs

    kw = parse_integrator_kwargs(dict(rtol=1e-8), 1e-3, 1e-6)
    assert kw["rtol"] == 1e-8 and kw["atol"] == 1e-6 and kw["max_step"] == np.inf
    with pytest.raises(TypeError):
        parse_integrator_kwargs(dict(hmax=1.0), 1e-3, 1e-6)
    with pytest.raises(ValueError):
        parse_integrator_kwargs(dict(max_step=0.0), 1e-3, 1e-6)
    with pytest.raises(NotImplementedError):
        parse_integrator_kwargs(dict(method="LSODA"), 1e-3, 1e-6)


def test_constructor_and_validation_without_gpu():
    from dynlab_b200 import diagnostics as D

    with pytest.raises(ValueError):
        D.FTLE(num_threads=0)
    with pytest.raises(ValueError):
        D.LCS(num_threads=-1)
    assert D.FTLE(num_threads=4).num_threads == 4
    assert set(["FTLE", "LCS", "AttractionRate", "RepulsionRate", "iLES", "TrajectoryRepulsionRate",
                "TrajectoryRepulsionRatio"]) <= set(dir(D))
    if not torch.cuda.is_available():  # the product fails loudly without a device
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            D.FTLE().compute([0, 1, 2], [0, 1], __import__("dynlab_b200").flows.double_gyre, (2, 0))


def test_missing_native_library_fails_loudly():
    """No CPU fallback: without the CUDA extension the package refuses to import its diagnostics."""
    import subprocess
    import sys

    code = "import dynlab_b200.diagnostics"
    env = dict(os.environ, DLB_LIB_PATH="/nonexistent/libdynlab_b200.so")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env=env)
    assert out.returncode != 0
    assert "NativeLibraryMissing" in out.stderr and "no CPU fallback" in out.stderr


def test_install_as_dynlab():
    import subprocess
    import sys

    code = ("import dynlab_b200; dynlab_b200.install_as_dynlab();"
            "from dynlab.diagnostics import FTLE, LCS, AttractionRate, RepulsionRate, iLES, "
            "TrajectoryRepulsionRate, TrajectoryRepulsionRatio;"
            "from dynlab.flows import double_gyre, lorenz; from dynlab.utils import odeint_wrapper, "
            "force_eigenvectors2D; import dynlab; print(dynlab.__name__, FTLE.__module__)")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    assert "dynlab_b200 dynlab_b200.diagnostics.lagrangian" in out.stdout


# ------------------------------------------------------------------ sharding (gloo)
def test_partition_covers_grid():
    from dynlab_b200.sharding import active_ranks, partition

    for ydim in (2, 3, 5, 16, 101, 4096):
        for size in (1, 2, 3, 8):
            slabs = [partition(ydim, size, r) for r in range(size)]
            n = active_ranks(ydim, size)
            assert sum(s.rows for s in slabs) == ydim
            rows = 0
            for r, s in enumerate(slabs):
                if r < n:
                    assert s.row0 == rows and s.rows >= min(2, ydim)
                    assert s.lo == max(s.row0 - 1, 0) and s.hi == min(s.row0 + s.rows + 1, ydim)
                    rows += s.rows
                else:
                    assert s.rows == 0
            assert max(s.rows for s in slabs) == slabs[0].rows_max


def _shard_worker(rank, size, port, ydim, xdim, edge_order, out_dir):
    """Each rank owns a row slab of a flow map, exchanges halos, runs the ORACLE stencil on its
    local rows (test infrastructure standing in for K2) and gathers; rank 0 checks the result
    against the unsharded oracle."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from dynlab_b200 import sharding
        from oracle import py_oracle as po

        rng = np.random.default_rng(3)
        x = np.sort(rng.uniform(0, 2, xdim))
        y = np.linspace(0, 1, ydim)
        fm = rng.normal(size=(ydim, xdim, 2))
        full = np.ma.getdata(po.ftle_from_flow_map(fm, x, y, (0, 2), edge_order))

        assert sharding.world() == (rank, size)
        slab = sharding.partition(ydim, size, rank)
        local = torch.zeros(max(slab.nloc, 1), xdim, 2, dtype=torch.float64)
        if slab.rows:  # "K1": own rows only
            local[slab.halo_lo:slab.halo_lo + slab.rows] = torch.from_numpy(fm[slab.row0:slab.row0 + slab.rows])
        sharding.exchange_halo(local, slab, ydim, rank, size)
        if slab.rows:
            assert np.array_equal(local[:slab.nloc].numpy(), fm[slab.lo:slab.hi])
        out = torch.zeros(slab.rows_max, xdim, dtype=torch.float64)
        if slab.rows:
            # local stencil: one-sided formulas only at global edges -> evaluate the oracle on
            # the local rows and keep the rows whose stencil lay inside
            loc = np.ma.getdata(po.ftle_from_flow_map(local[:slab.nloc].numpy(), x, y[slab.lo:slab.hi], (0, 2), edge_order))
            out[:slab.rows] = torch.from_numpy(loc[slab.halo_lo:slab.halo_lo + slab.rows])
        got = sharding.gather_rows(out, ydim, size).numpy()
        assert got.shape == full.shape
        # gather_slab: the same through the entry the diagnostics use (CPU tensors -> the
        # torch.distributed branch; on CUDA the peer-memory branch, checked by the gpu tests)
        part = out[:slab.rows] if slab.rows else None
        got2 = sharding.gather_slab(part, slab, ydim, (xdim,), torch.float64, "cpu").numpy()
        assert np.array_equal(got2, got)
        m8 = torch.full((max(slab.rows, 1), xdim), rank + 1, dtype=torch.uint8)
        gm = sharding.gather_slab(m8 if slab.rows else None, slab, ydim, (xdim,), torch.uint8, "cpu")
        for r in range(size):
            sr = sharding.partition(ydim, size, r)
            assert (gm[sr.row0:sr.row0 + sr.rows] == r + 1).all()
        # rows next to an interior cut used a one-sided oracle formula locally only if the slab
        # had no halo there — with halos every own row matches the unsharded result
        np.testing.assert_allclose(got, full, rtol=1e-12, atol=1e-12)
        sharding.set_distributed(False)
        assert sharding.world() == (0, 1)
        sharding.set_distributed(True)
        if rank == 0:
            open(os.path.join(out_dir, f"ok_{size}_{ydim}_{edge_order}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size,ydim,edge_order", [(2, 16, 2), (2, 7, 1), (3, 11, 2), (2, 3, 2)])
def test_sharded_halo_and_gather_gloo(tmp_path, size, ydim, edge_order):
    port = 29500 + (os.getpid() + size * 17 + ydim) % 2000
    mp.spawn(_shard_worker, args=(size, port, ydim, 9, edge_order, str(tmp_path)), nprocs=size, join=True)
    assert os.path.exists(tmp_path / f"ok_{size}_{ydim}_{edge_order}")


def _shared_worker(rank, size, port, out_dir):
    """sharding.shared_host_array: one segment mapped by every rank; each rank writes its own
    rows, after a barrier every rank sees the whole array; the segment is cached by size and
    its /dev/shm name is gone as soon as it is mapped."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from dynlab_b200 import sharding

        assert sharding.host_result_mode() == "shared"  # the default
        ydim, xdim = 11, 7
        for rep in range(2):
            arr, ten = sharding.shared_host_array((ydim, xdim))
            assert arr.shape == (ydim, xdim) and arr.dtype == np.float64
            slab = sharding.partition(ydim, size, rank)
            rows = torch.arange(slab.row0, slab.row0 + slab.rows, dtype=torch.float64)
            ten[slab.row0:slab.row0 + slab.rows] = (rows[:, None] * 100 + torch.arange(xdim) + rep)
            dist.barrier()
            exp = np.arange(ydim)[:, None] * 100.0 + np.arange(xdim) + rep
            assert np.array_equal(arr, exp)
            dist.barrier()
        assert len(sharding._segments) == 1
        assert not [n for n in os.listdir("/dev/shm") if n.startswith("dynlab_b200_")]
        # host_barrier: nobody leaves before everybody has arrived, generation after generation
        import time
        for gen in range(5):
            if rank == gen % size:
                time.sleep(0.05)
            arr[0, 0] = -1.0 if rank == 0 else arr[0, 0]
            stamp, _ = sharding.shared_host_array((size,), tag="stamps")
            stamp[rank] = gen + 1
            sharding.host_barrier()
            assert (stamp >= gen + 1).all()
            sharding.host_barrier()
        # SharedResultPool: a segment is handed out again only when NO rank holds a result in it
        pool = sharding.SharedResultPool.get((5, 4))
        r1, _ = pool.acquire()
        r1[:] = 1.0
        view = r1[1:3]                      # a derived view keeps the segment busy too
        r2, _ = pool.acquire()
        r2[:] = 2.0
        assert (r1 == 1.0).all() and len(pool.owners) == 2
        del r1
        r3, _ = pool.acquire()              # r1's segment is still seen through `view`
        r3[:] = 3.0
        assert (view == 1.0).all() and (r2 == 2.0).all() and len(pool.owners) == 3
        del view
        if rank != 0:
            del r2                          # rank 0 keeps r2: busy for everybody
        r4, _ = pool.acquire()              # -> the segment r1 / view lived in
        r4[:] = 4.0
        assert len(pool.owners) == 3 and (r3 == 3.0).all()
        dist.barrier()
        if rank == 0:
            assert (r2 == 2.0).all()
        dist.barrier()
        del r3, r4
        if rank == 0:
            del r2
        n_before = len(sharding._segments)
        other, _ = sharding.shared_host_array((3, 5))
        assert other.shape == (3, 5) and len(sharding._segments) == n_before + 1
        sharding.set_host_result("private")
        assert sharding.host_result_mode() == "private"
        sharding.set_host_result("shared")
        with pytest.raises(ValueError):
            sharding.set_host_result("everyone")
        # rotating result segments are distinct mappings
        a0, _ = sharding.shared_host_array((4, 4), tag=("result", 0))
        a1, _ = sharding.shared_host_array((4, 4), tag=("result", 1))
        a0[:] = 1.0
        a1[:] = 2.0
        assert (a0 == 1.0).all() and (a1 == 2.0).all()
        if rank == 0:
            open(os.path.join(out_dir, "ok_shared"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size", [2, 3])
def test_shared_host_segment_gloo(tmp_path, size):
    port = 31500 + (os.getpid() + size * 13) % 2000
    mp.spawn(_shared_worker, args=(size, port, str(tmp_path)), nprocs=size, join=True)
    assert os.path.exists(tmp_path / "ok_shared")


def test_sharding_mode_switches_validate():
    from dynlab_b200 import sharding

    assert sharding.halo_mode() == "redundant"
    sharding.set_halo_mode("exchange")
    assert sharding.halo_mode() == "exchange"
    sharding.set_halo_mode("redundant")
    with pytest.raises(ValueError):
        sharding.set_halo_mode("both")
    with pytest.raises(ValueError):
        sharding.set_host_result("nobody")
    assert sharding.host_result_mode() == "shared"
    sharding.set_host_result("all")  # round-1 name of the private-copy mode
    assert sharding.host_result_mode() == "private"
    sharding.set_host_result("shared")
    assert sharding.transport() == "peer"
    sharding.set_transport("nccl")
    assert sharding.transport() == "nccl"
    sharding.set_transport("peer")
    with pytest.raises(ValueError):
        sharding.set_transport("mpi")


def test_single_process_device_selection(monkeypatch):
    """peer.set_devices / the automatic choice of a plain process (no kernel launches): explicit
    lists are taken as they are, "auto" shards large grids over every visible device and never
    inside a torchrun rank, bad specs raise."""
    from dynlab_b200 import peer

    try:
        monkeypatch.setattr(peer.torch.cuda, "device_count", lambda: 8)
        monkeypatch.delenv("WORLD_SIZE", raising=False)
        peer.set_devices("auto")
        assert peer._single_process_devices(peer.AUTO_MIN_SEEDS) == list(range(8))
        assert peer._single_process_devices(peer.AUTO_MIN_SEEDS - 1) is None
        assert peer._single_process_devices(None) is None
        monkeypatch.setenv("WORLD_SIZE", "8")     # a torchrun rank owns one device
        assert peer._single_process_devices(1 << 30) is None
        monkeypatch.delenv("WORLD_SIZE")
        peer.set_devices("all")
        assert peer._single_process_devices(4) == list(range(8))
        peer.set_devices(3)
        assert peer._single_process_devices(4) == [0, 1, 2]
        peer.set_devices([0, 0, 2])
        assert peer._single_process_devices(4) == [0, 0, 2]
        peer.set_devices(None)
        assert peer.devices_spec() == 1 and peer._single_process_devices(1 << 30) is None
        peer.set_devices(9)
        with pytest.raises(ValueError):
            peer._single_process_devices(4)
        for bad in (0, []):
            with pytest.raises(ValueError):
                peer.set_devices(bad)
        monkeypatch.setattr(peer.torch.cuda, "device_count", lambda: 1)
        peer.set_devices("auto")
        assert peer._single_process_devices(1 << 30) is None
    finally:
        peer.set_devices("auto")


def test_weighted_bounds():
    """Cost-weighted slabs: cover the grid, keep >= 2 rows per active rank, equalise the cost,
    reduce to the uniform partition for a flat cost, and agree with `partition` through
    `slab_from_bounds`."""
    from dynlab_b200 import sharding as sh

    rng = np.random.default_rng(5)
    for ydim in (2, 3, 5, 16, 101, 4096):
        for size in (1, 2, 3, 8):
            ub = sh.uniform_bounds(ydim, size)
            assert [sh.slab_from_bounds(ub, ydim, r) for r in range(size)] == \
                   [sh.partition(ydim, size, r) for r in range(size)]
            n = sh.active_ranks(ydim, size)
            for cost in (rng.uniform(0.5, 2.0, ydim), np.linspace(1, 30, ydim),
                         np.r_[np.zeros(ydim // 2), np.ones(ydim - ydim // 2)]):
                wb = sh.weighted_bounds(ydim, size, cost)
                assert len(wb) == size + 1 and wb[0] == 0 and wb[-1] == ydim
                assert all(b1 - b0 >= min(2, ydim) for b0, b1 in zip(wb[:n], wb[1:n + 1]))
                assert all(b == ydim for b in wb[n:])
                slabs = [sh.slab_from_bounds(wb, ydim, r) for r in range(size)]
                assert sum(s.rows for s in slabs) == ydim
                for r, s in enumerate(slabs[:n]):
                    assert s.lo == max(s.row0 - 1, 0) and s.hi == min(s.row0 + s.rows + 1, ydim)
            assert sh.weighted_bounds(ydim, size, np.ones(ydim)) == ub or ydim % max(n, 1)
    cost = np.linspace(1, 30, 4096)
    wb = sh.weighted_bounds(4096, 8, cost)
    shares = [cost[a:b].sum() for a, b in zip(wb[:-1], wb[1:])]
    assert max(shares) / np.mean(shares) < 1.01
    assert sh.weighted_bounds(4096, 8, np.full(4096, np.nan)) == sh.uniform_bounds(4096, 8)
    assert sh.weighted_bounds(4096, 8, np.ones(7)) == sh.uniform_bounds(4096, 8)


def test_deferred_attribute_resolves_once():
    """_LazyHost + _Deferred (u, v of the fused Eulerian path): the producer runs once, only when
    an attribute is read, and both attributes share its result."""
    from dynlab_b200.diagnostics._base_classes import _Deferred

    calls = []

    def produce():
        calls.append(1)
        return ("U", "V")

    cache = {}
    du, dv = _Deferred(produce, 0, cache), _Deferred(produce, 1, cache)
    assert not calls
    assert dv.resolve() == "V" and du.resolve() == "U" and len(calls) == 1


def test_ctypes_signatures_match_header_arity():
    """Every prototype of include/dynlab_b200.h has as many parameters as the ctypes binding
    declares (a missing argument would shift everything after it silently)."""
    from dynlab_b200 import _native

    header = open(os.path.join(ROOT, "include", "dynlab_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = re.findall(r"\b(dlb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    assert len(protos) == len(_native.SIGNATURES)
    for name, params in protos:
        params = " ".join(params.split())
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_native.SIGNATURES[name][1]), (name, n, len(_native.SIGNATURES[name][1]))


def test_pipeline_schedule_invariants():
    """FTLE._pipeline's chunk plan (single GPU: the whole grid; multi-GPU shared mode: a slab with
    locally integrated halo rows): chunks tile [lo, hi), output ranges tile [own0, own1), and
    every row a stencil reads has been integrated when its output range is produced."""
    from dynlab_b200.diagnostics.lagrangian import pipeline_schedule
    from dynlab_b200.sharding import partition

    def check(lo, hi, own0, own1, ydim, eo, chunks, frac):
        steps = pipeline_schedule(lo, hi, own0, own1, ydim, eo, chunks, frac)
        assert steps[0][0] == lo and steps[-1][1] == hi
        assert all(a[1] == b[0] for a, b in zip(steps, steps[1:]))      # chunks tile [lo, hi)
        assert all(r1 > r0 for r0, r1, _, _ in steps)
        assert steps[0][2] == own0 and steps[-1][3] == own1
        assert all(a[3] == b[2] for a, b in zip(steps, steps[1:]))      # outputs tile [own0, own1)
        for _, r1, done, upto in steps:
            if upto == done:
                continue
            need_lo, need_hi = max(done - 1, 0), min(upto, ydim - 1)     # neighbours
            if eo == 2 and done == 0:
                need_hi = max(need_hi, min(2, ydim - 1))                  # one-sided, top edge
            if eo == 2 and upto == ydim:
                need_lo = min(need_lo, max(ydim - 3, 0))                  # one-sided, bottom edge
            assert lo <= need_lo and need_hi < r1, (lo, hi, own0, own1, ydim, eo, done, upto, r1)

    for ydim in (3, 4, 5, 63, 64, 255, 256, 257, 1000, 4096):
        for eo in (1, 2):
            for chunks, frac in ((4, 32), (1, 32), (2, 16), (8, 64)):
                check(0, ydim, 0, ydim, ydim, eo, chunks, frac)           # single GPU
                for size in (2, 3, 8):
                    for rank in range(size):
                        s = partition(ydim, size, rank)
                        if s.rows:
                            check(s.lo, s.hi, s.row0, s.row0 + s.rows, ydim, eo, chunks, frac)


def test_member_schedule_invariants():
    """Group path (several GPUs): for uniform and weighted partitions, both halo modes, every own
    output row is produced exactly once, in order, and only from rows that exist locally at that
    moment (integrated chunks; after the exchange also the halo slots)."""
    from dynlab_b200 import sharding as sh
    from dynlab_b200.diagnostics.lagrangian import counted_rows, member_schedule

    def need(a, b, ydim, eo):
        lo, hi = max(a - 1, 0), min(b, ydim - 1)
        if eo == 2 and a == 0:
            hi = max(hi, min(2, ydim - 1))
        if eo == 2 and b == ydim:
            lo = min(lo, max(ydim - 3, 0))
        return lo, hi

    rng = np.random.default_rng(11)
    for ydim in (3, 4, 5, 17, 64, 257, 1000, 4096):
        for size in (2, 3, 8):
            for bounds in (sh.uniform_bounds(ydim, size),
                           sh.weighted_bounds(ydim, size, rng.uniform(0.2, 3.0, ydim))):
                for eo in (1, 2):
                    for exchange in (False, True):
                        for rank in range(size):
                            s = sh.slab_from_bounds(bounds, ydim, rank)
                            steps, tail = member_schedule(s, ydim, eo, exchange, 4, 32)
                            if not s.rows:
                                assert steps == [] and tail == []
                                continue
                            k0, k1 = (s.row0, s.row0 + s.rows) if exchange else (s.lo, s.hi)
                            assert steps[0][0] == k0 and steps[-1][1] == k1
                            assert all(a[1] == b[0] for a, b in zip(steps, steps[1:]))
                            produced = []
                            for _, r1, done, upto in steps:
                                if upto > done:
                                    lo, hi = need(done, upto, ydim, eo)
                                    assert k0 <= lo and hi < r1
                                    produced.append((done, upto))
                            if not exchange:
                                assert tail == []
                            for a, b in tail:  # after the exchange the whole [s.lo, s.hi) is local
                                lo, hi = need(a, b, ydim, eo)
                                assert s.lo <= lo and hi < s.hi
                                produced.append((a, b))
                            rows = sorted(r for a, b in produced for r in range(a, b))
                            assert rows == list(range(s.row0, s.row0 + s.rows)), (ydim, size, rank, eo, exchange)
                            # integrator counters: every own row counted exactly once, halo
                            # rows never, and the counted range lies inside its chunk
                            counted = []
                            for r0, r1, _, _ in steps:
                                c0, cn = counted_rows(r0, r1, s)
                                assert r0 <= c0 and c0 + cn <= r1 and cn >= 0
                                counted += list(range(c0, c0 + cn))
                            assert counted == list(range(s.row0, s.row0 + s.rows))


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line
    with the contract's keys; it needs no GPU."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "seeds/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("FTLE seeds/sec") and line["value"] > 0
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"]
    assert abs(cb["value"] - line["value"]) <= 1e-9 * line["value"]
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert e2e["unit"] == line["unit"] and abs(e2e["value"] - line["value"]) <= 1e-9 * line["value"]


# ------------------------------------------------------------------ constants of the CUDA sources
def _c_double_table(path, name):
    """Values of `__constant__ double NAME[...] = { ... };` in a CUDA source, evaluated like C
    (hex floats, integer-looking literals with a decimal point, a / b quotients)."""
    src = open(path).read()
    body = re.search(name + r"\[\d+\]\s*=\s*\{(.*?)\};", src, flags=re.S).group(1)
    body = re.sub(r"//[^\n]*", "", body)
    out = []
    for tok in body.split(","):
        tok = tok.strip()
        if not tok:
            continue
        if "0x" in tok:
            out.append(float.fromhex(tok))
        elif "/" in tok:
            a, b = tok.split("/")
            out.append(float(a) / float(b))
        else:
            out.append(float(tok))
    return out


def test_rk45_tableau_in_cuda_source_is_scipys():
    """csrc/rk45.cu states the Dormand-Prince tableau as literals: it must be scipy's RK45
    (scipy/integrate/_ivp/rk.py:538-552), value for value."""
    from scipy.integrate._ivp.rk import RK45

    t = _c_double_table(os.path.join(ROOT, "dynlab_b200", "csrc", "rk45.cu"), "RKT")
    A, B, E, Cn = RK45.A, RK45.B, RK45.E, RK45.C
    a_flat = [A[i][j] for i in range(1, 6) for j in range(i)]
    assert t[:15] == a_flat
    assert t[15:20] == [B[0], B[2], B[3], B[4], B[5]] and B[1] == 0
    assert t[20:26] == [E[0], E[2], E[3], E[4], E[5], E[6]] and E[1] == 0
    assert t[26:30] == list(Cn[1:5]) and Cn[5] == 1.0


def test_device_trig_constants_and_halfturn_sine_model():
    """csrc/dmath.cuh: pi/2 and pi splits, 2/pi, 1/pi, and the Taylor coefficients of the
    half-turn sine; a float64 model of that sine (the kernel's operations, FMAs emulated in
    long double) stays within 2 ulp of sin on the stage-time arguments of an integration."""
    import math

    K = _c_double_table(os.path.join(ROOT, "dynlab_b200", "csrc", "dmath.cuh"), "DM_K")
    ld = np.longdouble
    assert K[0] == 2 / math.pi and K[30] == 1 / math.pi and K[18] == math.pi
    assert K[1] == math.pi / 2 and K[4] == 1.5 * 2.0 ** 52
    # pi/2 = P1 + P2 + P3 and pi = K18 + K19: the splits are the rounding residues of the previous terms
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 400
    assert K[2] == float(mp.pi / 2 - mp.mpf(K[1]))
    assert K[3] == float(mp.pi / 2 - mp.mpf(K[1]) - mp.mpf(K[2]))
    assert K[19] == 2 * K[2]
    for i in range(10):
        assert K[20 + i] == (-1) ** (i + 1) / math.factorial(2 * i + 3)

    def fma(a, b, c):
        return (ld(a) * ld(b) + ld(c)).astype(np.float64)

    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-7, 7, 50000), rng.uniform(-1e-3, 1e-3, 5000),
                        math.pi * np.arange(-2, 3) + rng.uniform(-1e-9, 1e-9, 5)])
    t = fma(x, K[30], K[4])
    q = t.view(np.int64) & 1
    fq = t - K[4]
    r = fma(-fq, K[19], fma(-fq, K[18], x))
    z = r * r
    p = fma(np.full_like(x, K[29]), z, K[28])
    for k in range(27, 19, -1):
        p = fma(p, z, K[k])
    s = np.where(q == 1, -1.0, 1.0) * fma(r * z, p, r)
    ref = np.sin(ld(x))
    ulp = np.abs(ld(s) - ref) / np.spacing(np.abs(ref.astype(np.float64)))
    # (the long-double stand-in for the FMA keeps 11 fraction bits next to the magic constant, so
    # a quotient within 2.4e-4 of a half-integer may round the other way here: |r| <= pi/2 + 1e-3)
    assert np.abs(r).max() <= math.pi / 2 + 1e-3
    assert float(ulp.max()) < 2.0 and float((ulp > 1).mean()) < 0.02
