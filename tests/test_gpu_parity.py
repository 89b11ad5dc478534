"""GPU parity tests proper: the CUDA path (through the public API and the C-ABI) against
(1) the reference's own known-answer vectors, (2) fixtures generated from the unmodified
reference (tests/golden/), (3) the oracle on the same seeded inputs.

Tolerances (north_star): flow-map endpoints within 10*rtol relative (floor: atol-scale),
FTLE within 1e-6 relative.  Measured agreement is orders of magnitude tighter and the tests
assert the tighter figures where the arithmetic allows (stencil/eigen kernels: 1e-11).
"""
import functools
import warnings

import numpy as np
import pytest

from oracle.make_golden import FTLE_CASES
from tests import reference_vectors as rv
from tests.test_oracle_golden import EUL_CASES, EUL_GRIDS

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dl():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dynlab_b200
    from dynlab_b200 import diagnostics, flows, peer, utils  # noqa: F401

    # the tests below talk to ONE device unless they say otherwise (a plain process on a
    # multi-GPU box would shard every large grid by itself: test_auto_multi_device_default)
    peer.set_devices(None)
    return dynlab_b200


def rel_err(a, b, floor):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


# ---------------------------------------------------------------- K0 / flows
def test_flows_reference_vectors(dl):
    """R:tests/test_flows.py through the device twins (K0, grid mode)."""
    from dynlab_b200._engine import Engine
    from dynlab_b200._resolve import resolve_flow

    eng = Engine.get()
    for name, (ue, ve) in rv.FLOWS_2D.items():
        _, fid, _, params = resolve_flow(getattr(dl.flows, name))
        u, v = eng.flow_eval_grid(fid, params, 0.0, np.array([0., 1., 2.]), np.array([0., 1.]))
        assert np.allclose(ue, u.cpu().numpy()), name
        assert np.allclose(ve, v.cpu().numpy()), name
    X, Y, Z = np.meshgrid([0, 1], [0, 1], [0, 1])
    pts = torch.tensor(np.stack([X.ravel(), Y.ravel(), Z.ravel()]), dtype=torch.float64, device=eng.device)
    for name, exp in rv.FLOWS_3D.items():
        _, fid, dim, params = resolve_flow(getattr(dl.flows, name))
        out = eng.flow_eval_points(fid, dim, params, 0.0, pts).cpu().numpy()
        for k in range(3):
            assert np.allclose(np.array(exp[k]).ravel(), out[k]), name


def test_flows_golden_points(dl, golden):
    """All 15 device twins on 257 random points x 4 times vs the reference's numpy output."""
    from dynlab_b200._engine import Engine
    from dynlab_b200._resolve import resolve_flow

    g = golden("flows")
    eng = Engine.get()
    for name, (fid, dim, _) in dl.flows.FLOW_REGISTRY.items():
        P = (g["pts3"] if dim == 3 else g["pts2"]) * (2000.0 if name == "bickley_jet" else 1.0)
        pts = torch.tensor(P, dtype=torch.float64, device=eng.device)
        _, fid, dim, params = resolve_flow(getattr(dl.flows, name))
        for k, t in enumerate(g["ts"]):
            out = eng.flow_eval_points(fid, dim, params, float(t), pts).cpu().numpy()
            ref = g[name][k]
            scale = np.abs(ref).max()
            assert np.abs(out - ref).max() <= 4e-15 * scale + 1e-300, (name, t)
    _, fid, dim, params = resolve_flow(functools.partial(dl.flows.double_gyre, A=0.25, w=1.3, e=0.1))
    pts = torch.tensor(g["pts2"], dtype=torch.float64, device=eng.device)
    out = eng.flow_eval_points(fid, dim, params, 0.7, pts).cpu().numpy()
    assert np.abs(out - g["double_gyre_params"]).max() < 4e-15
    _, fid, dim, params = resolve_flow(dl.flows.FlowSpec("bickley_jet", scale=4000.0))
    out = eng.flow_eval_points(fid, dim, params, 0.01, pts).cpu().numpy()
    assert np.abs(out - g["bickley_scaled"]).max() < 1e-14 * np.abs(g["bickley_scaled"]).max()


def test_device_trig_accuracy(dl):
    """dm::sin / dm::sincos (csrc/dmath.cuh) against numpy over small, large and
    near-multiple-of-pi arguments, through pendulum (v = -sin x) and abc (sin/cos of x,y,z)."""
    from dynlab_b200._engine import Engine
    from dynlab_b200._resolve import resolve_flow

    eng = Engine.get()
    rng = np.random.default_rng(7)
    k = np.arange(-200, 200)
    xs = np.concatenate([
        rng.uniform(-10, 10, 4000), rng.uniform(-1e5, 1e5, 4000), rng.uniform(-1e9, 1e9, 1000),
        k * (np.pi / 2), k * (np.pi / 2) + 1e-9, np.nextafter(k * np.pi, np.inf),
        [0.0, -0.0, 1e-300, 5e-324, 105614.9, 105615.1]])
    _, fid, dim, params = resolve_flow(functools.partial(dl.flows.pendulum, A=0.0, gamma=0.0))
    pts = torch.tensor(np.stack([xs, np.zeros_like(xs)]), dtype=torch.float64, device=eng.device)
    out = eng.flow_eval_points(fid, dim, params, 0.0, pts).cpu().numpy()
    ref = -np.sin(xs)
    ulp = np.spacing(np.abs(ref))
    assert (np.abs(out[1] - ref) <= 2 * ulp + 1e-300).all()
    _, fid, dim, params = resolve_flow(functools.partial(dl.flows.abc, A=1.0, B=0.0, C=0.0))
    p3 = torch.tensor(np.stack([xs, xs, xs]), dtype=torch.float64, device=eng.device)
    out = eng.flow_eval_points(fid, dim, params, 0.0, p3).cpu().numpy()
    for got, ref in ((out[0], np.sin(xs)), (out[1], np.cos(xs))):
        assert (np.abs(got - ref) <= 2 * np.spacing(np.abs(ref)) + 1e-300).all()


# ---------------------------------------------------------------- reference's own tests
def test_reference_test_ftle(dl):
    """R:tests/test_diagnostics/test_lagrangian.py:6-17, verbatim inputs, default integrator."""
    f = dl.diagnostics.FTLE(num_threads=1).compute(rv.FTLE_X, rv.FTLE_Y, dl.flows.double_gyre, rv.FTLE_T)
    assert isinstance(f, np.ma.MaskedArray) and f.shape == (2, 3)
    assert np.allclose(rv.FTLE_EXPECTED, f)
    assert f[0, 2] == 0 and f[1, 0] == 0  # lambda_max < 1 clamp


def test_reference_test_eulerian(dl):
    """R:tests/test_diagnostics/test_eulerian.py:8-123 (8 tests), verbatim inputs."""
    D, dg = dl.diagnostics, dl.flows.double_gyre
    u, v = dg(0, np.meshgrid(rv.EUL_X, rv.EUL_Y))
    for kw in (dict(f=dg, t=0), dict(u=u, v=v)):
        assert np.allclose(rv.ATTRACTION_EXPECTED, D.AttractionRate().compute(rv.EUL_X, rv.EUL_Y, **kw))
        assert np.allclose(rv.REPULSION_EXPECTED, D.RepulsionRate().compute(rv.EUL_X, rv.EUL_Y, **kw))
    u, v = dg(0, np.meshgrid(rv.TRAJ_X, rv.TRAJ_Y))
    for kw in (dict(f=dg, t=0), dict(u=u, v=v)):
        assert np.allclose(rv.RHODOT_EXPECTED, D.TrajectoryRepulsionRate().compute(rv.TRAJ_X, rv.TRAJ_Y, **kw))
        assert np.allclose(rv.NUDOT_EXPECTED, D.TrajectoryRepulsionRatio().compute(rv.TRAJ_X, rv.TRAJ_Y, **kw))


# ---------------------------------------------------------------- K1 + K2, config 1
def test_ftle_c1_against_reference(dl, golden):
    """BASELINE config 1 in full (R:README.md:24-26) vs the unmodified reference (RK45)."""
    g = golden("ftle_c1")
    d = dl.diagnostics.FTLE(num_threads=1)
    field = d.compute(g["x"], g["y"], dl.flows.double_gyre, (10, 0), edge_order=2, rtol=1e-8, atol=1e-8)
    rtol = 1e-8
    e = rel_err(d.flow_map, g["flow_map"], 1.0)
    assert e.max() < 10 * rtol, e.max()          # north_star bar
    assert e.max() < 1e-10, e.max()               # what the arithmetic actually gives
    ef = rel_err(np.ma.getdata(field), g["field"], 1e-3)
    excluded = (ef > 1e-6).mean()
    assert excluded == 0.0                        # no ridge-cusp exclusions needed
    assert ef.max() < 1e-8, ef.max()
    assert d.stats["nfev"] == int(g["nfev"].sum())  # identical step sequences in total
    assert d.stats["failed"] == 0 and d.stats["seeds"] == 101 * 101


def test_flow_map_per_seed_nfev(dl, golden):
    """Per-seed nfev of K1 equals scipy's on every seed of config 1."""
    from dynlab_b200._engine import Engine
    from dynlab_b200._resolve import resolve_flow

    g = golden("ftle_c1")
    eng = Engine.get()
    _, fid, _, params = resolve_flow(dl.flows.double_gyre)
    nfev = torch.zeros(101, 101, dtype=torch.int32, device=eng.device)
    status = torch.full((101, 101), 7, dtype=torch.int8, device=eng.device)
    fm = eng.flow_map(fid, params, g["x"], g["y"], 0, 101, 10.0, 0.0, 1e-8, 1e-8, nfev=nfev, status=status)
    assert np.array_equal(nfev.cpu().numpy(), g["nfev"])
    assert (status.cpu().numpy() == 0).all()
    assert np.abs(fm.cpu().numpy() - g["flow_map"]).max() < 1e-10


@pytest.mark.parametrize("name", list(FTLE_CASES))
def test_ftle_cases_against_reference(dl, golden, name):
    """Every 2-D flow, both time directions, uniform / non-uniform / mixed grids, edge_order
    1 and 2, loose and tight tolerances — vs fixtures from the unmodified reference."""
    g = golden("ftle_cases")
    fl, kw, x, y, t, eo, tol = FTLE_CASES[name]
    atol = tol if name != "dg_loose" else 1e-6
    f = getattr(dl.flows, fl)
    if kw:
        f = functools.partial(f, **kw)
    d = dl.diagnostics.FTLE()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        field = d.compute(x, y, f, t, edge_order=eo, rtol=tol, atol=atol)
    ref_fm, ref_f, st = g[name + "__flow_map"], g[name + "__field"], g[name + "__status"]
    ok = st == 0
    assert d.stats["failed"] == int((~ok).sum())
    assert bool(w) == bool((~ok).any())  # RuntimeWarning iff scipy reported status -1 seeds
    e = rel_err(d.flow_map, ref_fm, 1.0)[ok]
    assert e.max() < 10 * tol, (name, e.max())
    if tol <= 1e-8:
        assert e.max() < 1e-9, (name, e.max())
    # FTLE: compare where all stencil inputs are regular seeds
    if ok.all():
        ef = rel_err(np.ma.getdata(field), ref_f, 1e-3)
        assert ef.max() < (1e-6 if tol <= 1e-8 else 1e-2), (name, ef.max())
    assert d.stats["nfev"] >= int(g[name + "__nfev"][ok].sum())


def test_ftle_stencil_alone(dl, golden):
    """K2 on the reference's own flow maps: np.gradient both branches, both edge orders."""
    from dynlab_b200._engine import Engine

    eng = Engine.get()
    g = golden("ftle_cases")
    for name, (fl, kw, x, y, t, eo, tol) in FTLE_CASES.items():
        if name in ("glider_blowup",):
            continue
        fm = torch.tensor(g[name + "__flow_map"], device=eng.device)
        x64, y64 = np.ascontiguousarray(x, dtype=float), np.ascontiguousarray(y, dtype=float)
        out, _ = eng.ftle_stencil(fm, 0, x64, y64, eo, abs(t[1] - t[0]), 0, len(y))
        ref = g[name + "__field"]
        # gradient cancellation noise: eps * |f| / dx relative to the gradient scale
        assert rel_err(out.cpu().numpy(), ref, 1e-3).max() < 1e-11, name
    gc = golden("ftle_c1")
    fm = torch.tensor(gc["flow_map"], device=eng.device)
    out, _ = eng.ftle_stencil(fm, 0, gc["x"], gc["y"], 2, 10.0, 0, 101)
    assert rel_err(out.cpu().numpy(), gc["field"], 1e-3).max() < 1e-11


def test_stencils_on_random_grids(dl):
    """K2 / K3 (bulk-copy and register-window interior kernels + edge kernel) against numpy on
    seeded random fields over random shapes (2..300, odd and even), uniform / non-uniform axes
    and both edge orders — tails of strips, row chunks and ring refills included."""
    import os

    from dynlab_b200 import _native as nat
    from dynlab_b200._engine import Engine
    from oracle import py_oracle as po

    eng = Engine.get()
    rng = np.random.default_rng(20261020)
    shapes = [(2, 2), (3, 3), (2, 9), (9, 2), (3, 130), (130, 3), (64, 126), (65, 127), (66, 128), (67, 129),
              (129, 255), (200, 300), (131, 254)]
    shapes += [tuple(int(v) for v in rng.integers(4, 300, 2)) for _ in range(12)]
    try:
        for bulk in ("1", "0"):
            os.environ["DLB_STENCIL_BULK"] = bulk
            for (ny, nx) in shapes:
                ux, uy = bool(rng.integers(2)), bool(rng.integers(2))
                x = np.arange(nx) * 0.37 if ux else np.sort(rng.uniform(0, 10, nx))
                y = np.arange(ny) * 0.5 - 3 if uy else np.sort(rng.uniform(-5, 5, ny))
                eo = int(rng.integers(1, 3))
                if min(nx, ny) < eo + 1:
                    eo = 1
                fm = rng.normal(size=(ny, nx, 2)) * 3
                gs = np.abs(fm).max() / min(np.diff(x).min(), np.diff(y).min())
                out, _ = eng.ftle_stencil(torch.tensor(fm, device=eng.device), 0, x, y, eo, 2.0, 0, ny)
                ref = np.ma.getdata(po.ftle_from_flow_map(fm, x, y, (0, 2.0), eo))
                # ln(lambda) with lambda ~ gs^2: compare in absolute terms scaled by the rounding
                # noise of the gradients (eps * |f| / dx relative to the gradient)
                assert np.abs(out.cpu().numpy() - ref).max() < 1e-9, (bulk, ny, nx, eo, ux, uy)
                u, v = fm[..., 0].copy(), fm[..., 1].copy()
                v[rng.integers(ny), rng.integers(nx)] = 0.0
                u[v == 0.0] = 0.0  # a zero-velocity node: masked by rate / ratio
                ud, vd = torch.tensor(u, device=eng.device), torch.tensor(v, device=eng.device)
                for mode, idx in ((nat.EULER_S_MIN, 0), (nat.EULER_S_MAX, -1)):
                    f, _, _ = eng.euler_stencil(mode, ud, vd, 0, x, y, eo, 0, ny)
                    r = np.ma.getdata(po.attraction_repulsion_rate(x, y, u=u, v=v, edge_order=eo, eigenvalue_index=idx))
                    assert np.abs(f.cpu().numpy() - r).max() < 1e-12 * gs, (bulk, ny, nx, eo, mode)
                for mode, rr in ((nat.EULER_RATE, "rate"), (nat.EULER_RATIO, "ratio")):
                    f, m, _ = eng.euler_stencil(mode, ud, vd, 0, x, y, eo, 0, ny)
                    r = po.trajectory_repulsion_vec(x, y, u=u, v=v, edge_order=eo, rate_or_ratio=rr)
                    mk = np.ma.getmaskarray(r)
                    assert np.array_equal(m.cpu().numpy().astype(bool), mk) and mk.any()
                    assert np.abs(f.cpu().numpy() - np.ma.getdata(r))[~mk].max() < 4e-12 * gs
    finally:
        os.environ.pop("DLB_STENCIL_BULK", None)


def test_stencil_slabs_match_full_grid(dl, golden):
    """Row-slab geometry of the C-ABI (grow0 / nloc / halo): any partition gives the bits of
    the single-slab result (this is what each rank runs in the multi-GPU path)."""
    from dynlab_b200 import sharding
    from dynlab_b200._engine import Engine

    eng = Engine.get()
    gc = golden("ftle_c1")
    fm = torch.tensor(gc["flow_map"], device=eng.device)
    x, y = gc["x"], gc["y"]
    for eo in (1, 2):
        full, xi_full = eng.ftle_stencil(fm, 0, x, y, eo, 10.0, 0, 101, xi_mode=1)
        for size in (2, 3, 8, 64):
            parts, xis = [], []
            for r in range(size):
                s = sharding.partition(101, size, r)
                if not s.rows:
                    continue
                o, xi = eng.ftle_stencil(fm[s.lo:s.hi].contiguous(), s.lo, x, y, eo, 10.0, s.row0, s.rows, xi_mode=1)
                parts.append(o)
                xis.append(xi)
            assert torch.equal(torch.cat(parts), full)
            assert torch.equal(torch.cat(xis), xi_full)


# ---------------------------------------------------------------- K3 / ridge field
@pytest.mark.parametrize("name", EUL_CASES)
def test_eulerian_against_reference(dl, golden, name):
    g = golden("eulerian")
    D = dl.diagnostics
    x, y, eo = EUL_GRIDS[name]
    u, v = g[name + "__u"], g[name + "__v"]
    gs = max(np.abs(u).max(), np.abs(v).max()) / min(np.diff(x).min(), np.diff(y).min())  # gradient scale

    def close(a, b, tol=1e-12):
        assert np.abs(a - b).max() <= tol * gs, (name, np.abs(a - b).max(), gs)

    close(np.ma.getdata(D.AttractionRate().compute(x, y, u=u, v=v, edge_order=eo)), g[name + "__att"])
    close(np.ma.getdata(D.RepulsionRate().compute(x, y, u=u, v=v, edge_order=eo)), g[name + "__rep"])
    for key, cls in (("rate", D.TrajectoryRepulsionRate), ("ratio", D.TrajectoryRepulsionRatio)):
        r = cls().compute(x, y, u=u, v=v, edge_order=eo)
        mask = g[f"{name}__{key}_mask"]
        assert np.array_equal(np.ma.getmaskarray(r), mask)
        close(np.ma.getdata(r)[~mask], g[f"{name}__{key}"][~mask], 4e-12)
    for kind in ("attracting", "repelling"):
        for forced in (True, False):
            tag = f"{name}__iles_{kind}_{'forced' if forced else 'raw'}"
            il = D.iLES()
            il.compute(x, y, u=u, v=v, kind=kind, edge_order=eo, percentile=60,
                       force_eigenvectors=forced, debug=True)
            close(np.ma.getdata(il.field), g[tag + "__field"])
            xi_ref = g[tag + "__xi"]
            xi = np.ma.getdata(il.Xi_max)
            # Eigenvectors carry LAPACK's sign convention.  Compare where they are defined:
            #  * skip degenerate S (equal eigenvalues: no unique eigenvector);
            #  * where S11 - S22 is rounding noise (|a-d| <= 1e-12 |b|) dlanv2's sign(p) — and
            #    with it the overall sign, and the |xi0| == |xi1| tie of force_eigenvectors2D —
            #    is decided by the last bit of the gradients (numpy: separate roundings; kernel:
            #    FMA), so there the comparison is up to sign.
            dudy, dudx = np.gradient(u, y, x, edge_order=eo)
            dvdy, dvdx = np.gradient(v, y, x, edge_order=eo)
            sep = np.abs(g[name + "__rep"] - g[name + "__att"]) > 1e-9 * gs
            noise_p = np.abs(dudx - dvdy) <= 1e-12 * np.abs(0.5 * (dudy + dvdx)) + 1e-300
            strict = sep & ~noise_p
            assert strict.mean() > 0.4 or name.startswith("aduffing"), (tag, strict.mean())
            if strict.any():
                assert np.abs(xi - xi_ref)[strict].max() < 1e-7, (tag, np.abs(xi - xi_ref)[strict].max())
            loose = sep & noise_p
            if loose.any():
                d_same = np.abs(xi - xi_ref)[loose].max(axis=-1)
                d_flip = np.abs(xi + xi_ref)[loose].max(axis=-1)
                assert np.minimum(d_same, d_flip).max() < 1e-7, tag


def test_eulerian_function_route_matches_velocity_route(dl):
    """f,t route (K0 on the device) == u,v route (host arrays uploaded)."""
    D, fl = dl.diagnostics, dl.flows
    x, y = np.linspace(0, 20000, 300), np.linspace(-4000, 4000, 200)
    u, v = fl.bickley_jet(0.25, np.meshgrid(x, y))
    a = D.AttractionRate().compute(x, y, u=u, v=v, edge_order=2)
    d = D.AttractionRate()
    b = d.compute(x, y, f=fl.bickley_jet, t=0.25, edge_order=2)
    gs = np.abs(u).max() / np.diff(x).min()
    assert np.abs(a - b).max() < 1e-12 * gs
    assert np.abs(d.u - u).max() < 1e-11 * np.abs(u).max()  # lazily downloaded K0 output
    # CUDA tensors in, CUDA tensor out: nothing crosses PCIe
    ud, vd = torch.tensor(u, device="cuda"), torch.tensor(v, device="cuda")
    c = D.AttractionRate().compute_device(x, y, u=ud, v=vd, edge_order=2)
    assert c.is_cuda and np.array_equal(c.cpu().numpy(), np.ma.getdata(a))


@pytest.mark.parametrize("name,x,y,t,eo,pct", [
    ("dg_lcs", np.linspace(0, 2, 20), np.linspace(0, 1, 10), (0.1, 0), 1, None),
    ("dg_lcs_p80", np.linspace(0, 2, 41), np.linspace(0, 1, 21), (10, 0), 2, 80)])
def test_lcs_fields_against_reference(dl, golden, name, x, y, t, eo, pct):
    """LCS: FTLE, Xi_max (LAPACK sign, raw and forced), ridge field and its masks."""
    g = golden("lcs")
    for forced in (False, True):
        tag = f"{name}_{'forced' if forced else 'raw'}"
        l = dl.diagnostics.LCS()
        lines = l.compute(x, y, f=dl.flows.double_gyre, t=t, edge_order=eo, percentile=pct,
                          force_eigenvectors=forced, debug=True, rtol=1e-8, atol=1e-8)
        assert isinstance(lines, list)
        assert rel_err(np.ma.getdata(l.ftle), g[tag + "__ftle"], 1e-3).max() < 1e-7
        xi, xi_ref = np.ma.getdata(l.Xi_max), g[tag + "__xi"]
        assert (np.sum(xi * xi_ref, axis=-1) > 0).all()
        assert np.abs(xi - xi_ref).max() < 1e-6
        dd_ref, m_ref = g[tag + "__dd"], g[tag + "__dd_mask"]
        m = np.ma.getmaskarray(l.directional_derivative)
        assert (m != m_ref).mean() <= 0.01  # masks may differ where concavity ~ 0 to rounding
        both = ~m & ~m_ref
        sc = np.nanmax(np.abs(dd_ref))
        assert np.abs(np.ma.getdata(l.directional_derivative) - dd_ref)[both].max() < 1e-6 * sc


def test_readme_examples_through_dynlab_alias(dl, golden):
    """The five usage examples of the reference's README (R:README.md:20-26, :31-38, :44-51,
    :64-71, :78-84), unmodified apart from the plotting lines, after
    ``dynlab_b200.install_as_dynlab()`` - i.e. what a user who switches packages runs."""
    import importlib

    from oracle import py_oracle as po

    dl.install_as_dynlab()
    FTLE = importlib.import_module("dynlab.diagnostics").FTLE
    from dynlab.diagnostics import LCS, AttractionRate, TrajectoryRepulsionRate, iLES
    from dynlab.flows import bead_on_a_rotating_hoop, bickley_jet, double_gyre

    # FTLE (README :20-26) against the reference's own output for exactly this call
    x = np.linspace(0, 2, 101)
    y = np.linspace(0, 1, 101)
    ftle = FTLE(num_threads=1).compute(x, y, double_gyre, (10, 0), edge_order=2, rtol=1e-8, atol=1e-8)
    ref = golden("ftle_c1")["field"]
    assert isinstance(ftle, np.ma.MaskedArray) and ftle.shape == (101, 101)
    assert rel_err(np.ma.getdata(ftle), ref, 1e-3).max() < 1e-6
    # LCS (README :31-38)
    x = np.linspace(0, 2, 201)
    lcs = LCS()
    attracting_lcs = lcs.compute(x, y, f=double_gyre, t=(10, 0), percentile=80)
    assert len(attracting_lcs) > 0 and all(l.ndim == 2 and l.shape[1] == 2 for l in attracting_lcs)
    assert lcs.ftle.shape == (101, 201) and not hasattr(lcs, "flow_map")
    # iLES (README :44-51)
    a = iLES().compute(x, y, f=double_gyre, t=0, kind='attracting', force_eigenvectors=True)
    r = iLES().compute(x, y, f=double_gyre, t=0, kind='repelling', force_eigenvectors=True)
    assert len(a) > 0 and len(r) > 0
    # AttractionRate on the Bickley jet (README :64-71)
    x = np.linspace(0, 20000, 101)
    y = np.linspace(-4000, 4000, 101)
    u, v = bickley_jet(0, np.meshgrid(x, y))
    attraction_rate = AttractionRate().compute(x, y, u=u, v=v, edge_order=2)
    exp = po.attraction_repulsion_rate(x, y, u=u, v=v, edge_order=2)
    sc = np.abs(np.ma.getdata(exp)).max()
    assert np.abs(np.ma.getdata(attraction_rate) - np.ma.getdata(exp)).max() < 1e-9 * sc
    assert (-attraction_rate).shape == (101, 101)  # the README plots the negated field
    # TrajectoryRepulsionRate of the bead (README :78-84; a coarser grid for the Python oracle)
    x = np.linspace(-1, 1, 61) * 3
    y = np.linspace(-1, 1, 41) * 2.5
    rhodot = TrajectoryRepulsionRate().compute(x, y, f=bead_on_a_rotating_hoop, t=0, edge_order=2)
    exp = po.trajectory_repulsion(x, y, f=po.flow("bead_on_a_rotating_hoop"), t=0, edge_order=2)
    assert np.array_equal(np.ma.getmaskarray(rhodot), np.ma.getmaskarray(exp))
    ok = ~np.ma.getmaskarray(exp)
    sc = np.abs(np.ma.getdata(exp)[ok]).max()
    assert np.abs(np.ma.getdata(rhodot) - np.ma.getdata(exp))[ok].max() < 1e-9 * sc


def test_reference_ridge_tests_through_api(dl):
    """R:tests/test_diagnostics/test_lagrangian.py:20-44 (test_lcs) and
    R:tests/test_diagnostics/test_eulerian.py:126-193 (test_iles, test_iles_with_forced_
    eigenvectors), verbatim inputs through the public API: same number of lines, same order,
    same vertices.  (test_iles' expected lines are what the reference produces with
    force_eigenvectors=False; with its default True it produces the *_forced set.)"""
    x, y = np.linspace(0, 2, 20), np.linspace(0, 1, 10)
    dg = dl.flows.double_gyre

    def same(expected, got):
        assert len(expected) == len(got), (len(expected), len(got))
        assert all(a.shape == b.shape and np.isclose(a, b).all() for a, b in zip(expected, got))

    same(rv.LCS_LINES, dl.diagnostics.LCS().compute(x, y, f=dg, t=(0.1, 0)))
    iles = dl.diagnostics.iLES
    same(rv.ILES_FORCED_ATTRACTING, iles().compute(x, y, f=dg, t=0, kind='attracting', force_eigenvectors=True))
    same(rv.ILES_FORCED_REPELLING, iles().compute(x, y, f=dg, t=0, kind='repelling', force_eigenvectors=True))
    same(rv.ILES_FORCED_ATTRACTING, iles().compute(x, y, f=dg, t=0, kind='attracting'))
    same(rv.ILES_ATTRACTING, iles().compute(x, y, f=dg, t=0, kind='attracting', force_eigenvectors=False))
    same(rv.ILES_REPELLING, iles().compute(x, y, f=dg, t=0, kind='repelling', force_eigenvectors=False))


def test_pipelined_compute_is_bit_identical(dl):
    """FTLE.compute on large single-GPU grids integrates in row chunks and overlaps the
    device-to-host copy; field, flow map and counters equal the one-shot path bit for bit."""
    x, y = np.linspace(0, 2, 2048), np.linspace(0, 1, 2051)
    a = dl.diagnostics.FTLE()
    assert a.pipeline_chunks > 1
    fa = a.compute(x, y, dl.flows.double_gyre, (1.5, 0), edge_order=2, rtol=1e-6, atol=1e-6)
    b = dl.diagnostics.FTLE()
    b.pipeline_chunks = 1
    fb = b.compute(x, y, dl.flows.double_gyre, (1.5, 0), edge_order=2, rtol=1e-6, atol=1e-6)
    assert np.array_equal(np.ma.getdata(fa), np.ma.getdata(fb))
    assert np.array_equal(a.flow_map, b.flow_map)
    assert a.stats == b.stats and a.stats["seeds"] == 2048 * 2051
    c = dl.diagnostics.FTLE().compute_device(x, y, dl.flows.double_gyre, (1.5, 0), edge_order=2, rtol=1e-6, atol=1e-6)
    assert np.array_equal(c.cpu().numpy(), np.ma.getdata(fb))


def test_time_sweep_matches_single_calls(dl):
    """BASELINE config 5 pattern at small size: the sweep driver returns, slice by slice, the
    bits FTLE().compute / LCS(debug) give for the same interval."""
    from dynlab_b200.sweep import ftle_time_sweep

    x, y = np.linspace(0, 2, 65), np.linspace(0, 1, 33)
    spans = [(k * 10 / 4 + 5.0, k * 10 / 4) for k in range(4)]
    res = ftle_time_sweep(x, y, dl.flows.double_gyre, spans, edge_order=2, rtol=1e-8, atol=1e-8)
    assert sorted(res) == [0, 1, 2, 3]
    for k, span in enumerate(spans):
        ref = dl.diagnostics.FTLE().compute(x, y, dl.flows.double_gyre, span, edge_order=2, rtol=1e-8, atol=1e-8)
        assert np.array_equal(np.ma.getdata(res[k]), np.ma.getdata(ref))
    res = ftle_time_sweep(x, y, dl.flows.double_gyre, spans[:2], edge_order=2, percentile=80, ridge=True,
                          rtol=1e-8, atol=1e-8)
    for k in (0, 1):
        l = dl.diagnostics.LCS()
        l.compute(x, y, f=dl.flows.double_gyre, t=spans[k], edge_order=2, percentile=80, debug=True,
                  rtol=1e-8, atol=1e-8)
        ftle, dd, lines = res[k]
        assert len(lines) == len(l.lcs) and all(np.array_equal(a, b) for a, b in zip(lines, l.lcs))
        assert np.array_equal(np.ma.getdata(ftle), np.ma.getdata(l.ftle))
        assert np.array_equal(np.ma.getmaskarray(dd), np.ma.getmaskarray(l.directional_derivative))
        assert np.array_equal(np.ma.getdata(dd), np.ma.getdata(l.directional_derivative))


def test_no_out_of_bounds_writes(dl):
    """compute-sanitizer is closed on this pool, so bounds are checked by hand: outputs are
    carved out of a larger buffer whose guard bands (canary values before and after) must
    survive every kernel, over odd / even / tiny shapes and both interior-kernel variants."""
    import os

    from dynlab_b200._engine import Engine
    from dynlab_b200._resolve import resolve_flow

    eng = Engine.get()
    _, fid, _, params = resolve_flow(dl.flows.double_gyre)
    guard = 4096
    canary = -7.25e300
    try:
        for bulk in ("1", "0"):
            os.environ["DLB_STENCIL_BULK"] = bulk
            for (nx, ny, eo) in ((2, 2, 1), (3, 3, 2), (127, 5, 2), (128, 66, 1), (129, 67, 2), (257, 131, 2), (1000, 3, 2)):
                x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
                big = torch.full((guard + ny * nx * 2 + guard,), canary, dtype=torch.float64, device=eng.device)
                fm = big[guard:guard + ny * nx * 2].view(ny, nx, 2)
                eng.flow_map(fid, params, x, y, 0, ny, 2.0, 0.0, 1e-6, 1e-6, out=fm)
                bigf = torch.full((guard + ny * nx + guard,), canary, dtype=torch.float64, device=eng.device)
                fo = bigf[guard:guard + ny * nx].view(ny, nx)
                eng.ftle_stencil(fm, 0, x, y, eo, 2.0, 0, ny, out=fo)
                torch.cuda.synchronize()
                for b, n in ((big, ny * nx * 2), (bigf, ny * nx)):
                    assert (b[:guard] == canary).all() and (b[guard + n:] == canary).all(), (bulk, nx, ny, eo)
                    assert not (b[guard:guard + n] == canary).any(), (bulk, nx, ny, eo)  # every element written
                # row-slab call: only the requested rows are written
                if ny >= 6:
                    bigf.fill_(canary)
                    eng.ftle_stencil(fm[1:ny - 1], 1, x, y, eo, 2.0, 2, ny - 4, out=fo[2:ny - 2])
                    torch.cuda.synchronize()
                    assert (fo[:2] == canary).all() and (fo[ny - 2:] == canary).all()
                    assert not (fo[2:ny - 2] == canary).any()
    finally:
        os.environ.pop("DLB_STENCIL_BULK", None)


def test_fused_flow_stencil_equals_two_kernel_path(dl):
    """csrc/eulerflow.cu (K0 fused into K3, x/y-separable evaluation) is bit-identical to
    dlb_flow_eval_grid + dlb_euler_stencil: every 2-D flow, all four modes, both eigenvector
    modes, uniform and non-uniform axes, both edge orders, row slabs, tiny grids; and the lazily
    evaluated ``u, v`` attributes of the public classes."""
    from dynlab_b200 import _native as nat
    from dynlab_b200._engine import Engine
    from dynlab_b200.flows import FLOW_REGISTRY
    from oracle import py_oracle as po

    eng = Engine.get()
    rng = np.random.default_rng(8)
    grids = [(np.linspace(0.1, 1.9, 97), np.linspace(0.05, 0.95, 71)),
             (np.sort(rng.uniform(0.1, 1.9, 64)), np.linspace(0.05, 0.95, 130)),
             (np.linspace(0.1, 1.9, 33), np.sort(rng.uniform(0.05, 0.95, 67))),
             (np.sort(rng.uniform(0.1, 1.9, 5)), np.sort(rng.uniform(0.05, 0.95, 3)))]
    modes = [(nat.EULER_S_MIN, nat.XI_NONE), (nat.EULER_S_MAX, nat.XI_LAPACK),
             (nat.EULER_S_MIN, nat.XI_FORCED), (nat.EULER_RATE, nat.XI_NONE),
             (nat.EULER_RATIO, nat.XI_NONE)]
    for name, (fid, dim, _) in FLOW_REGISTRY.items():
        if dim != 2:
            continue
        params = po.flow_params(name)
        scale = 1.0
        if name == "bickley_jet":
            scale = 1.0e4  # its natural length scale (Mm-sized jet)
        for gi, (x0, y0) in enumerate(grids):
            x, y = x0 * scale, (y0 - 0.5) * scale
            ydim = len(y)
            eo = 1 + gi % 2
            u, v = eng.flow_eval_grid(fid, params, 0.37, x, y)
            # grid mode (separable tile kernel where the flow allows) == per-point evaluation
            X, Y = np.meshgrid(x, y)
            pts = torch.tensor(np.stack([X.ravel(), Y.ravel()]), device=eng.device)
            up = eng.flow_eval_points(fid, 2, params, 0.37, pts)
            assert torch.equal(up[0].view(torch.int64), u.reshape(-1).view(torch.int64)), name
            assert torch.equal(up[1].view(torch.int64), v.reshape(-1).view(torch.int64)), name
            for mode, xim in modes:
                neg = mode == nat.EULER_S_MIN and xim == nat.XI_FORCED
                ref = eng.euler_stencil(mode, u, v, 0, x, y, eo, 0, ydim, neg, xim)
                got = eng.euler_stencil_flow(mode, fid, params, 0.37, x, y, eo, 0, ydim, neg, xim)
                for a, b in zip(ref, got):
                    assert (a is None) == (b is None)
                    if a is not None:  # bit for bit: flows.cuh leaves no contraction to the compiler
                        assert torch.equal(a, b) or (a.dtype == torch.float64 and torch.equal(
                            a.view(torch.int64), b.view(torch.int64))), (name, gi, mode, xim)
                if ydim >= 8:  # a slab in the middle and the two boundary slabs
                    for r0, rows in ((0, 3), (2, ydim - 5), (ydim - 2, 2)):
                        part = eng.euler_stencil_flow(mode, fid, params, 0.37, x, y, eo, r0, rows, neg, xim)
                        assert torch.equal(part[0].view(torch.int64), got[0][r0:r0 + rows].view(torch.int64))
    # public API: f, t route == u, v route; u / v attributes appear on demand
    x, y = np.linspace(0, 2, 201), np.linspace(0, 1, 101)
    d = dl.diagnostics.TrajectoryRepulsionRate()
    f1 = d.compute(x, y, f=dl.flows.double_gyre, t=0.3, edge_order=2)
    assert "_u_host" not in d.__dict__
    u, v = d.u, d.v
    assert u.shape == (101, 201) and "_u_host" in d.__dict__
    f2 = dl.diagnostics.TrajectoryRepulsionRate().compute(x, y, u=u, v=v, edge_order=2)
    assert np.array_equal(np.ma.getdata(f1), np.ma.getdata(f2))
    assert np.array_equal(np.ma.getmaskarray(f1), np.ma.getmaskarray(f2))


def test_staged_host_to_device_copy(dl, monkeypatch):
    """Engine.to_device / to_host: the multi-threaded pinned-staging paths (large arrays) move
    exactly the bytes of the plain copy - ragged last chunk, fewer chunks than workers,
    non-contiguous and integer inputs, repeated calls reusing the staging buffers."""
    from dynlab_b200._engine import Engine

    eng = Engine.get()
    monkeypatch.setattr(Engine, "H2D_CHUNK", 1 << 13)  # 1024 doubles per chunk
    rng = np.random.default_rng(2)
    for shape in ((4097,), (5, 1024), (33, 1000), (257, 513), (1, 4096 + 7)):
        a = rng.normal(size=shape)
        assert a.nbytes >= 4 * Engine.H2D_CHUNK or a.size < 4 * 1024
        d = eng.to_device(a)
        assert tuple(d.shape) == shape and np.array_equal(d.cpu().numpy(), a)
        back = eng.to_host(d)  # staged device -> pageable host
        assert back.shape == shape and np.array_equal(back, a) and back.flags.writeable
    a = rng.normal(size=(300, 400))
    assert np.array_equal(eng.to_device(a[::2, ::3]).cpu().numpy(), a[::2, ::3])
    assert np.array_equal(eng.to_device(a.T).cpu().numpy(), a.T)
    i = rng.integers(-5, 5, size=(64, 1024))
    assert np.array_equal(eng.to_device(i).cpu().numpy(), i.astype(float))
    # through the public API: a velocity field large enough for the staged path
    x, y = np.linspace(0, 2, 301), np.linspace(0, 1, 201)
    u, v = dl.flows.double_gyre(0.3, np.meshgrid(x, y))
    f_staged = dl.diagnostics.AttractionRate().compute(x, y, u=u, v=v, edge_order=2)
    monkeypatch.setattr(Engine, "H2D_CHUNK", 16 << 20)
    f_plain = dl.diagnostics.AttractionRate().compute(x, y, u=u, v=v, edge_order=2)
    assert np.array_equal(np.ma.getdata(f_staged), np.ma.getdata(f_plain))


def test_device_percentile_bit_identical_to_numpy(dl):
    """csrc/select.cu (radix select of the two neighbouring order statistics) + numpy's lerp
    reproduces ``np.percentile`` (R:_base_classes.py:266-268) bit for bit: random signed fields,
    heavy ties, +-0, infinities, tiny arrays, q at the ends, a real FTLE field; NaN -> NaN."""
    from dynlab_b200._engine import Engine

    eng = Engine.get()
    rng = np.random.default_rng(5)

    def check(a, qs):
        d = torch.tensor(a, device=eng.device)
        for q in qs:
            got, exp = eng.percentile(d, q), float(np.percentile(a, q))
            assert got == exp or (np.isnan(got) and np.isnan(exp)), (a.shape, q, got, exp)

    qs = (0, 1e-9, 0.5, 10, 33.3, 50, 60, 80, 90, 99.9, 99.999999, 100)
    check(rng.normal(size=(257, 513)), qs)
    check(rng.normal(size=1_000_003) * 10.0 ** rng.integers(-300, 300, size=1_000_003), qs)
    check(rng.integers(-3, 4, size=(100, 1000)).astype(float), qs)       # ties
    check(np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 5e-324, -5e-324]), qs)
    check(np.array([3.25]), qs)
    check(np.array([2.0, -7.5]), qs)
    check(np.abs(rng.normal(size=(2048, 4096))), (60, 80, 95))
    with pytest.raises(ValueError):
        eng.percentile(torch.zeros(4, device=eng.device, dtype=torch.float64), 101)
    a = rng.normal(size=1000)
    a[17] = np.nan
    assert np.isnan(eng.percentile(torch.tensor(a, device=eng.device), 50))
    x, y = np.linspace(0, 2, 301), np.linspace(0, 1, 151)
    f = dl.diagnostics.FTLE().compute(x, y, dl.flows.double_gyre, (0, 8), edge_order=2)
    check(np.ma.getdata(f), qs)


def test_device_contour_segments_match_host_tracer(dl, golden):
    """csrc/contour.cu (cell classification on the device) + host linking gives exactly the
    polylines of the all-host tracer: the reference's ridge fields, random masked fields with
    saddles, single masked nodes (corner triangles), closed loops, non-finite values."""
    from dynlab_b200._contour import zero_contour_lines_device
    from oracle.contour_oracle import marching_squares
    from dynlab_b200._engine import Engine

    eng = Engine.get()

    def both(x, y, z):
        zd = np.ma.getdata(z).astype(float)
        mk = np.ma.getmaskarray(z)
        host = marching_squares(x, y, z)
        dev = zero_contour_lines_device(eng, torch.tensor(zd, device=eng.device),
                                        torch.tensor(mk.astype(np.uint8), device=eng.device), x, y)
        assert len(host) == len(dev)
        for a, b in zip(host, dev):
            assert np.array_equal(a, b)
        return host

    g = golden("contour_inputs")
    for tag, exp in (("lcs", rv.LCS_LINES), ("iles_attracting", rv.ILES_FORCED_ATTRACTING),
                     ("iles_raw_repelling", rv.ILES_REPELLING)):
        lines = both(g["x"], g["y"], np.ma.MaskedArray(g[tag + "__dd"], mask=g[tag + "__mask"]))
        assert len(lines) == len(exp) and all(np.isclose(a, b).all() for a, b in zip(exp, lines))
    rng = np.random.default_rng(11)
    for (ny, nx) in ((2, 2), (3, 7), (40, 33), (129, 200), (257, 511)):
        x, y = np.sort(rng.uniform(0, 3, nx)), np.linspace(-1, 1, ny)
        z = rng.normal(size=(ny, nx))                      # rough field: many saddles, loops
        both(x, y, np.ma.MaskedArray(z, mask=rng.random((ny, nx)) < 0.15))
        X, Y = np.meshgrid(x, y)
        zs = np.sin(3 * X) * np.cos(4 * Y) + 0.1 * X       # smooth field: long lines and loops
        zs[rng.integers(ny), rng.integers(nx)] = np.nan
        assert len(both(x, y, np.ma.MaskedArray(zs, mask=(X - 1.5) ** 2 + Y ** 2 < 0.1))) >= (1 if ny > 3 else 0)
    assert zero_contour_lines_device(eng, torch.ones(5, 5, dtype=torch.float64, device=eng.device), None,
                                     np.arange(5.0), np.arange(5.0)) == []


# ---------------------------------------------------------------- errors / edge cases
def test_error_paths(dl):
    D, fl = dl.diagnostics, dl.flows
    x, y = np.linspace(0, 2, 5), np.linspace(0, 1, 4)
    with pytest.raises(ValueError):
        D.FTLE(num_threads=0)
    with pytest.raises(ValueError):
        D.FTLE().compute(x, y, fl.double_gyre, (0, 1, 2))
    with pytest.raises(ValueError):
        D.FTLE().compute(np.zeros((2, 2)), y, fl.double_gyre, (0, 1))
    with pytest.raises(ValueError):
        D.FTLE().compute(x, y, fl.double_gyre, (0, 1), edge_order=3)
    with pytest.raises(ValueError):  # numpy: axis shorter than edge_order + 1
        D.FTLE().compute([0.0, 1.0], y, fl.double_gyre, (0, 1), edge_order=2)
    with pytest.raises(NotImplementedError):  # no device twin, no CPU fallback
        D.FTLE().compute(x, y, lambda t, Y: (Y[1], -Y[0]), (0, 1))
    with pytest.raises(NotImplementedError):  # foreign integrator
        D.FTLE(integrator=lambda f, t, Y: None).compute(x, y, fl.double_gyre, (0, 1))
    with pytest.raises(TypeError):
        D.FTLE().compute(x, y, fl.double_gyre, (0, 1), hmax=0.1)
    with pytest.raises(ZeroDivisionError):  # T == 0: the reference divides by 2|T| in Python
        D.FTLE().compute(x, y, fl.double_gyre, (1.0, 1.0))
    with pytest.raises(RuntimeError):
        D.AttractionRate().compute(x, y)
    with pytest.raises(ValueError):
        D.AttractionRate().compute(x, y, u=np.zeros(5), v=np.zeros(5))
    with pytest.raises(ValueError):
        D.iLES().compute(x, y, f=fl.double_gyre, t=0)  # default kind='attacting' (reference typo)
    with pytest.raises(ValueError):
        D.TrajectoryRepulsionRate.__mro__[1]().compute(x, y, f=fl.double_gyre, t=0, rate_or_ratio="x")


def test_small_and_degenerate_grids(dl):
    """2x2 (edge_order 1), 3x3 (edge_order 2), T == 0, integer coordinates."""
    from oracle import py_oracle as po

    D, fl = dl.diagnostics, dl.flows
    for (nx, ny, eo) in ((2, 2, 1), (3, 3, 2), (2, 5, 1), (7, 3, 2)):
        x, y = np.linspace(0.1, 1.9, nx), np.linspace(0.1, 0.9, ny)
        d = D.FTLE()
        f = d.compute(x, y, fl.double_gyre, (0, 3), edge_order=eo, rtol=1e-8, atol=1e-8)
        fm = po.flow_map(x, y, po.flow("double_gyre"), (0, 3), rtol=1e-8, atol=1e-8)
        ref = po.ftle_from_flow_map(fm, x, y, (0, 3), eo)
        assert np.abs(d.flow_map - fm).max() < 1e-10
        assert np.abs(np.ma.getdata(f) - np.ma.getdata(ref)).max() < 1e-8
    # integer coordinates (np.gradient casts to float64)
    d = D.FTLE()
    f = d.compute([0, 1, 2], [0, 1], fl.double_gyre, (2, 0))
    assert d.x.dtype.kind == "i" and np.allclose(rv.FTLE_EXPECTED, f)


def test_negative_and_sign_crossing_time_spans(dl):
    """Time intervals with negative times and crossing t = 0, both directions (the step-size
    floor uses the spacing of doubles at t in the direction of integration)."""
    from oracle import py_oracle as po

    x, y = np.linspace(-1.5, 1.5, 9), np.linspace(-1.0, 1.0, 7)
    for name, kw in (("duffing_oscillator", {}), ("double_gyre", {}), ("pendulum", {})):
        for span in ((-3.0, 2.0), (2.0, -3.0), (-4.0, -1.0), (-1.0, -4.0), (0.0, -2.5)):
            d = dl.diagnostics.FTLE()
            d.compute(x, y, getattr(dl.flows, name), span, edge_order=2, rtol=1e-8, atol=1e-8)
            ref = po.flow_map(x, y, po.flow(name), span, rtol=1e-8, atol=1e-8)
            assert (np.abs(d.flow_map - ref) / np.maximum(np.abs(ref), 1.0)).max() < 1e-9, (name, span)
            nfev = sum(po.rk45_stats(po.flow(name), span, (xx, yy), rtol=1e-8, atol=1e-8)[1]
                       for yy in y for xx in x)
            assert d.stats["nfev"] == nfev, (name, span)


def test_single_trajectory_wrappers(dl):
    """utils.odeint_wrapper / rk45_wrapper keep the integrator protocol (R:utils.py:6-25)."""
    from oracle import py_oracle as po

    tr = dl.utils.rk45_wrapper(dl.flows.lorenz, (0, 1.5), (1.0, 2.0, 3.0), rtol=1e-9, atol=1e-9)
    ref = po.rk45_wrapper(po.flow("lorenz"), (0, 1.5), (1.0, 2.0, 3.0), rtol=1e-9, atol=1e-9)
    assert tr.shape == (2, 3) and np.array_equal(tr[0], [1.0, 2.0, 3.0])
    assert np.abs(tr[-1] - ref[-1]).max() < 1e-9
    tr = dl.utils.odeint_wrapper(dl.flows.abc, (2.0, 0.0), (0.3, 0.2, 0.1))
    ref = po.rk45_wrapper(po.flow("abc"), (2.0, 0.0), (0.3, 0.2, 0.1), rtol=1.49012e-8, atol=1.49012e-8)
    assert np.abs(tr[-1] - ref[-1]).max() < 1e-10


# ---------------------------------------------------------------- full BASELINE sizes
def test_ftle_c3_full_size_properties(dl):
    """BASELINE config 3 (8192x4096, T=10, tol 1e-8) at full size:
    (a) a strided sample of seeds equals the C oracle's endpoints and nfev;
    (b) FTLE rows recomputed by the oracle on a patch agree;
    (c) the double gyre's symmetry (x, y) -> (2 - x, 1 - y) holds across the whole field."""
    from oracle import c_oracle as co
    from oracle import py_oracle as po

    x, y = np.linspace(0, 2, 8192), np.linspace(0, 1, 4096)
    d = dl.diagnostics.FTLE()
    field = np.ma.getdata(d.compute(x, y, dl.flows.double_gyre, (10, 0), edge_order=2, rtol=1e-8, atol=1e-8))
    assert field.shape == (4096, 8192) and np.isfinite(field).all()
    assert d.stats["seeds"] == 8192 * 4096 and d.stats["failed"] == 0
    fm = d.flow_map
    p = po.flow_params("double_gyre")
    # (a) every 64th row / column: 128 x 64 seeds
    ref, nfev, st = co.flow_map("double_gyre", p, x[::64], y[::64], 10.0, 0.0, 1e-8, 1e-8)
    assert np.abs(fm[::64, ::64] - ref).max() < 1e-10
    # (b) a 48 x 48 patch in the interior and the four corners (edge formulas)
    for (i0, j0) in ((2000, 4000), (0, 0), (4096 - 48, 8192 - 48), (0, 8192 - 48), (4096 - 48, 0)):
        ys, xs = y[i0:i0 + 48], x[j0:j0 + 48]
        pf, _, _ = co.flow_map("double_gyre", p, xs, ys, 10.0, 0.0, 1e-8, 1e-8)
        assert np.abs(fm[i0:i0 + 48, j0:j0 + 48] - pf).max() < 1e-10
    # stencil on the GPU flow map vs numpy on the same flow map, three row bands incl. edges
    for sl in (slice(0, 64), slice(2016, 2080), slice(4032, 4096)):
        lo, hi = max(sl.start - 2, 0), min(sl.stop + 2, 4096)
        ref_rows = po.ftle_from_flow_map(fm[lo:hi], x, y[lo:hi], (10, 0), 2)
        a = field[sl]
        b = np.ma.getdata(ref_rows)[sl.start - lo: sl.start - lo + 64]
        if sl.start > 0 and sl.stop < 4096:
            assert rel_err(a, b, 1e-3).max() < 1e-8  # cancellation noise eps*|f|/dx at dx=2.4e-4
        else:  # band touches a global edge: rows next to lo/hi cut use a different stencil
            inner = slice(0, 62) if sl.start == 0 else slice(2, 64)
            assert rel_err(a[inner], b[inner], 1e-3).max() < 1e-8
    # (c) invariants of the double gyre: the walls are invariant (u = 0 on x = 0, v = 0 on
    # y = 0 exactly; to rounding on x = 2, y = 1), the domain is invariant, and the 1-D flow
    # along the wall y = 0 preserves the order of the seeds
    assert np.array_equal(fm[:, 0, 0], np.zeros(4096)) and np.array_equal(fm[0, :, 1], np.zeros(8192))
    assert np.abs(fm[:, -1, 0] - 2.0).max() < 1e-9 and np.abs(fm[-1, :, 1] - 1.0).max() < 1e-9
    assert fm[..., 0].min() >= 0 and fm[..., 0].max() <= 2 + 1e-9
    assert fm[..., 1].min() >= 0 and fm[..., 1].max() <= 1 + 1e-9
    assert (np.diff(fm[0, :, 0]) > 0).all()
    # (d) round trip: integrate a strided sample forward 0 -> 10, then back 10 -> 0
    from dynlab_b200._engine import Engine
    from dynlab_b200._resolve import resolve_flow

    eng = Engine.get()
    _, fid, _, params = resolve_flow(dl.flows.double_gyre)
    X, Y = np.meshgrid(x[::64], y[::64])
    seeds = torch.tensor(np.stack([X.ravel(), Y.ravel()], axis=1), device=eng.device)
    fwd = eng.rk45_endpoints(fid, params, seeds, 0.0, 10.0, 1e-10, 1e-10)
    back = eng.rk45_endpoints(fid, params, fwd, 10.0, 0.0, 1e-10, 1e-10)
    assert (back - seeds).abs().max().item() < 1e-5


def test_four_times_config3_no_index_overflow(dl):
    """16384 x 8192 = 134 M seeds (4 x BASELINE config 3, 2.1 GB flow map: element offsets pass
    2^28, byte offsets 2^31): K1 + K2 on the device, strided seeds and corner / interior patches
    against the C oracle, the field against numpy on the same flow map, the LCS-side kernels
    (percentile, contour) on the full field."""
    from dynlab_b200._engine import Engine
    from oracle import c_oracle as co
    from oracle import py_oracle as po

    nx, ny = 16384, 8192
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    d = dl.diagnostics.FTLE()
    f_dev = d.compute_device(x, y, dl.flows.double_gyre, (2, 0), edge_order=2, rtol=1e-8, atol=1e-8)
    eng = Engine.get()
    st = eng.read_counters()
    assert st["seeds"] == nx * ny and st["failed"] == 0
    fm_dev = d.__dict__["_flow_map_dev"]
    p = po.flow_params("double_gyre")
    ref, _, _ = co.flow_map("double_gyre", p, x[::128], y[::128], 2.0, 0.0, 1e-8, 1e-8)
    assert np.abs(fm_dev[::128, ::128].cpu().numpy() - ref).max() < 1e-11
    for (i0, j0) in ((ny - 40, nx - 40), (0, nx - 40), (ny - 40, 0), (5000, 9000)):
        fm = fm_dev[i0:i0 + 40, j0:j0 + 40].cpu().numpy()
        pf, _, _ = co.flow_map("double_gyre", p, x[j0:j0 + 40], y[i0:i0 + 40], 2.0, 0.0, 1e-8, 1e-8)
        assert np.abs(fm - pf).max() < 1e-11
    for sl in (slice(0, 32), slice(ny - 32, ny), slice(6000, 6032)):
        lo, hi = max(sl.start - 2, 0), min(sl.stop + 2, ny)
        ref_rows = np.ma.getdata(po.ftle_from_flow_map(fm_dev[lo:hi].cpu().numpy(), x, y[lo:hi], (2, 0), 2))
        a = f_dev[sl].cpu().numpy()
        b = ref_rows[sl.start - lo: sl.start - lo + 32]
        inner = slice(0, 30) if sl.start == 0 else slice(2, 32) if sl.stop == ny else slice(0, 32)
        assert rel_err(a[inner], b[inner], 1e-3).max() < 1e-7
    # order statistics and contour cells over 134 M nodes
    q = eng.percentile(f_dev, 90)
    sample = f_dev[::16, ::16].cpu().numpy()
    assert abs(q - np.percentile(sample, 90)) < 0.05 * max(abs(q), 1e-3)
    cell, ka, kb, pts = eng.contour_segments(f_dev, None, x, y, level=q)
    assert len(cell) > 1000 and cell.min() >= 0 and cell.max() < (nx - 1) * (ny - 1)
    assert pts[:, 0].min() >= 0 and pts[:, 0].max() <= 2 and pts[:, 1].min() >= 0 and pts[:, 1].max() <= 1
    assert (ka < nx * ny * nx * ny).all() and (ka >= 0).all() and (kb >= 0).all()


def test_config4_full_size_samples(dl):
    """BASELINE config 4 at full size (4096x4096): FTLE of bead_on_a_rotating_hoop t=(0,2) and
    of duffing_oscillator t=(10,0) (step counts 476..2456 per seed) against the C oracle on a
    strided sample, and TrajectoryRepulsionRate(bead) against the oracle on crops."""
    from oracle import c_oracle as co
    from oracle import py_oracle as po

    D, fl = dl.diagnostics, dl.flows
    xb, yb = np.linspace(-1, 1, 4096) * 3, np.linspace(-1, 1, 4096) * 2.5
    xd = yd = np.linspace(-2, 2, 4096)
    for name, f, x, y, t in (("bead_on_a_rotating_hoop", fl.bead_on_a_rotating_hoop, xb, yb, (0.0, 2.0)),
                             ("duffing_oscillator", fl.duffing_oscillator, xd, yd, (10.0, 0.0))):
        d = D.FTLE()
        field = np.ma.getdata(d.compute(x, y, f, t, edge_order=2, rtol=1e-8, atol=1e-8))
        assert np.isfinite(field).all() and d.stats["failed"] == 0
        ref, nfev, st = co.flow_map(name, po.flow_params(name), x[::64], y[::64], t[0], t[1], 1e-8, 1e-8)
        got = d.flow_map[::64, ::64]
        assert (np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)).max() < 1e-7, name  # 10 * rtol
        # stencil rows recomputed by numpy on the GPU flow map (interior band)
        lo, hi = 1000, 1040
        rows = np.ma.getdata(po.ftle_from_flow_map(d.flow_map[lo - 2:hi + 2], x, y[lo - 2:hi + 2], t, 2))
        assert rel_err(field[lo:hi], rows[2:-2], 1e-3).max() < 1e-8, name
    tr = D.TrajectoryRepulsionRate()
    rho = tr.compute(xb, yb, f=fl.bead_on_a_rotating_hoop, t=0, edge_order=2)
    assert rho.shape == (4096, 4096)
    u, v = tr.u, tr.v
    gs = max(np.abs(u).max(), np.abs(v).max()) / (xb[1] - xb[0])
    for (i0, j0) in ((0, 0), (2048 - 20, 2048 - 20), (4096 - 40, 4096 - 40)):
        sl = (slice(i0, i0 + 40), slice(j0, j0 + 40))
        ref = po.trajectory_repulsion_vec(xb[sl[1]], yb[sl[0]], u=u[sl], v=v[sl], edge_order=2)
        inner = (slice(1, -1), slice(1, -1)) if 0 < i0 < 4000 else (
            (slice(0, -1), slice(0, -1)) if i0 == 0 else (slice(1, None), slice(1, None)))
        a, b = rho[sl][inner], ref[inner]
        assert np.array_equal(np.ma.getmaskarray(a), np.ma.getmaskarray(b))
        m = ~np.ma.getmaskarray(b)
        assert np.abs(np.ma.getdata(a)[m] - np.ma.getdata(b)[m]).max() < 1e-10 * gs


def test_config5_time_sweep_full_grid(dl):
    """BASELINE config 5 pattern on the full 4096x2048 grid, T = 20 (3 of the 64 slices):
    strided samples of every slice against the C oracle; ridge field produced on the device."""
    from dynlab_b200.sweep import ftle_time_sweep
    from oracle import c_oracle as co
    from oracle import py_oracle as po

    x, y = np.linspace(0, 2, 4096), np.linspace(0, 1, 2048)
    spans = [(k * 10 / 64 + 20.0, k * 10 / 64) for k in (0, 21, 63)]
    res = ftle_time_sweep(x, y, dl.flows.double_gyre, spans, edge_order=2, percentile=80, ridge=True,
                          to_host=False, rtol=1e-8, atol=1e-8)
    p = po.flow_params("double_gyre")
    for k, span in enumerate(spans):
        ftle, dd, mask, lines = res[k]
        assert ftle.is_cuda and ftle.shape == (2048, 4096)
        assert len(lines) > 0 and all(l.ndim == 2 and l.shape[1] == 2 for l in lines)
        fm, _, st = co.flow_map("double_gyre", p, x[::128], y[::128], span[0], span[1], 1e-8, 1e-8)
        fo = co.ftle(fm, y[::128], x[::128], span[0], span[1], 2)  # coarse-grid FTLE: only for scale
        assert np.isfinite(fo).all() and (st == 0).all()
        f_host = ftle.cpu().numpy()
        assert np.isfinite(f_host).all() and f_host.min() >= 0.0
        frac_masked = mask.float().mean().item()
        assert 0.8 <= frac_masked < 1.0  # percentile 80 + concavity/positivity masks
        # FTLE on the full grid vs the oracle on a dense patch
        i0, j0 = 700, 1500
        pf, _, _ = co.flow_map("double_gyre", p, x[j0:j0 + 40], y[i0:i0 + 40], span[0], span[1], 1e-8, 1e-8)
        ref = co.ftle(pf, y[i0:i0 + 40], x[j0:j0 + 40], span[0], span[1], 2)
        got = f_host[i0 + 1:i0 + 39, j0 + 1:j0 + 39]
        assert rel_err(got, ref[1:-1, 1:-1], 1e-3).max() < 1e-6


def test_eulerian_c2_full_size_properties(dl):
    """BASELINE config 2 (8192x8192 bickley_jet, edge_order 2): crops vs the oracle; and
    attraction + repulsion = trace(S) = du/dx + dv/dy = 0 for this (divergence-free) jet."""
    from oracle import py_oracle as po

    D, fl = dl.diagnostics, dl.flows
    x, y = np.linspace(0, 20000, 8192), np.linspace(-4000, 4000, 8192)
    a = D.AttractionRate()
    att = np.ma.getdata(a.compute(x, y, f=fl.bickley_jet, t=0.0, edge_order=2))
    rep = np.ma.getdata(D.RepulsionRate().compute(x, y, f=fl.bickley_jet, t=0.0, edge_order=2))
    assert att.shape == (8192, 8192) and (att <= rep).all()
    u, v = a.u, a.v
    gs = np.abs(u).max() / (x[1] - x[0])
    for (i0, j0) in ((0, 0), (4000, 3000), (8192 - 40, 8192 - 40)):
        sl = (slice(i0, i0 + 40), slice(j0, j0 + 40))
        uo, vo = po.flow("bickley_jet")(0.0, np.meshgrid(x[sl[1]], y[sl[0]]))
        assert np.abs(u[sl] - uo).max() < 1e-11 * np.abs(u).max()
        ref = po.attraction_repulsion_rate(x[sl[1]], y[sl[0]], u=u[sl], v=v[sl], edge_order=2)
        inner = (slice(1, -1), slice(1, -1)) if 0 < i0 < 8000 else (
            (slice(0, -1), slice(0, -1)) if i0 == 0 else (slice(1, None), slice(1, None)))
        assert np.abs(att[sl][inner] - np.ma.getdata(ref)[inner]).max() < 1e-11 * gs
    # trace identity (second-order stencil of a smooth divergence-free field: small, not zero)
    tr = att + rep
    assert np.abs(tr).max() < 1e-3 * np.abs(rep).max()


def _torchrun_check(n, extra, port):
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(root, "tools", "check_multigpu.py")] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert "MULTIGPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    return out.stdout


def test_multi_gpu_sharding_matches_single(dl):
    """N >= 2 GPUs, one process per GPU: row-sharded results — peer-memory transport
    (dlb_gather_rows / dlb_halo_exchange over CUDA IPC) and NCCL transport, both halo modes,
    shared / private host result, device result, work- and row-cut slabs — are bit-identical to
    the single-GPU ones.  Skipped on a 1-GPU box: ranks that wait for each other's flags must
    not share a GPU (nothing guarantees that their kernels run at the same time; see the
    profiling guide's Xid 109 note) - the next test covers the data path on one GPU inside ONE
    process, where every cross-member dependency is also a stream dependency and no kernel
    ever spins."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    _torchrun_check(min(n, 4), [], 29617)


def test_peer_abi_on_one_device(dl):
    """The multi-GPU entry points of the C-ABI called directly (ctypes, raw pointers), with two
    "members" on cuda:0: symmetric blocks, IPC handle export, row gather push + flag release,
    flag wait (already satisfied: no spinning), halo exchange, one-sided peer copy, and K2 fused
    with the gather (peer stores from the epilogue) against the plain K2."""
    import ctypes as C

    from dynlab_b200 import _native as nat
    from dynlab_b200._engine import Engine
    from dynlab_b200.peer import _Raw

    eng = Engine.get()
    lib = nat.lib
    n = C.c_int(0)
    nat.check(lib.dlb_device_count(C.byref(n)))
    assert n.value == torch.cuda.device_count()
    nat.check(lib.dlb_peer_enable(0, 0))
    stream = torch.cuda.current_stream().cuda_stream
    rows, xdim = 6, 40
    nbytes = rows * xdim * 8
    blocks = []
    for _ in range(2):
        p = C.c_void_p()
        nat.check(lib.dlb_peer_alloc(nbytes + 256, C.byref(p)))
        blocks.append(p.value)
    t = [torch.as_tensor(_Raw(b, nbytes + 256), device=eng.device) for b in blocks]
    assert all(int(x.sum()) == 0 for x in t)  # zero-filled
    handle = C.create_string_buffer(nat.IPC_HANDLE_BYTES)
    nat.check(lib.dlb_ipc_export(blocks[0], handle))
    assert any(handle.raw)
    a = t[0][:nbytes].view(torch.float64).view(rows, xdim)
    b = t[1][:nbytes].view(torch.float64).view(rows, xdim)
    a.copy_(torch.arange(rows * xdim, dtype=torch.float64, device=eng.device).view(rows, xdim))
    flags = [blk + nbytes for blk in blocks]  # uint64 words behind the rows
    # rows 1..3 of block 0 -> same place of block 1, then flag 7 in both
    off = 1 * xdim * 8
    dst = (C.c_void_p * 2)(blocks[0] + off, blocks[1] + off)
    flg = (C.c_void_p * 2)(flags[0], flags[1])
    nat.check(lib.dlb_gather_rows(blocks[0] + off, 3 * xdim * 8, dst, 2, flg, 7, stream))
    nat.check(lib.dlb_wait_flags(flags[1], 1, 7, 5.0, None, stream))
    timed_out = torch.zeros(1, dtype=torch.int32, pin_memory=True)
    nat.check(lib.dlb_wait_flags(flags[0], 1, 7, 5.0, timed_out.data_ptr(), stream))
    torch.cuda.synchronize()
    assert int(timed_out[0]) == 0
    assert torch.equal(b[1:4], a[1:4]) and float(b[0].abs().sum()) == 0 and float(b[4:].abs().sum()) == 0
    fl = t[1][nbytes:nbytes + 8].view(torch.int64)
    assert int(fl[0]) == 7
    # a wait for a value nobody publishes gives up and reports (bounded, no hang)
    nat.check(lib.dlb_wait_flags(flags[1], 1, 99, 0.05, timed_out.data_ptr(), stream))
    torch.cuda.synchronize()
    assert int(timed_out[0]) == 1
    # halo exchange: first own row (1) -> "up" slot, last own row (4) -> "down" slot, flags 3
    up = torch.zeros(xdim, dtype=torch.float64, device=eng.device)
    down = torch.zeros(xdim, dtype=torch.float64, device=eng.device)
    nat.check(lib.dlb_halo_exchange(blocks[0], xdim * 8, 1, 4, up.data_ptr(), down.data_ptr(),
                                    flags[0] + 8, flags[1] + 8, 3, stream))
    torch.cuda.synchronize()
    assert torch.equal(up, a[1]) and torch.equal(down, a[4])
    assert int(t[0][nbytes + 8:nbytes + 16].view(torch.int64)[0]) == 3
    # one-sided copy
    c = torch.zeros(2, xdim, dtype=torch.float64, device=eng.device)
    nat.check(lib.dlb_peer_copy(c.data_ptr(), blocks[0] + 2 * xdim * 8, 2 * xdim * 8, stream))
    torch.cuda.synchronize()
    assert torch.equal(c, a[2:4])
    # argument errors
    assert lib.dlb_gather_rows(blocks[0], 8, dst, 99, flg, 1, stream) == nat.DLB_ERR_ARG
    assert lib.dlb_wait_flags(flags[0], 1, 1, -1.0, None, stream) == nat.DLB_ERR_ARG
    for blk in blocks:
        nat.check(lib.dlb_peer_free(blk))
    # K2 fused with the gather == plain K2, in the local field and in two "peer" fields
    rng = np.random.default_rng(5)
    for (ny, nx, eo) in ((40, 300, 2), (7, 131, 1)):
        x, y = np.sort(rng.uniform(0, 2, nx)), np.linspace(0, 1, ny)
        fm = torch.tensor(rng.normal(size=(ny, nx, 2)), device=eng.device)
        ref, _ = eng.ftle_stencil(fm, 0, x, y, eo, 3.0, 0, ny)
        loc = eng.empty(ny, nx)
        peers = [torch.full((ny, nx), -1.0, dtype=torch.float64, device=eng.device) for _ in range(2)]
        fl = torch.zeros(4, dtype=torch.int64, device=eng.device)
        r0, r1 = 2, ny - 1  # a slab in the middle: only these rows are produced and pushed
        eng.ftle_stencil_gather(fm, 0, x, y, eo, 3.0, r0, r1 - r0, loc[r0:r1],
                                [p[r0:r1].data_ptr() for p in peers],
                                [fl[k:].data_ptr() for k in range(3)], 41)
        torch.cuda.synchronize()
        assert torch.equal(loc[r0:r1], ref[r0:r1])
        for p_ in peers:
            assert torch.equal(p_[r0:r1], ref[r0:r1])
            assert float((p_[:r0] + 1).abs().sum()) == 0 and float((p_[r1:] + 1).abs().sum()) == 0
        assert fl.tolist() == [41, 41, 41, 0]


def test_single_process_group_matches_single_device(dl):
    """One plain process driving a group of GPUs (the reference creates and tears
    down its parallelism inside one blocking call, R:_base_classes.py:139-148).  With one GPU
    the group is three members on cuda:0, with more it is every visible device; either way
    compute(), compute_device() and flow_map equal the single-device results bit for bit."""
    from dynlab_b200 import peer, sharding

    D, fl = dl.diagnostics, dl.flows
    n = torch.cuda.device_count()
    spec = list(range(n)) if n > 1 else [0, 0, 0]
    try:
        for (nx, ny, eo, t) in ((257, 131, 2, (6, 0)), (64, 5, 1, (6, 0)), (33, 3, 2, (6, 0)),
                                (2048, 1030, 2, (3, 0))):
            x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
            kw = dict(edge_order=eo, rtol=1e-8, atol=1e-8)
            peer.set_devices(None)
            d1 = D.FTLE()
            f1 = np.ma.getdata(d1.compute(x, y, fl.double_gyre, t, **kw)).copy()
            fm1 = d1.flow_map.copy()
            peer.set_devices(spec)
            for halo in ("exchange", "redundant"):
                sharding.set_halo_mode(halo)
                for weighted, overlap, chunks in ((True, False, 4), (False, True, 4),
                                                  (True, True, (0.45, 0.45, 0.07, 0.03))):
                    d = D.FTLE()
                    d.weighted_slabs, d.overlap_chunks, d.pipeline_chunks = weighted, overlap, chunks
                    d.fused_gather_chunks = 2 if overlap else (0 if weighted else 1)
                    f = np.ma.getdata(d.compute(x, y, fl.double_gyre, t, **kw))
                    assert np.array_equal(f, f1), (nx, ny, halo, weighted)
                    # counters of the sharded run = the single-device ones in BOTH halo modes
                    # (redundantly integrated halo rows are not counted)
                    assert d.stats == d1.stats, (nx, ny, halo, d.stats, d1.stats)
                    assert np.array_equal(d.flow_map, fm1), (nx, ny, halo, weighted)
                    for rep in range(4):  # rotating result buffers
                        fd = d.compute_device(x, y, fl.double_gyre, t, **kw)
                        assert np.array_equal(fd.cpu().numpy(), f1), (nx, ny, halo, weighted, rep)
            peer.current_group(nx * ny).check_timeouts()
        g = peer.current_group(1 << 30)
        assert len(g._sym) > 1
        peer.release_all()  # symmetric blocks are freed; the flags block stays for the next call
        assert list(g._sym) == [("flags", peer.FLAG_WORDS * 8)]
        d = D.FTLE()
        f = np.ma.getdata(d.compute(x, y, fl.double_gyre, t, **kw))
        assert np.array_equal(f, f1)
    finally:
        peer.set_devices(None)
        sharding.set_halo_mode("redundant")


def test_auto_multi_device_default(dl):
    """Library default in a plain process on a multi-GPU box: a grid of >= 4 M nodes is sharded
    over every visible device without the caller doing anything, with the single-device field,
    flow map and integrator counters (redundant halo rows are not counted)."""
    from dynlab_b200 import peer

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    D, fl = dl.diagnostics, dl.flows
    x, y = np.linspace(0, 2, 2048), np.linspace(0, 1, 2051)
    kw = dict(edge_order=2, rtol=1e-6, atol=1e-6)
    try:
        peer.set_devices(None)
        a = D.FTLE()
        fa = np.ma.getdata(a.compute(x, y, fl.double_gyre, (1.5, 0), **kw)).copy()
        fma = a.flow_map.copy()
        peer.set_devices("auto")
        assert peer.current_group(len(x) * len(y)).size == torch.cuda.device_count()
        assert peer.current_group(1000) is None  # small grids stay on one device
        b = D.FTLE()
        fb = np.ma.getdata(b.compute(x, y, fl.double_gyre, (1.5, 0), **kw))
        assert np.array_equal(fa, fb) and np.array_equal(b.flow_map, fma)
        assert b.stats == a.stats and b.stats["seeds"] == len(x) * len(y)
        c = D.FTLE().compute_device(x, y, fl.double_gyre, (1.5, 0), **kw)
        assert np.array_equal(c.cpu().numpy(), fa)
    finally:
        peer.set_devices(None)


