"""CPU: pin the oracle (oracle/py_oracle.py, oracle/c/dynlab_oracle.c) against
(1) the reference's own known-answer vectors and (2) fixtures generated from the
unmodified reference (tests/golden/, made by oracle/make_golden.py)."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import py_oracle as po
from oracle.make_golden import FTLE_CASES
from tests import reference_vectors as rv


# ---------------- (1) the reference's own test vectors ----------------
@pytest.mark.parametrize("name", list(rv.FLOWS_2D))
def test_flow_reference_vectors_2d(name):
    x, y = np.meshgrid([0, 1, 2], [0, 1])
    u, v = po.flow(name)(0, (x, y))
    ue, ve = rv.FLOWS_2D[name]
    assert np.allclose(ue, u) and np.allclose(ve, v)
    # C oracle, scalar evaluation per point
    p = po.flow_params(name)
    for i in range(2):
        for j in range(3):
            o = co.flow_eval(name, p, 0.0, [x[i, j], y[i, j]])
            assert np.allclose([ue[i][j], ve[i][j]], o)


@pytest.mark.parametrize("name", list(rv.FLOWS_3D))
def test_flow_reference_vectors_3d(name):
    x, y, z = np.meshgrid([0, 1], [0, 1], [0, 1])
    out = po.flow(name)(0, (x, y, z))
    for e, o in zip(rv.FLOWS_3D[name], out):
        assert np.allclose(e, o)
    p = po.flow_params(name)
    for idx in np.ndindex(2, 2, 2):
        o = co.flow_eval(name, p, 0.0, [x[idx], y[idx], z[idx]])
        assert np.allclose([np.array(e)[idx] for e in rv.FLOWS_3D[name]], o)


def test_ftle_reference_vector_lsoda_and_rk45():
    f = po.flow("double_gyre")
    # reference default: LSODA at odeint's default tolerances
    field, _ = po.ftle(rv.FTLE_X, rv.FTLE_Y, f, rv.FTLE_T, integrator=po.odeint_wrapper)
    assert np.allclose(rv.FTLE_EXPECTED, field)
    # RK45 at the same default tolerances (what the CUDA path implements)
    field, _ = po.ftle(rv.FTLE_X, rv.FTLE_Y, f, rv.FTLE_T, rtol=1.49012e-8, atol=1.49012e-8)
    assert np.allclose(rv.FTLE_EXPECTED, field)
    assert field[0, 2] == 0 and field[1, 0] == 0  # lambda_max < 1 clamp, R:lagrangian.py:70-73
    # C oracle
    fm, _, st = co.flow_map("double_gyre", po.flow_params("double_gyre"), rv.FTLE_X, rv.FTLE_Y,
                            2.0, 0.0, 1.49012e-8, 1.49012e-8)
    assert (st == 0).all()
    assert np.allclose(rv.FTLE_EXPECTED, co.ftle(fm, rv.FTLE_Y, rv.FTLE_X, 2.0, 0.0, 1))


def test_eulerian_reference_vectors():
    f = po.flow("double_gyre")
    u, v = f(0, np.meshgrid(rv.EUL_X, rv.EUL_Y))
    for kw in (dict(f=f, t=0), dict(u=u, v=v)):
        assert np.allclose(rv.ATTRACTION_EXPECTED,
                           po.attraction_repulsion_rate(rv.EUL_X, rv.EUL_Y, eigenvalue_index=0, **kw))
        assert np.allclose(rv.REPULSION_EXPECTED,
                           po.attraction_repulsion_rate(rv.EUL_X, rv.EUL_Y, eigenvalue_index=-1, **kw))
    u, v = f(0, np.meshgrid(rv.TRAJ_X, rv.TRAJ_Y))
    for kw in (dict(f=f, t=0), dict(u=u, v=v)):
        for fn in (po.trajectory_repulsion, po.trajectory_repulsion_vec):
            assert np.allclose(rv.RHODOT_EXPECTED, fn(rv.TRAJ_X, rv.TRAJ_Y, rate_or_ratio="rate", **kw))
            assert np.allclose(rv.NUDOT_EXPECTED, fn(rv.TRAJ_X, rv.TRAJ_Y, rate_or_ratio="ratio", **kw))


def test_force_eigenvectors_reference_vectors():
    for a, e in rv.FORCE_EIG:
        assert (np.array([[e]]) == po.force_eigenvectors2D(np.array([[a]]))).all()


# ---------------- (2) fixtures from the unmodified reference ----------------
def test_flows_golden(golden):
    g = golden("flows")
    pts2, pts3, ts = g["pts2"], g["pts3"], g["ts"]
    for name in po.FLOW_NAMES:
        P = (pts3 if po.FLOW_DIM[name] == 3 else pts2) * (2000.0 if name == "bickley_jet" else 1.0)
        f = po.flow(name)
        p = po.flow_params(name)
        for k, t in enumerate(ts):
            out = np.stack(np.broadcast_arrays(*f(t, tuple(P))))
            assert np.array_equal(out, g[name][k]), name  # same numpy ops -> bit-identical
            for m in range(0, P.shape[1], 16):  # C oracle: libm vs numpy SIMD, few ulp
                o = co.flow_eval(name, p, float(t), P[:, m])
                np.testing.assert_allclose(o, g[name][k][:, m], rtol=1e-13, atol=1e-13 * np.abs(g[name][k]).max())
    assert np.array_equal(np.stack(po.flow("double_gyre", A=0.25, w=1.3, e=0.1)(0.7, tuple(pts2))),
                          g["double_gyre_params"])
    assert np.array_equal(np.stack(po.flow("bickley_jet", scale=4000.0)(0.01, tuple(pts2))),
                          g["bickley_scaled"])
    assert np.array_equal(np.stack(po.flow("abc", ABC_Amplitude=0.5)(0.3, tuple(pts3))), g["abc_forced"])


def test_ftle_c1_golden_c_oracle(golden):
    """BASELINE config 1 in full: C restatement of scipy RK45 vs the reference."""
    g = golden("ftle_c1")
    x, y = g["x"], g["y"]
    fm, nfev, st = co.flow_map("double_gyre", po.flow_params("double_gyre"), x, y, 10.0, 0.0, 1e-8, 1e-8)
    assert (st == 0).all()
    assert np.array_equal(nfev, g["nfev"])  # identical step sequence on all 10 201 seeds
    assert np.abs(fm - g["flow_map"]).max() < 1e-10
    f = co.ftle(fm, y, x, 10.0, 0.0, 2)
    assert np.abs(f - g["field"]).max() < 1e-10
    # stencil+eigen alone on the reference's own flow map
    f = co.ftle(g["flow_map"], y, x, 10.0, 0.0, 2)
    np.testing.assert_allclose(f, g["field"], rtol=1e-12, atol=1e-14)


def test_ftle_c1_golden_py_oracle_subgrid(golden):
    """py_oracle on a strided sub-grid of C1 (full grid takes a minute)."""
    g = golden("ftle_c1")
    x, y = g["x"][::10], g["y"][::20]
    fm = po.flow_map(x, y, po.flow("double_gyre"), (10, 0), rtol=1e-8, atol=1e-8)
    assert np.array_equal(fm, g["flow_map"][::20, ::10])
    f = po.ftle_from_flow_map(g["flow_map"], g["x"], g["y"], (10, 0), edge_order=2)
    assert np.array_equal(np.ma.getdata(f), g["field"])


@pytest.mark.parametrize("name", list(FTLE_CASES))
def test_ftle_cases_golden(golden, name):
    g = golden("ftle_cases")
    fl, kw, x, y, t, eo, tol = FTLE_CASES[name]
    atol = tol if name != "dg_loose" else 1e-6
    ref_fm, ref_f = g[name + "__flow_map"], g[name + "__field"]
    # py oracle: identical scipy/numpy calls -> bit-identical
    fm = po.flow_map(x, y, po.flow(fl, **kw), t, rtol=tol, atol=atol)
    assert np.array_equal(fm, ref_fm)
    assert np.array_equal(np.ma.getdata(po.ftle_from_flow_map(fm, x, y, t, eo)), ref_f)
    # C oracle
    cfm, nfev, st = co.flow_map(fl, po.flow_params(fl, **kw), x, y, float(t[0]), float(t[1]), tol, atol)
    assert np.array_equal(st, g[name + "__status"])  # same seeds end with "step size too small"
    ok = st == 0
    assert (~ok).any() == (name == "glider_blowup")
    assert (nfev[ok] == g[name + "__nfev"][ok]).mean() > 0.98  # same step sequences
    scale = np.maximum(np.abs(ref_fm), 1.0)
    assert (np.abs(cfm - ref_fm) / scale)[ok].max() < 10 * tol * 1e-2
    cf = co.ftle(ref_fm, y, x, float(t[0]), float(t[1]), eo)
    np.testing.assert_allclose(cf, ref_f, rtol=1e-10, atol=1e-13)


def test_ftle_default_golden(golden):
    g = golden("ftle_default")
    assert np.allclose(rv.FTLE_EXPECTED, g["field_lsoda"])
    assert np.allclose(rv.FTLE_EXPECTED, g["field_rk45"])
    f, fm = po.ftle(rv.FTLE_X, rv.FTLE_Y, po.flow("double_gyre"), rv.FTLE_T, integrator=po.odeint_wrapper)
    assert np.array_equal(fm, g["flow_map_lsoda"])


EUL_CASES = ["bickley_eo2", "bickley_eo1", "dg_eo2", "dg_uniform", "bead_eo2", "hills_eo1", "aduffing_eo2"]
EUL_GRIDS = {
    "bickley_eo2": (np.linspace(0, 20000, 48), np.linspace(-4000, 4000, 33), 2),
    "bickley_eo1": (np.linspace(0, 20000, 48), np.linspace(-4000, 4000, 33), 1),
    "dg_eo2": (np.linspace(0, 2, 41), np.linspace(0, 1, 21), 2),
    "dg_uniform": (np.arange(33) / 16.0, np.arange(17) / 16.0, 2),
    "bead_eo2": (np.linspace(-1, 1, 31) * 3, np.linspace(-1, 1, 21) * 2.5, 2),
    "hills_eo1": (np.linspace(-1.5, 1.5, 25), np.linspace(-1.5, 1.5, 17), 1),
    "aduffing_eo2": (np.linspace(-2, 2, 21), np.linspace(-2, 2, 21), 2),
}


@pytest.mark.parametrize("name", EUL_CASES)
def test_eulerian_golden(golden, name):
    g = golden("eulerian")
    x, y, eo = EUL_GRIDS[name]
    u, v = g[name + "__u"], g[name + "__v"]
    assert np.array_equal(np.ma.getdata(po.attraction_repulsion_rate(x, y, u=u, v=v, edge_order=eo)),
                          g[name + "__att"])
    assert np.array_equal(
        np.ma.getdata(po.attraction_repulsion_rate(x, y, u=u, v=v, edge_order=eo, eigenvalue_index=-1)),
        g[name + "__rep"])
    for key, rr in (("rate", "rate"), ("ratio", "ratio")):
        ref, mask = g[f"{name}__{key}"], g[f"{name}__{key}_mask"]
        r = po.trajectory_repulsion(x, y, u=u, v=v, edge_order=eo, rate_or_ratio=rr)
        assert np.array_equal(np.ma.getmaskarray(r), mask)
        assert np.array_equal(np.ma.filled(r, np.nan), ref, equal_nan=True)
        rvv = po.trajectory_repulsion_vec(x, y, u=u, v=v, edge_order=eo, rate_or_ratio=rr)
        assert np.array_equal(np.ma.getmaskarray(rvv), mask)
        sc = np.nanmax(np.abs(ref))
        np.testing.assert_allclose(np.ma.filled(rvv, 0)[~mask], ref[~mask], rtol=1e-9, atol=1e-12 * sc)
    for kind, idx in (("attracting", 0), ("repelling", -1)):
        f, xi = po.attraction_repulsion_rate(x, y, u=u, v=v, edge_order=eo, eigenvalue_index=idx, want_xi=True)
        f = np.ma.getdata(f)
        if kind == "attracting":
            f = -f
        for forced in (True, False):
            tag = f"{name}__iles_{kind}_{'forced' if forced else 'raw'}"
            X = po.force_eigenvectors2D(xi) if forced else xi
            assert np.array_equal(f, g[tag + "__field"])
            assert np.array_equal(X, g[tag + "__xi"])
            dd, conc = po.ridge_field(f, X, x, y, 60, eo)
            assert np.array_equal(np.ma.getmaskarray(dd), g[tag + "__dd_mask"])
            np.testing.assert_allclose(np.ma.filled(dd, np.nan), g[tag + "__dd"], rtol=1e-12, atol=0, equal_nan=True)
            np.testing.assert_allclose(np.ma.getdata(conc), g[tag + "__conc"], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("name,x,y,t,eo,pct", [
    ("dg_lcs", np.linspace(0, 2, 20), np.linspace(0, 1, 10), (0.1, 0), 1, None),
    ("dg_lcs_p80", np.linspace(0, 2, 41), np.linspace(0, 1, 21), (10, 0), 2, 80)])
def test_lcs_golden(golden, name, x, y, t, eo, pct):
    g = golden("lcs")
    fm = po.flow_map(x, y, po.flow("double_gyre"), t, rtol=1e-8, atol=1e-8)
    f, xi = po.ftle_from_flow_map(fm, x, y, t, eo, want_xi=True)
    for forced in (False, True):
        tag = f"{name}_{'forced' if forced else 'raw'}"
        X = po.force_eigenvectors2D(xi) if forced else xi
        assert np.array_equal(np.ma.getdata(f), g[tag + "__ftle"])
        assert np.array_equal(X, g[tag + "__xi"])
        dd, conc = po.ridge_field(f, X, x, y, pct, eo)
        assert np.array_equal(np.ma.getmaskarray(dd), g[tag + "__dd_mask"])
        np.testing.assert_allclose(np.ma.filled(dd, np.nan), g[tag + "__dd"], rtol=1e-12, atol=0, equal_nan=True)


def test_c_gradient_matches_numpy():
    rng = np.random.default_rng(5)
    for (ny, nx) in [(2, 2), (3, 3), (7, 5), (16, 33)]:
        f = rng.normal(size=(ny, nx))
        for ux in (True, False):
            for uy in (True, False):
                x = np.arange(nx) * 0.25 if ux else np.sort(rng.uniform(0, 3, nx))
                y = np.arange(ny) * 0.5 if uy else np.sort(rng.uniform(0, 3, ny))
                for eo in (1, 2):
                    if min(nx, ny) < eo + 1:
                        with pytest.raises(ValueError):
                            co.gradient(f, y, x, eo)
                        continue
                    gy, gx = np.gradient(f, y, x, edge_order=eo)
                    cy, cx = co.gradient(f, y, x, eo)
                    assert np.array_equal(gy, cy) and np.array_equal(gx, cx)


# ---------------- (3) the live reference (oracle/_ref or /root/reference), when present ----------------
_LIVE = r"""
import json, sys, numpy as np
sys.path.insert(0, sys.argv[1])
from oracle import ref_loader, py_oracle as po
ref_loader.load_reference()
from dynlab.diagnostics import FTLE, AttractionRate
from dynlab.flows import double_gyre, bickley_jet
x, y = np.linspace(0, 2, 9), np.linspace(0, 1, 6)
out = {}
for name, integ, pint in (("lsoda", None, po.odeint_wrapper), ("rk45", ref_loader.rk45_wrapper, po.rk45_wrapper)):
    d = FTLE(num_threads=2, **({} if integ is None else {"integrator": integ}))
    ref = d.compute(x, y, double_gyre, (3, 0), edge_order=2, rtol=1e-8, atol=1e-8)
    mine, fm = po.ftle(x, y, po.flow("double_gyre"), (3, 0), edge_order=2, integrator=pint, rtol=1e-8, atol=1e-8)
    out[name] = [bool(np.array_equal(np.ma.getdata(ref), np.ma.getdata(mine))), bool(np.array_equal(d.flow_map, fm))]
xe, ye = np.linspace(0, 20, 12), np.linspace(-4, 4, 7)
u, v = bickley_jet(0, np.meshgrid(xe, ye))
ref = AttractionRate().compute(xe, ye, u=u, v=v, edge_order=2)
mine = po.attraction_repulsion_rate(xe, ye, u=u, v=v, edge_order=2, eigenvalue_index=0)
out["attraction"] = [bool(np.array_equal(np.ma.getdata(ref), np.ma.getdata(mine)))]
print(json.dumps(out))
"""


def test_oracle_vs_live_reference():
    """py_oracle against the unmodified reference package itself (fresh interpreter: the GPU
    tests register dynlab_b200 under the name `dynlab` in this one)."""
    import hashlib
    import json
    import os
    import subprocess
    import sys

    from oracle import ref_loader

    root = ref_loader.reference_root()
    if root is None:
        pytest.skip("no reference package here (oracle/build_ref.sh not run)")
    sums = os.path.join(ref_loader.REF_COPY, "SHA256SUMS")
    if root == ref_loader.REF_COPY and os.path.exists(sums):
        for line in open(sums):  # the travelling copy is what build_ref.sh copied, unmodified
            digest, rel = line.split()
            assert hashlib.sha256(open(os.path.join(root, rel), "rb").read()).hexdigest() == digest
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", _LIVE, repo], capture_output=True, text=True,
                         timeout=600, cwd=repo)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out == {"lsoda": [True, True], "rk45": [True, True], "attraction": [True]}, out
