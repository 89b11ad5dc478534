"""TEST INFRASTRUCTURE — generate ``tests/golden/*.npz`` from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py            # all fixtures
    python oracle/make_golden.py flows ...  # selected groups

Every array stored here is an output of hokiepete/dynlab's own code
(``dynlab.diagnostics.*`` / ``dynlab.flows.*``) imported through
``oracle/ref_loader.py``; scipy's ``solve_ivp`` is attached through the
reference's pluggable ``integrator=`` argument (R:_base_classes.py:118-137).
The GPU parity tests compare against these files, so the GPU box never needs
the reference tree.
"""
from __future__ import annotations

import os
import sys
import warnings
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_loader import load_reference, rk45_wrapper  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
NT = len(os.sched_getaffinity(0))


def save(name, **arrays):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path} ({os.path.getsize(path)/1024:.1f} KiB)")


def gen_flows(dl):
    """All 15 flows on fixed pseudo-random points (R:src/dynlab/flows.py)."""
    rng = np.random.default_rng(20261019)
    out = {}
    pts2 = rng.uniform(-2.5, 2.5, size=(2, 257))
    pts3 = rng.uniform(-2.5, 2.5, size=(3, 257))
    ts = np.array([0.0, 0.37, -1.25, 7.5])
    out["pts2"], out["pts3"], out["ts"] = pts2, pts3, ts
    names2 = ["double_gyre", "autonomous_double_gyre", "bickley_jet", "glider", "hills_vortex",
              "sigmoidal", "autonomous_duffing_oscillator", "duffing_oscillator", "pendulum",
              "hurricane", "van_der_pol_oscillator", "bead_on_a_rotating_hoop", "lotka_volterra"]
    for n in names2:
        f = getattr(dl.flows, n)
        P = pts2 * (2000.0 if n == "bickley_jet" else 1.0)
        out[n] = np.stack([np.stack(f(t, (P[0], P[1]))) for t in ts])
    for n in ["abc", "lorenz"]:
        f = getattr(dl.flows, n)
        out[n] = np.stack([np.stack(f(t, (pts3[0], pts3[1], pts3[2]))) for t in ts])
    # parameter overrides
    out["double_gyre_params"] = np.stack(
        dl.flows.double_gyre(0.7, (pts2[0], pts2[1]), A=0.25, w=1.3, e=0.1))
    out["bickley_scaled"] = np.stack(
        dl.flows.bickley_jet(0.01, (pts2[0], pts2[1]), scale=4000.0))
    out["abc_forced"] = np.stack(
        dl.flows.abc(0.3, (pts3[0], pts3[1], pts3[2]), ABC_Amplitude=0.5))
    save("flows", **out)


def gen_ftle_c1(dl):
    """BASELINE config 1 (R:README.md:24-26) through the reference with the
    RK45 integrator — the K1/K2 parity pin — and with its default LSODA."""
    x = np.linspace(0, 2, 101)
    y = np.linspace(0, 1, 101)
    d = dl.diagnostics.FTLE(integrator=rk45_wrapper, num_threads=NT)
    field = d.compute(x, y, dl.flows.double_gyre, (10, 0), edge_order=2, rtol=1e-8, atol=1e-8)
    d2 = dl.diagnostics.FTLE(num_threads=NT)
    field_lsoda = d2.compute(x, y, dl.flows.double_gyre, (10, 0), edge_order=2,
                             rtol=1e-8, atol=1e-8)
    # per-seed scipy work counters for the same config (solver statistics)
    from scipy.integrate import solve_ivp
    import multiprocess

    def stats(yx):
        s = solve_ivp(dl.flows.double_gyre, (10, 0), yx[::-1], method="RK45",
                      rtol=1e-8, atol=1e-8)
        return (s.nfev, len(s.t) - 1)

    from itertools import product
    with multiprocess.Pool(NT) as p:
        st = np.array(p.map(stats, list(product(y, x)))).reshape(101, 101, 2)
    save("ftle_c1", x=x, y=y, t=np.array([10.0, 0.0]), flow_map=d.flow_map,
         field=np.ma.getdata(field), flow_map_lsoda=d2.flow_map,
         field_lsoda=np.ma.getdata(field_lsoda), nfev=st[..., 0], naccept=st[..., 1])


FTLE_CASES = {
    # name: (flow, partial-kwargs, x, y, t, edge_order, tol)
    "dg_fwd_eo1": ("double_gyre", {}, np.linspace(0, 2, 23), np.linspace(0, 1, 12), (0, 10), 1, 1e-8),
    "dg_uniform_eo2": ("double_gyre", {}, np.arange(17) * 0.125, np.arange(9) * 0.125, (5, 0), 2, 1e-8),
    "dg_uniform_eo1": ("double_gyre", {}, np.arange(17) * 0.125, np.arange(9) * 0.125, (5, 0), 1, 1e-6),
    "dg_mixed_eo2": ("double_gyre", {}, np.arange(17) * 0.125, np.linspace(0, 1, 10), (0, 4), 2, 1e-8),
    "dg_params": ("double_gyre", dict(A=0.2, w=1.0, e=0.1), np.linspace(0, 2, 15), np.linspace(0, 1, 8), (0, 6), 2, 1e-8),
    "dg_stretched": ("double_gyre", {}, np.linspace(0, 1, 19) ** 1.5 * 2, np.linspace(0, 1, 11) ** 2, (3, 0), 2, 1e-8),
    "adg": ("autonomous_double_gyre", {}, np.linspace(0, 2, 15), np.linspace(0, 1, 9), (0, 1.5), 2, 1e-8),
    "bickley": ("bickley_jet", {}, np.linspace(0, 20000, 17), np.linspace(-4000, 4000, 11), (0, 1.0), 2, 1e-8),
    "glider": ("glider", {}, np.linspace(-1, 1, 11), np.linspace(0.5, 2, 9), (0.15, 0), 1, 1e-8),
    # finite-time blow-up: scipy ends 8 seeds with status -1 (step size too small) and the
    # reference silently keeps the last accepted state (R:_base_classes.py:180-182)
    "glider_blowup": ("glider", {}, np.linspace(-1, 1, 11), np.linspace(0.5, 2, 9), (0.2, 0), 1, 1e-8),
    "hills": ("hills_vortex", {}, np.linspace(-1.5, 1.5, 13), np.linspace(0, 1.5, 9), (0, 0.4), 2, 1e-8),
    "sigmoidal": ("sigmoidal", {}, np.linspace(-1, 1, 13), np.linspace(-1, 1, 7), (0, 2.0), 2, 1e-8),
    "aduffing": ("autonomous_duffing_oscillator", {}, np.linspace(-2, 2, 13), np.linspace(-2, 2, 11), (0, 3.0), 2, 1e-8),
    "duffing_bwd": ("duffing_oscillator", {}, np.linspace(-2, 2, 13), np.linspace(-2, 2, 11), (10, 0), 2, 1e-8),
    "duffing_fwd": ("duffing_oscillator", {}, np.linspace(-2, 2, 13), np.linspace(-2, 2, 11), (0, 10), 2, 1e-8),
    "pendulum": ("pendulum", dict(gamma=0.1), np.linspace(-3, 3, 13), np.linspace(-2, 2, 9), (0, 6.0), 2, 1e-8),
    "hurricane": ("hurricane", {}, np.linspace(-2, 2, 13), np.linspace(-1, 3, 11), (0, 4.0), 2, 1e-8),
    "vdp": ("van_der_pol_oscillator", {}, np.linspace(-3, 3, 13), np.linspace(-3, 3, 11), (0, 5.0), 2, 1e-8),
    "bead": ("bead_on_a_rotating_hoop", {}, np.linspace(-1, 1, 15) * 3, np.linspace(-1, 1, 11) * 2.5, (0, 2.0), 2, 1e-8),
    "lotka": ("lotka_volterra", {}, np.linspace(0.2, 3, 13), np.linspace(0.2, 3, 11), (0, 5.0), 2, 1e-8),
    "dg_loose": ("double_gyre", {}, np.linspace(0, 2, 21), np.linspace(0, 1, 11), (0, 15), 2, 1e-3),
}


def gen_ftle_cases(dl):
    out = {}
    for name, (fl, kw, x, y, t, eo, tol) in FTLE_CASES.items():
        f = getattr(dl.flows, fl)
        if kw:
            f = partial(f, **kw)
        d = dl.diagnostics.FTLE(integrator=rk45_wrapper, num_threads=1)
        atol = tol if name != "dg_loose" else 1e-6
        field = d.compute(x, y, f, t, edge_order=eo, rtol=tol, atol=atol)
        out[name + "__flow_map"] = d.flow_map
        out[name + "__field"] = np.ma.getdata(field)
        # scipy's own per-seed counters for the same solves (solver statistics, not dynlab output)
        from itertools import product
        from scipy.integrate import solve_ivp
        st = [solve_ivp(f, t, yx[::-1], method="RK45", rtol=tol, atol=atol) for yx in product(y, x)]
        out[name + "__nfev"] = np.array([s.nfev for s in st]).reshape(len(y), len(x))
        out[name + "__status"] = np.array([s.status for s in st]).reshape(len(y), len(x))
        print(name, "ok", float(field.max()))
    save("ftle_cases", **out)


def gen_ftle_default(dl):
    """R:tests/test_diagnostics/test_lagrangian.py:6-17 inputs through the
    reference default (LSODA) and through RK45 at LSODA's default tolerances."""
    x, y = [0, 1, 2], [0, 1]
    a = dl.diagnostics.FTLE(num_threads=1)
    fa = a.compute(x, y, dl.flows.double_gyre, (2, 0))
    b = dl.diagnostics.FTLE(integrator=rk45_wrapper, num_threads=1)
    fb = b.compute(x, y, dl.flows.double_gyre, (2, 0), rtol=1.49012e-8, atol=1.49012e-8)
    save("ftle_default", field_lsoda=np.ma.getdata(fa), flow_map_lsoda=a.flow_map,
         field_rk45=np.ma.getdata(fb), flow_map_rk45=b.flow_map)


def gen_eulerian(dl):
    D = dl.diagnostics
    out = {}
    cases = {
        "bickley_eo2": ("bickley_jet", np.linspace(0, 20000, 48), np.linspace(-4000, 4000, 33), 0.0, 2),
        "bickley_eo1": ("bickley_jet", np.linspace(0, 20000, 48), np.linspace(-4000, 4000, 33), 0.5, 1),
        "dg_eo2": ("double_gyre", np.linspace(0, 2, 41), np.linspace(0, 1, 21), 0.3, 2),
        "dg_uniform": ("double_gyre", np.arange(33) / 16.0, np.arange(17) / 16.0, 0.0, 2),
        "bead_eo2": ("bead_on_a_rotating_hoop", np.linspace(-1, 1, 31) * 3, np.linspace(-1, 1, 21) * 2.5, 0.0, 2),
        "hills_eo1": ("hills_vortex", np.linspace(-1.5, 1.5, 25), np.linspace(-1.5, 1.5, 17), 0.0, 1),
        "aduffing_eo2": ("autonomous_duffing_oscillator", np.linspace(-2, 2, 21), np.linspace(-2, 2, 21), 0.0, 2),
    }
    for name, (fl, x, y, t, eo) in cases.items():
        f = getattr(dl.flows, fl)
        u, v = f(t, np.meshgrid(x, y))
        out[name + "__u"], out[name + "__v"] = np.asarray(u, float), np.asarray(v, float)
        out[name + "__att"] = np.ma.getdata(D.AttractionRate().compute(x, y, u=u, v=v, edge_order=eo))
        out[name + "__rep"] = np.ma.getdata(D.RepulsionRate().compute(x, y, f=f, t=t, edge_order=eo))
        r = D.TrajectoryRepulsionRate().compute(x, y, u=u, v=v, edge_order=eo)
        out[name + "__rate"] = np.ma.filled(r, np.nan)
        out[name + "__rate_mask"] = np.ma.getmaskarray(r)
        r = D.TrajectoryRepulsionRatio().compute(x, y, f=f, t=t, edge_order=eo)
        out[name + "__ratio"] = np.ma.filled(r, np.nan)
        out[name + "__ratio_mask"] = np.ma.getmaskarray(r)
        for kind in ("attracting", "repelling"):
            for force in (True, False):
                il = D.iLES()
                try:
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        il.compute(x, y, u=u, v=v, kind=kind, edge_order=eo, percentile=60,
                                   force_eigenvectors=force, debug=True)
                except RuntimeError as e:  # contour step unavailable (no contourpy)
                    assert "contourpy" in str(e)
                tag = f"{name}__iles_{kind}_{'forced' if force else 'raw'}"
                out[tag + "__field"] = np.ma.getdata(il.field)
                out[tag + "__xi"] = np.ma.getdata(il.Xi_max)
                dd = il.directional_derivative
                out[tag + "__dd"] = np.ma.filled(dd, np.nan)
                out[tag + "__dd_mask"] = np.ma.getmaskarray(dd)
                out[tag + "__conc"] = np.ma.getdata(il.concavity)
        print(name, "ok")
    save("eulerian", **out)


def gen_lcs(dl):
    """LCS eigen + ridge-field stage (everything before contourpy) on small
    grids, R:lagrangian.py:127-176, R:_base_classes.py:216-268."""
    out = {}
    cases = {
        "dg_lcs": (np.linspace(0, 2, 20), np.linspace(0, 1, 10), (0.1, 0), 1, None),
        "dg_lcs_p80": (np.linspace(0, 2, 41), np.linspace(0, 1, 21), (10, 0), 2, 80),
    }
    for name, (x, y, t, eo, pct) in cases.items():
        for force in (False, True):
            l = dl.diagnostics.LCS(integrator=rk45_wrapper, num_threads=NT)
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    l.compute(x, y, f=dl.flows.double_gyre, t=t, edge_order=eo, percentile=pct,
                              force_eigenvectors=force, debug=True, rtol=1e-8, atol=1e-8)
            except RuntimeError as e:
                assert "contourpy" in str(e)
            tag = f"{name}_{'forced' if force else 'raw'}"
            out[tag + "__ftle"] = np.ma.getdata(l.ftle)
            out[tag + "__xi"] = np.ma.getdata(l.Xi_max)
            out[tag + "__dd"] = np.ma.filled(l.directional_derivative, np.nan)
            out[tag + "__dd_mask"] = np.ma.getmaskarray(l.directional_derivative)
            out[tag + "__conc"] = np.ma.getdata(l.concavity)
        print(name, "ok")
    save("lcs", **out)


def gen_contour(dl):
    """Masked directional-derivative fields (the contour generator's input) of the
    reference's own ridge tests: R:tests/test_diagnostics/test_lagrangian.py:20-44 (test_lcs)
    and R:tests/test_diagnostics/test_eulerian.py:126-193 (test_iles*)."""
    out = {}
    x = np.linspace(0, 2, 20)
    y = np.linspace(0, 1, 10)

    def grab(obj, tag, **kw):
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                obj.compute(x, y, debug=True, **kw)
        except RuntimeError as e:
            assert "contourpy" in str(e)
        dd = obj.directional_derivative
        out[tag + "__dd"] = np.ma.filled(dd, np.nan)
        out[tag + "__mask"] = np.ma.getmaskarray(dd)

    grab(dl.diagnostics.LCS(), "lcs", f=dl.flows.double_gyre, t=(0.1, 0))
    for kind in ("attracting", "repelling"):
        grab(dl.diagnostics.iLES(), "iles_" + kind, f=dl.flows.double_gyre, t=0, kind=kind)
        grab(dl.diagnostics.iLES(), "iles_raw_" + kind, f=dl.flows.double_gyre, t=0, kind=kind,
             force_eigenvectors=False)
    out["x"], out["y"] = x, y
    save("contour_inputs", **out)


GROUPS = {"contour": gen_contour, "flows": gen_flows, "ftle_default": gen_ftle_default, "ftle_cases": gen_ftle_cases,
          "eulerian": gen_eulerian, "lcs": gen_lcs, "ftle_c1": gen_ftle_c1}

if __name__ == "__main__":
    dl = load_reference()
    which = sys.argv[1:] or list(GROUPS)
    for g in which:
        GROUPS[g](dl)
