"""dynlab_b200.utils — mirrors R:src/dynlab/utils.py (odeint_wrapper, force_eigenvectors2D).

``odeint_wrapper`` keeps the reference's integrator protocol ``integrator(f, t, Y, **kw) ->
ndarray[len(t), n]`` (R:utils.py:6-25) but, as ``north_star`` requires, integrates with the
scipy-RK45-exact CUDA kernel instead of LSODA.  Its default tolerances stay those of the
reference's default integrator (scipy ``odeint``: rtol = atol = 1.49012e-8), so
R:tests/test_diagnostics/test_lagrangian.py:6-17 still holds.  The LSODA-vs-RK45 gap is a
documented semantic deviation, not a parity failure.

The diagnostics recognise these two functions as sentinels and run the whole seed grid in one
kernel launch; any other ``integrator=`` raises (no CPU fallback).
"""
from __future__ import annotations

import math

import numpy as np

ODEINT_DEFAULT_TOL = 1.49012e-8          # scipy/integrate/_odepack_py.py:148
RK45_DEFAULT_RTOL, RK45_DEFAULT_ATOL = 1e-3, 1e-6  # scipy/integrate/_ivp/rk.py:86

_ALLOWED_KW = ("rtol", "atol", "max_step", "first_step")


def parse_integrator_kwargs(kwargs: dict, default_rtol: float, default_atol: float) -> dict:
    """Validate integrator kwargs the way scipy's RK45 does (rk.py:84-104, common.py:10-60)."""
    unknown = [k for k in kwargs if k not in _ALLOWED_KW and k != "method"]
    if unknown:
        raise TypeError(f"unsupported integrator keyword(s) {unknown}; the CUDA RK45 integrator "
                        f"accepts {_ALLOWED_KW}")
    if kwargs.get("method", "RK45") != "RK45":
        raise NotImplementedError("only method='RK45' is implemented on the device")
    out = dict(rtol=float(kwargs.get("rtol", default_rtol)), atol=kwargs.get("atol", default_atol),
               max_step=float(kwargs.get("max_step", math.inf)),
               first_step=kwargs.get("first_step", None))
    if out["max_step"] <= 0:
        raise ValueError("`max_step` must be positive.")
    if out["first_step"] is not None and out["first_step"] <= 0:
        raise ValueError("`first_step` must be positive.")
    return out


def _single_trajectory(f, t, Y, default_rtol, default_atol, **kwargs):
    import torch

    from ._engine import Engine
    from ._resolve import resolve_flow

    if len(t) != 2:
        raise ValueError("t must only have 2 values, t_0 and t_final")
    name, fid, dim, params = resolve_flow(f)
    y0 = np.asarray(Y, dtype=np.float64).reshape(-1)
    if y0.size != dim:
        raise ValueError(f"{name} is a {dim}-D flow but Y has {y0.size} components")
    kw = parse_integrator_kwargs(kwargs, default_rtol, default_atol)
    t0, tf = float(t[0]), float(t[1])
    if kw["first_step"] is not None and kw["first_step"] > abs(tf - t0):
        raise ValueError("`first_step` exceeds bounds.")
    eng = Engine.get()
    seeds = torch.from_numpy(y0.reshape(1, dim)).to(eng.device)
    out = eng.rk45_endpoints(fid, params, seeds, t0, tf, kw["rtol"], kw["atol"], kw["max_step"],
                             kw["first_step"] or 0.0)
    return np.stack([y0, out.cpu().numpy()[0]])


def odeint_wrapper(f, t, Y, **kwargs):
    """Integrate one trajectory from ``t[0]`` to ``t[1]``; returns ``[[Y], [Y(t[1])]]``."""
    return _single_trajectory(f, t, Y, ODEINT_DEFAULT_TOL, ODEINT_DEFAULT_TOL, **kwargs)


def rk45_wrapper(f, t, Y, **kwargs):
    """Same kernel with ``scipy.integrate.solve_ivp``'s default tolerances (1e-3, 1e-6)."""
    return _single_trajectory(f, t, Y, RK45_DEFAULT_RTOL, RK45_DEFAULT_ATOL, **kwargs)


def force_eigenvectors2D(Xi):
    """Flip each eigenvector so that its larger-magnitude component is non-negative
    (R:utils.py:28-48; host utility for user code — the diagnostics fuse the same rule into
    the stencil kernels' eigenvector epilogue, csrc/stencil.cu: force_eigenvector)."""
    Xi = np.asarray(Xi)
    first, second = Xi[..., 0], Xi[..., 1]
    big_first = np.abs(first) >= np.abs(second)
    flip = np.where(big_first, first < 0, second < 0)
    forced = np.empty(Xi.shape)
    forced[...] = np.where(flip[..., None], -Xi, Xi)
    return forced
