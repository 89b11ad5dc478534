"""TEST INFRASTRUCTURE — numpy/scipy restatement of dynlab's hot path.

This file is the CPU oracle for the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it; the product package ``dynlab_b200`` never does.

It does not import the reference.  Each function cites the reference lines it
restates (``R:`` = /root/reference/, ``SP:`` = site-packages).  scipy/numpy are
called directly where the reference calls them (``odeint``, ``solve_ivp``,
``np.gradient``, ``np.linalg.eig``), so the third-party arithmetic is the very
same code the reference runs.

Pinning: ``tests/test_oracle_golden.py`` checks this oracle against
 * every golden vector the reference's own tests hold for the path
   (R:tests/test_flows.py, R:tests/test_diagnostics/*.py, R:tests/test_utils.py),
 * fixtures under ``tests/golden/`` produced by ``oracle/make_golden.py`` from the
   *unmodified reference* run in the build container.
Ridge *contour* geometry (contourpy) is "parity unpinned": contourpy is absent.
"""
from __future__ import annotations

import os
from itertools import product

import numpy as np

# ----------------------------------------------------------------------------
# Flow library — R:src/dynlab/flows.py:13-331.  ``FLOWS[name] = (fn, defaults)``
# fn(t, Y, *params) -> tuple.  Operation order follows the reference so that
# scalar results are bit-identical to it.
# ----------------------------------------------------------------------------
_pi = np.pi


def _double_gyre(t, Y, A, w, e):  # R:flows.py:27-33
    s = np.sin(w * t)
    a = e * s
    b = 1 - 2 * e * s
    f = a * Y[0] ** 2 + b * Y[0]
    dfdx = 2 * a * Y[0] + b
    return (-_pi * A * np.sin(_pi * f) * np.cos(Y[1] * _pi),
            _pi * A * np.cos(_pi * f) * np.sin(Y[1] * _pi) * dfdx)


def _autonomous_double_gyre(t, Y):  # R:flows.py:46-48
    return (-_pi * np.sin(_pi * Y[0]) * np.cos(Y[1] * _pi),
            _pi * np.cos(_pi * Y[0]) * np.sin(Y[1] * _pi))


def _bickley_jet(t, Y, U0, L, re, A2, A3, c2, c3, k2, k3, scale):  # R:flows.py:84-96
    U0 = U0 / scale
    L = L / scale
    re = re / scale
    c2 = c2 * U0
    c3 = c3 * U0
    k2 = k2 / re
    k3 = k3 / re
    sech2 = (1 / np.cosh(Y[1] / L)) ** 2
    th = np.tanh(Y[1] / L)
    u = (U0 * sech2 + 2 * A3 * U0 * th * sech2 * np.cos(k3 * (Y[0] - c3 * t))
         + 2 * A2 * U0 * th * sech2 * np.cos(k2 * (Y[0] - c2 * t)))
    v = (-k3 * A3 * L * U0 * sech2 * np.sin(k3 * (Y[0] - c3 * t))
         - k2 * A2 * L * U0 * sech2 * np.sin(k2 * (Y[0] - c2 * t)))
    return u, v


def _glider(t, Y):  # R:flows.py:109-113
    r = np.sqrt(Y[0] ** 2 + Y[1] ** 2)
    th2 = -2 * np.arctan2(Y[1], Y[0])
    s, c = 1.2 * np.sin(th2), 1.4 - np.cos(th2)
    return r * (-Y[1] * s - c * Y[0]), r * (Y[0] * s - c * Y[1]) - 1


def _hills_vortex(t, Y):  # R:flows.py:126-128
    return 2.0 * Y[0] * Y[1], -2.0 * (2.0 * Y[0] * Y[0] - 1.0) - 2.0 * Y[1] * Y[1]


def _sigmoidal(t, Y):  # R:flows.py:140-142
    return 0.4 * np.tanh(10.0 * (Y[0] - 0.5 * t)), Y[1] * 0


def _autonomous_duffing(t, Y):  # R:flows.py:155-157
    return Y[1], Y[0] - Y[0] ** 3


def _duffing(t, Y, A, gamma, omega):  # R:flows.py:174-176
    return Y[1], Y[0] - Y[0] ** 3 - gamma * Y[1] + A * np.sin(omega * t)


def _pendulum(t, Y, A, gamma, omega):  # R:flows.py:193-195
    return Y[1], -np.sin(Y[0]) - gamma * Y[1] + A * np.sin(omega * t)


def _hurricane(t, Y, eye_alpha, eye_omega, eye_amplitude):  # R:flows.py:216-221
    sq = np.sqrt(1.0 + 4.0 * eye_alpha)
    beta = 1.0 / sq
    yt = Y[1] - 0.5 * (1.0 + sq)
    den = Y[0] ** 2 + yt ** 2 + eye_alpha
    return -yt / den - beta, Y[0] / den + eye_amplitude * Y[1] * np.cos(eye_omega * t)


def _van_der_pol(t, Y, A, mu, gamma, omega):  # R:flows.py:239-241
    return Y[1], mu * (1 - Y[0] ** 2) * Y[1] - Y[0] - gamma * Y[1] + A * np.sin(omega * t)


def _bead(t, Y, epsilon, gamma):  # R:flows.py:260-262
    return Y[1], (1 / epsilon) * (np.sin(Y[0]) * (gamma * np.cos(Y[0]) - 1) - Y[1])


def _lotka_volterra(t, Y, alpha, beta, gamma, delta):  # R:flows.py:282-284
    return alpha * Y[0] - beta * Y[0] * Y[1], delta * Y[0] * Y[1] - gamma * Y[1]


def _abc(t, Y, amp, A, B, C):  # R:flows.py:304-307
    At = A + amp * np.sin(_pi * t)
    return (At * np.sin(Y[2]) + C * np.cos(Y[1]),
            B * np.sin(Y[0]) + At * np.cos(Y[2]),
            C * np.sin(Y[1]) + B * np.cos(Y[0]))


def _lorenz(t, Y, sigma, rho, beta):  # R:flows.py:328-331
    return sigma * (Y[1] - Y[0]), Y[0] * (rho - Y[2]) - Y[1], Y[0] * Y[1] - beta * Y[2]


# name -> (function, ordered (param, default) pairs); order = flow-id order of the C-ABI
FLOWS = {
    "double_gyre": (_double_gyre, (("A", 0.1), ("w", 0.2 * np.pi), ("e", 0.25))),
    "autonomous_double_gyre": (_autonomous_double_gyre, ()),
    "bickley_jet": (_bickley_jet, (("U0", 5413.824), ("L", 1770.0), ("re", 6371.0),
                                   ("A2", 0.1), ("A3", 0.3), ("c2", 0.205), ("c3", 0.461),
                                   ("k2", 4.0), ("k3", 6.0), ("scale", 1.0))),
    "glider": (_glider, ()),
    "hills_vortex": (_hills_vortex, ()),
    "sigmoidal": (_sigmoidal, ()),
    "autonomous_duffing_oscillator": (_autonomous_duffing, ()),
    "duffing_oscillator": (_duffing, (("A", 0.3), ("gamma", 0.2), ("omega", 1.0))),
    "pendulum": (_pendulum, (("A", 0.5), ("gamma", 0.0), ("omega", np.pi))),
    "hurricane": (_hurricane, (("eye_alpha", 0.2), ("eye_omega", 0.8), ("eye_amplitude", 0.4))),
    "van_der_pol_oscillator": (_van_der_pol, (("A", 0.2), ("mu", 2.0), ("gamma", 0.1),
                                              ("omega", np.pi))),
    "bead_on_a_rotating_hoop": (_bead, (("epsilon", 0.1), ("gamma", 2.3))),
    "lotka_volterra": (_lotka_volterra, (("alpha", 2.0 / 3.0), ("beta", 4.0 / 3.0),
                                         ("gamma", 1.0), ("delta", 1.0))),
    "abc": (_abc, (("ABC_Amplitude", 0.0), ("A", np.sqrt(3)), ("B", np.sqrt(2)), ("C", 1.0))),
    "lorenz": (_lorenz, (("sigma", 10.0), ("rho", 28.0), ("beta", 8.0 / 3.0))),
}
FLOW_NAMES = list(FLOWS)
FLOW_DIM = {n: (3 if n in ("abc", "lorenz") else 2) for n in FLOW_NAMES}


def flow(name: str, **overrides):
    """Return ``f(t, Y)`` for the named flow with parameter overrides."""
    fn, plist = FLOWS[name]
    vals = [overrides.pop(k, d) for k, d in plist]
    if overrides:
        raise TypeError(f"unknown parameters for {name}: {sorted(overrides)}")

    def f(t, Y):
        return fn(t, Y, *vals)

    f.__name__ = name
    return f


def flow_params(name: str, **overrides):
    return [float(overrides.get(k, d)) for k, d in FLOWS[name][1]]


# ----------------------------------------------------------------------------
# Integrators — R:src/dynlab/utils.py:6-25 (LSODA default) and the RK45 wrapper
# the parity contract is stated against.
# ----------------------------------------------------------------------------
def odeint_wrapper(f, t, Y, **kwargs):
    """R:utils.py:23-25 — scipy LSODA with tfirst=True."""
    from scipy.integrate import odeint

    return odeint(f, Y, t, tfirst=True, **kwargs)


def rk45_wrapper(f, t, Y, **kwargs):
    """solve_ivp(method='RK45') in the reference's integrator protocol
    (R:_base_classes.py:180-182: result[-1, :] is the endpoint)."""
    from scipy.integrate import solve_ivp

    return solve_ivp(f, t, Y, method="RK45", **kwargs).y.T


def rk45_stats(f, t, Y, **kwargs):
    """Endpoint plus (nfev, accepted steps, status) from scipy's RK45."""
    from scipy.integrate import solve_ivp

    s = solve_ivp(f, t, Y, method="RK45", **kwargs)
    return s.y[:, -1], s.nfev, len(s.t) - 1, s.status


# ----------------------------------------------------------------------------
# Diagnostic2D / LagrangianDiagnostic2D — R:_base_classes.py:26-36, 175-189
# ----------------------------------------------------------------------------
def _coords(x, y):
    """R:_base_classes.py:26-36."""
    y = (np.array(y) if type(y) is not np.ndarray else y).squeeze()
    x = (np.array(x) if type(x) is not np.ndarray else x).squeeze()
    if len(y.shape) > 1:
        raise ValueError("y must be a 1-dimensional array.")
    if len(x.shape) > 1:
        raise ValueError("x must be a 1-dimensional array.")
    return x, y


class _Seed:
    """Picklable seed -> endpoint map (the reference uses a lambda + dill)."""

    def __init__(self, integrator, f, t, kwargs):
        self.integrator, self.f, self.t, self.kwargs = integrator, f, t, kwargs

    def __call__(self, yx):
        # R:_base_classes.py:180-182 — seed is (y, x); reversed to (x, y)
        return self.integrator(self.f, self.t, yx[::-1], **self.kwargs)[-1, :]


def flow_map(x, y, f, t, integrator=rk45_wrapper, num_threads=1, **kwargs):
    """R:_base_classes.py:171-189.  y outer, x inner; result [ydim, xdim, 2].

    num_threads>1 uses a process pool over seeds like the reference's pathos
    pool (R:_base_classes.py:141-146); stdlib multiprocessing stands in for
    pathos, which is not installed.  ``f`` must then be picklable (use
    ``FlowSpec`` below)."""
    x, y = _coords(x, y)
    if len(t) != 2:
        raise ValueError("t must only have 2 values, t_0 and t_final")
    job = _Seed(integrator, f, t, kwargs)
    seeds = list(product(y, x))
    if num_threads == 1:
        out = list(map(job, seeds))
    elif num_threads > 1:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(num_threads) as p:
            out = p.map(job, seeds)
    else:
        raise ValueError("num_threads must be an integer greater than 0.")
    return np.array(out).reshape([len(y), len(x), 2])


class FlowSpec:
    """Picklable named flow (for the process pool)."""

    def __init__(self, name, **overrides):
        self.name, self.overrides = name, overrides
        self._vals = [overrides.get(k, d) for k, d in FLOWS[name][1]]

    def __call__(self, t, Y):
        return FLOWS[self.name][0](t, Y, *self._vals)


# ----------------------------------------------------------------------------
# FTLE / LCS eigen stage — R:lagrangian.py:50-78, 130-174
# ----------------------------------------------------------------------------
def _eig_sorted_2x2(M):
    """Per-point ``np.linalg.eig`` + argsort over a stack ``M[..., 2, 2]``.
    The stacked call dispatches LAPACK dgeev once per matrix, i.e. the same
    routine the reference's per-point loop reaches (R:lagrangian.py:69,148-151;
    R:eulerian.py:54-56)."""
    w, V = np.linalg.eig(M)
    w = np.real(w)
    V = np.real(V)
    idx = np.argsort(w, axis=-1)
    return w, V, idx


def ftle_from_flow_map(fm, x, y, t, edge_order=1, want_xi=False):
    """R:lagrangian.py:50-76 (and :130-167 for Xi_max)."""
    x, y = _coords(x, y)
    dfxdy, dfxdx = np.gradient(fm[:, :, 0].squeeze(), y, x, edge_order=edge_order)
    dfydy, dfydx = np.gradient(fm[:, :, 1].squeeze(), y, x, edge_order=edge_order)
    JF = np.empty(fm.shape[:2] + (2, 2))
    JF[..., 0, 0], JF[..., 0, 1] = dfxdx, dfxdy
    JF[..., 1, 0], JF[..., 1, 1] = dfydx, dfydy
    C = np.matmul(np.swapaxes(JF, -1, -2), JF)
    w, V, idx = _eig_sorted_2x2(C)
    lam = np.take_along_axis(w, idx[..., -1:], axis=-1)[..., 0]
    T = abs(t[-1] - t[0])
    with np.errstate(divide="ignore", invalid="ignore"):
        field = np.where(lam >= 1, 1.0 / (2.0 * T) * np.log(np.where(lam >= 1, lam, 1.0)), 0.0)
    field = np.ma.MaskedArray(field)
    if not want_xi:
        return field
    xi = np.take_along_axis(V, idx[..., None, -1:], axis=-1)[..., 0]
    return field, xi


def ftle(x, y, f, t, edge_order=1, integrator=rk45_wrapper, num_threads=1, **kwargs):
    fm = flow_map(x, y, f, t, integrator=integrator, num_threads=num_threads, **kwargs)
    return ftle_from_flow_map(fm, x, y, t, edge_order), fm


# ----------------------------------------------------------------------------
# Eulerian — R:_base_classes.py:91-113, R:eulerian.py:47-59,159-195,317-343
# ----------------------------------------------------------------------------
def eulerian_gradients(x, y, u=None, v=None, f=None, t=None, edge_order=1):
    x, y = _coords(x, y)
    if u is not None and v is not None:
        v = (np.array(v) if type(v) is not np.ndarray else v).squeeze()
        u = (np.array(u) if type(u) is not np.ndarray else u).squeeze()
        if len(v.shape) != 2:
            raise ValueError("v must be a 2-dimensional array.")
        if len(u.shape) != 2:
            raise ValueError("u must be a 2-dimensional array.")
    elif f is not None and t is not None:
        u, v = f(t, np.meshgrid(x, y))
    else:
        raise RuntimeError("Either a velocity field (u, v) or a function and timestep from which "
                           "to calculate a velocity field must be passed to the compute function.")
    dudy, dudx = np.gradient(u, y, x, edge_order=edge_order)
    dvdy, dvdx = np.gradient(v, y, x, edge_order=edge_order)
    return u, v, dudy, dudx, dvdy, dvdx


def _strain(dudy, dudx, dvdy, dvdx):
    G = np.empty(dudx.shape + (2, 2))
    G[..., 0, 0], G[..., 0, 1] = dudx, dudy
    G[..., 1, 0], G[..., 1, 1] = dvdx, dvdy
    return 0.5 * (G + np.swapaxes(G, -1, -2))


def attraction_repulsion_rate(x, y, u=None, v=None, f=None, t=None, edge_order=1,
                              eigenvalue_index=0, want_xi=False):
    """R:eulerian.py:41-61 (index 0 = attraction, -1 = repulsion); with
    ``want_xi`` also the eigenvector of R:eulerian.py:317-333."""
    _, _, dudy, dudx, dvdy, dvdx = eulerian_gradients(x, y, u, v, f, t, edge_order)
    S = _strain(dudy, dudx, dvdy, dvdx)
    w, V, idx = _eig_sorted_2x2(S)
    k = idx[..., eigenvalue_index]
    field = np.ma.MaskedArray(np.take_along_axis(w, k[..., None], axis=-1)[..., 0])
    if not want_xi:
        return field
    xi = np.take_along_axis(V, k[..., None, None], axis=-1)[..., 0]
    return field, xi


def trajectory_repulsion(x, y, u=None, v=None, f=None, t=None, edge_order=1,
                         rate_or_ratio="rate"):
    """R:eulerian.py:157-196, evaluated per point with the reference's own
    matrix expressions (small grids only — Python loop)."""
    u, v, dudy, dudx, dvdy, dvdx = eulerian_gradients(x, y, u, v, f, t, edge_order)
    if rate_or_ratio not in ("rate", "ratio"):
        raise ValueError('rate_or_ratio parameter must be either "rate" or "ratio".')
    J = np.array([[0, 1], [-1, 0]])
    ydim, xdim = dudx.shape
    field = np.ma.empty([ydim, xdim])
    for i, j in product(range(ydim), range(xdim)):
        G = np.array([[dudx[i, j], dudy[i, j]], [dvdx[i, j], dvdy[i, j]]])
        S = 0.5 * (G + G.T)
        V = np.array([u[i, j], v[i, j]])
        vv = np.dot(V, V)
        if vv:
            if rate_or_ratio == "rate":
                field[i, j] = np.dot(V, np.dot(np.matmul(J.T, np.matmul(S, J)), V)) / vv
            else:
                field[i, j] = np.dot(V, np.dot(np.trace(S) * np.identity(2) - 2 * S, V)) / vv
        else:
            field[i, j] = np.ma.masked
    return field


def trajectory_repulsion_vec(x, y, u=None, v=None, f=None, t=None, edge_order=1,
                             rate_or_ratio="rate"):
    """Vectorised algebraic expansion of the same quadratic forms;
    checked against ``trajectory_repulsion`` in the oracle tests.  Used for
    large-grid comparisons."""
    u, v, dudy, dudx, dvdy, dvdx = eulerian_gradients(x, y, u, v, f, t, edge_order)
    a, d, b = dudx, dvdy, 0.5 * (dudy + dvdx)
    vv = u * u + v * v
    with np.errstate(divide="ignore", invalid="ignore"):
        if rate_or_ratio == "rate":
            num = d * u * u - 2 * b * u * v + a * v * v
        else:
            num = (d - a) * u * u - 4 * b * u * v + (a - d) * v * v
        out = num / vv
    return np.ma.MaskedArray(out, mask=(vv == 0))


def force_eigenvectors2D(Xi):
    """R:utils.py:41-45, vectorised."""
    Xi = np.asarray(Xi, dtype=float)
    a0, a1 = np.abs(Xi[..., 0]), np.abs(Xi[..., 1])
    flip = ((a0 >= a1) & (Xi[..., 0] < 0)) | ((a0 < a1) & (Xi[..., 1] < 0))
    return np.where(flip[..., None], -Xi, Xi)


# ----------------------------------------------------------------------------
# Ridge field — R:_base_classes.py:216-268 (everything before the contour call)
# ----------------------------------------------------------------------------
def ridge_field(field, Xi, x, y, percentile, edge_order):
    x, y = _coords(x, y)
    field = np.ma.MaskedArray(field)
    dfdy, dfdx = np.gradient(field, y, x, edge_order=edge_order)
    dfdydy, dfdydx = np.gradient(dfdy, y, x, edge_order=edge_order)
    dfdxdy, dfdxdx = np.gradient(dfdx, y, x, edge_order=edge_order)
    dd = dfdx * Xi[..., 0] + dfdy * Xi[..., 1]
    conc = ((dfdxdx * Xi[..., 0] + dfdxdy * Xi[..., 1]) * Xi[..., 0]
            + (dfdydx * Xi[..., 0] + dfdydy * Xi[..., 1]) * Xi[..., 1])
    dd = np.ma.masked_where(field <= 0, np.ma.MaskedArray(dd))
    dd = np.ma.masked_where(conc >= 0, dd)
    if percentile:
        dd = np.ma.masked_where(field <= np.percentile(np.ma.getdata(field), percentile), dd)
    return dd, conc


def cpu_count() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1
