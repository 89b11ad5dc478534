"""dynlab_b200.flows — the flow library of dynlab (R:src/dynlab/flows.py:13-331).

Every flow keeps the reference's public signature ``f(t, Y, **params)`` and works on
scalars and on numpy arrays (e.g. ``np.meshgrid`` output), so user code that evaluates a
flow directly for plotting keeps working.  The functions here are *descriptors* as far as the
diagnostics are concerned: ``dynlab_b200.diagnostics`` never calls them on the host — it
resolves the callable to ``(flow_id, params)`` (``dynlab_b200._resolve``) and runs the
``__device__`` twin in ``csrc/flows.cuh``.  A callable that cannot be resolved raises; there is
no CPU fallback.

Ids and parameter order are the C-ABI's (``include/dynlab_b200.h``: ``enum dlb_flow``).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "double_gyre", "autonomous_double_gyre", "bickley_jet", "glider", "hills_vortex",
    "sigmoidal", "autonomous_duffing_oscillator", "duffing_oscillator", "pendulum", "hurricane",
    "van_der_pol_oscillator", "bead_on_a_rotating_hoop", "lotka_volterra", "abc", "lorenz",
    "FLOW_REGISTRY", "FlowSpec",
]

_PI = np.pi
_sin, _cos, _tanh, _cosh = np.sin, np.cos, np.tanh, np.cosh

# name -> (flow_id, dim, parameter names in C-ABI order); filled by @_flow
FLOW_REGISTRY: dict[str, tuple[int, int, tuple[str, ...]]] = {}
_BY_FUNCTION: dict[object, str] = {}


def _flow(flow_id: int, dim: int = 2):
    def deco(fn):
        code = fn.__code__
        names = code.co_varnames[2:code.co_argcount]
        FLOW_REGISTRY[fn.__name__] = (flow_id, dim, tuple(names))
        _BY_FUNCTION[fn] = fn.__name__
        fn.flow_id = flow_id
        fn.flow_dim = dim
        return fn
    return deco


@_flow(0)
def double_gyre(t, Y, A=0.1, w=0.2 * np.pi, e=0.25):
    """Time-dependent double gyre on [0, 2] x [0, 1] (R:flows.py:13-33).

    A: gyre velocity amplitude, w: forcing frequency, e: forcing strength."""
    x, y = Y[0], Y[1]
    forcing = _sin(w * t)
    a = e * forcing
    b = 1 - 2 * e * forcing
    f = a * x ** 2 + b * x
    dfdx = 2 * a * x + b
    return (-_PI * A * _sin(_PI * f) * _cos(y * _PI),
            _PI * A * _cos(_PI * f) * _sin(y * _PI) * dfdx)


@_flow(1)
def autonomous_double_gyre(_, Y):
    """Time-independent double gyre on [0, 2] x [0, 1] (R:flows.py:36-48)."""
    x, y = Y[0], Y[1]
    return -_PI * _sin(_PI * x) * _cos(y * _PI), _PI * _cos(_PI * x) * _sin(y * _PI)


@_flow(2)
def bickley_jet(t, Y, U0=5413.824, L=1770.0, re=6371.0, A2=0.1, A3=0.3, c2=0.205, c3=0.461,
                k2=4.0, k3=6.0, scale=1.0):
    """Time-dependent Bickley jet on [0, 20000] x [-4000, 4000] km (R:flows.py:51-96).

    ``scale`` shrinks the domain (scale=4000 -> [0, 5] x [-1, 1])."""
    x, y = Y[0], Y[1]
    U0, L, re = U0 / scale, L / scale, re / scale
    c2, c3 = c2 * U0, c3 * U0
    k2, k3 = k2 / re, k3 / re
    sech2 = (1 / _cosh(y / L)) ** 2
    th = _tanh(y / L)
    ph3, ph2 = k3 * (x - c3 * t), k2 * (x - c2 * t)
    u = U0 * sech2 + 2 * A3 * U0 * th * sech2 * _cos(ph3) + 2 * A2 * U0 * th * sech2 * _cos(ph2)
    v = -k3 * A3 * L * U0 * sech2 * _sin(ph3) - k2 * A2 * L * U0 * sech2 * _sin(ph2)
    return u, v


@_flow(3)
def glider(_, Y):
    """Time-independent glider flow (R:flows.py:99-113)."""
    x, y = Y[0], Y[1]
    speed = np.sqrt(x ** 2 + y ** 2)
    ang = -2 * np.arctan2(y, x)
    lift, drag = 1.2 * _sin(ang), 1.4 - _cos(ang)
    return speed * (-y * lift - drag * x), speed * (x * lift - drag * y) - 1


@_flow(4)
def hills_vortex(_, Y):
    """Time-independent Hill's vortex (R:flows.py:116-128)."""
    x, y = Y[0], Y[1]
    return 2.0 * x * y, -2.0 * (2.0 * x * x - 1.0) - 2.0 * y * y


@_flow(5)
def sigmoidal(t, Y):
    """Time-dependent sigmoidal flow (R:flows.py:131-142); v is identically zero."""
    x, y = Y[0], Y[1]
    return 0.4 * _tanh(10.0 * (x - 0.5 * t)), y * 0


@_flow(6)
def autonomous_duffing_oscillator(_, Y):
    """Time-independent Duffing oscillator (R:flows.py:145-157)."""
    x, y = Y[0], Y[1]
    return y, x - x ** 3


@_flow(7)
def duffing_oscillator(t, Y, A=0.3, gamma=0.2, omega=1.0):
    """Forced, damped Duffing oscillator (R:flows.py:160-176)."""
    x, y = Y[0], Y[1]
    return y, x - x ** 3 - gamma * y + A * _sin(omega * t)


@_flow(8)
def pendulum(t, Y, A=0.5, gamma=0.0, omega=np.pi):
    """Forced, damped pendulum (R:flows.py:179-195)."""
    x, y = Y[0], Y[1]
    return y, -_sin(x) - gamma * y + A * _sin(omega * t)


@_flow(9)
def hurricane(t, Y, eye_alpha=0.2, eye_omega=0.8, eye_amplitude=0.4):
    """Time-dependent hurricane centred on x=0, y=1 (R:flows.py:198-221)."""
    x, y = Y[0], Y[1]
    root = np.sqrt(1.0 + 4.0 * eye_alpha)
    beta = 1.0 / root
    yc = y - 0.5 * (1.0 + root)
    den = x ** 2 + yc ** 2 + eye_alpha
    return -yc / den - beta, x / den + eye_amplitude * y * _cos(eye_omega * t)


@_flow(10)
def van_der_pol_oscillator(t, Y, A=0.2, mu=2.0, gamma=0.1, omega=np.pi):
    """Forced Van der Pol oscillator (R:flows.py:224-241)."""
    x, y = Y[0], Y[1]
    return y, mu * (1 - x ** 2) * y - x - gamma * y + A * _sin(omega * t)


@_flow(11)
def bead_on_a_rotating_hoop(_, Y, epsilon=0.1, gamma=2.3):
    """Bead on a rotating hoop (R:flows.py:244-262); stiff-ish for small epsilon."""
    x, y = Y[0], Y[1]
    return y, (1 / epsilon) * (_sin(x) * (gamma * _cos(x) - 1) - y)


@_flow(12)
def lotka_volterra(_, Y, alpha=2.0 / 3.0, beta=4.0 / 3.0, gamma=1.0, delta=1.0):
    """Lotka-Volterra predator-prey flow (R:flows.py:265-284)."""
    x, y = Y[0], Y[1]
    return alpha * x - beta * x * y, delta * x * y - gamma * y


@_flow(13, dim=3)
def abc(t, Y, ABC_Amplitude=0.0, A=np.sqrt(3), B=np.sqrt(2), C=1.0):
    """Time-dependent Arnold-Beltrami-Childress flow, 3-D (R:flows.py:287-307)."""
    x, y, z = Y[0], Y[1], Y[2]
    At = A + ABC_Amplitude * _sin(_PI * t)
    return (At * _sin(z) + C * _cos(y), B * _sin(x) + At * _cos(z), C * _sin(y) + B * _cos(x))


@_flow(14, dim=3)
def lorenz(_, Y, sigma=10.0, rho=28.0, beta=8.0 / 3.0):
    """Lorenz system, 3-D (R:flows.py:310-331)."""
    x, y, z = Y[0], Y[1], Y[2]
    return sigma * (y - x), x * (rho - z) - y, x * y - beta * z


class FlowSpec:
    """Explicit flow descriptor: ``FlowSpec('double_gyre', A=0.2)``.  Callable like the
    function it names (host evaluation) and accepted by every diagnostic."""

    def __init__(self, name: str, **params):
        if name not in FLOW_REGISTRY:
            raise KeyError(f"unknown flow {name!r}; known: {sorted(FLOW_REGISTRY)}")
        unknown = set(params) - set(FLOW_REGISTRY[name][2])
        if unknown:
            raise TypeError(f"{name}() got unexpected parameters {sorted(unknown)}")
        self.name, self.params = name, dict(params)

    def __call__(self, t, Y):
        return globals()[self.name](t, Y, **self.params)

    def __repr__(self):
        return f"FlowSpec({self.name!r}, **{self.params!r})"
