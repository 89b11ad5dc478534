"""Resolve a user-supplied flow callable to ``(flow_id, dim, params)`` for the device twins.

Accepted (north_star: "resolved from the passed callable"):
  * the functions of ``dynlab_b200.flows``;
  * the reference's own ``dynlab.flows.<name>`` functions (matched by module + name and a
    signature check, so a user can keep ``from dynlab.flows import double_gyre``);
  * ``functools.partial`` of either, with keyword overrides only;
  * ``dynlab_b200.flows.FlowSpec``.
Anything else — lambdas, closures, user-written fields — raises ``UnsupportedFlowError``:
there is no CPU fallback.
"""
from __future__ import annotations

import functools
import inspect

from . import flows as _flows


class UnsupportedFlowError(NotImplementedError):
    """The callable has no ``__device__`` twin (csrc/flows.cuh)."""


def _signature_defaults(fn, names):
    sig = inspect.signature(fn)
    params = list(sig.parameters.values())
    got = tuple(p.name for p in params[2:])
    if got != tuple(names):
        raise UnsupportedFlowError(
            f"{fn.__module__}.{fn.__name__} has parameters {got}, expected {tuple(names)}")
    return {p.name: p.default for p in params[2:]}


def resolve_flow(f):
    """Return ``(name, flow_id, dim, [param values in C-ABI order])``."""
    overrides = {}
    while isinstance(f, functools.partial):
        if f.args:
            raise UnsupportedFlowError(
                "functools.partial flows may only bind keyword parameters (t and Y stay free)")
        for k, v in (f.keywords or {}).items():
            overrides.setdefault(k, v)  # outermost partial wins, like a real call
        f = f.func
    if isinstance(f, _flows.FlowSpec):
        name = f.name
        base = _flows.__dict__[name]
        spec_over = dict(f.params)
        spec_over.update(overrides)
        overrides = spec_over
        f = base
    name = _flows._BY_FUNCTION.get(f)
    if name is None:
        mod = getattr(f, "__module__", None)
        nm = getattr(f, "__name__", None)
        if mod == "dynlab.flows" and nm in _flows.FLOW_REGISTRY and inspect.isfunction(f):
            name = nm
        else:
            raise UnsupportedFlowError(
                f"flow callable {f!r} has no device twin; pass one of dynlab_b200.flows.* "
                f"(optionally through functools.partial for parameters) — there is no CPU "
                f"fallback for arbitrary Python callables")
    flow_id, dim, pnames = _flows.FLOW_REGISTRY[name]
    defaults = _signature_defaults(f, pnames)
    unknown = set(overrides) - set(pnames)
    if unknown:
        raise TypeError(f"{name}() got an unexpected keyword argument {sorted(unknown)[0]!r}")
    values = [float(overrides.get(k, defaults[k])) for k in pnames]
    return name, flow_id, dim, values
