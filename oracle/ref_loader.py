"""TEST / BENCH INFRASTRUCTURE — loader for the real, unmodified hokiepete/dynlab reference.

Two places can hold the reference package:

* ``oracle/_ref/dynlab`` — the byte-for-byte copy made by ``oracle/build_ref.sh`` (git-ignored,
  travels to the GPU box with the gpurun snapshot; ``SHA256SUMS`` beside it records what was
  copied).  This is what ``bench.py --impl reference`` and the ``cpu_baseline`` leg time.
* ``/root/reference/src`` — the read-only tree of the build container, used by
  ``oracle/make_golden.py`` to generate the fixtures under ``tests/golden/`` and by
  ``tests/test_oracle_golden.py::test_oracle_vs_live_reference`` (skipped when neither exists).

Nothing in ``dynlab_b200/`` imports this module; the product never runs reference code.

The reference imports ``contourpy`` and ``pathos`` at module scope
(/root/reference/src/dynlab/diagnostics/_base_classes.py:7-8); neither is in
this image.  We pre-seed ``sys.modules`` with two tiny stand-ins so that the
unmodified reference package imports:

* ``pathos.multiprocessing.ProcessingPool`` -> ``multiprocess.Pool`` (pathos's
  own backend, which *is* installed);
* ``contourpy.contour_generator`` -> a stub that raises when ``.lines`` is
  called (only the ridge-line step of LCS/iLES needs it).
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(_HERE, "_ref")           # oracle/build_ref.sh
REFERENCE_TREE = "/root/reference/src"           # build container only


def reference_root():
    """Directory to put on sys.path for ``import dynlab`` (the travelling copy first), or None."""
    for root in (REF_COPY, REFERENCE_TREE):
        if os.path.isfile(os.path.join(root, "dynlab", "diagnostics", "lagrangian.py")):
            return root
    return None


REFERENCE_SRC = reference_root() or REFERENCE_TREE


def reference_available() -> bool:
    return reference_root() is not None


def _install_shims() -> None:
    if "contourpy" not in sys.modules:
        try:
            import contourpy  # noqa: F401
        except ModuleNotFoundError:
            m = types.ModuleType("contourpy")

            class _NoContour:
                def __init__(self, *a, **k):
                    self.args = (a, k)

                def lines(self, level):
                    raise RuntimeError("contourpy is not installed in this image")

            m.contour_generator = lambda *a, **k: _NoContour(*a, **k)
            sys.modules["contourpy"] = m
    if "pathos" not in sys.modules:
        try:
            import pathos  # noqa: F401
        except ModuleNotFoundError:
            import multiprocess

            p = types.ModuleType("pathos")
            pm = types.ModuleType("pathos.multiprocessing")
            pm.ProcessingPool = multiprocess.Pool
            p.multiprocessing = pm
            sys.modules["pathos"] = p
            sys.modules["pathos.multiprocessing"] = pm


def load_reference():
    """Return the unmodified reference package ``dynlab`` (with diagnostics)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference package not found: run oracle/build_ref.sh where "
                           "/root/reference exists")
    mod = sys.modules.get("dynlab")
    if mod is not None and not (getattr(mod, "__file__", "") or "").startswith(root):
        raise RuntimeError("`dynlab` is already imported from somewhere else "
                           f"({getattr(mod, '__file__', None)}): dynlab_b200.install_as_dynlab() "
                           "and the real reference cannot share a process")
    _install_shims()
    if root not in sys.path:
        sys.path.insert(0, root)
    import dynlab  # noqa: F401
    import dynlab.diagnostics  # noqa: F401
    import dynlab.flows  # noqa: F401
    import dynlab.utils  # noqa: F401

    return sys.modules["dynlab"]


def rk45_wrapper(f, t, Y, **kwargs):
    """Integrator with the reference's protocol (R:_base_classes.py:180-182)
    that drives scipy's RK45: the like-for-like CPU algorithm for kernel K1."""
    from scipy.integrate import solve_ivp

    return solve_ivp(f, t, Y, method="RK45", **kwargs).y.T
