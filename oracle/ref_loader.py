"""TEST INFRASTRUCTURE — loader for the real hokiepete/dynlab reference.

Only usable where ``/root/reference`` exists (the build container).  Nothing in
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may import this module: the GPU box
has no reference tree.  It is used by ``oracle/make_golden.py`` to generate the
fixtures under ``tests/golden/`` and by ``tests/test_oracle_vs_reference.py``
(skipped when the tree is absent) to pin ``oracle/py_oracle.py``.

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

REFERENCE_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "dynlab"))


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
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_SRC)
    _install_shims()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
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
