"""dynlab_b200 — B200-native drop-in for dynlab's FTLE / Eulerian hot path.

    from dynlab_b200.diagnostics import FTLE
    from dynlab_b200.flows import double_gyre
    ftle = FTLE().compute(x, y, double_gyre, (10, 0), edge_order=2, rtol=1e-8, atol=1e-8)

The public surface mirrors hokiepete/dynlab (R:src/dynlab/diagnostics/__init__.py:1-4,
R:src/dynlab/flows.py, R:src/dynlab/utils.py); the arithmetic runs in hand-written sm_100a
CUDA kernels behind the C-ABI of ``include/dynlab_b200.h``.  ``install_as_dynlab()`` registers
the package under the name ``dynlab`` so unmodified user scripts import it.
"""
from __future__ import annotations

import sys

__version__ = "0.1.0"


def install_as_dynlab() -> None:
    """Make ``import dynlab`` / ``from dynlab.diagnostics import FTLE`` resolve to this package."""
    from . import diagnostics, flows, utils  # noqa: F401  (loads the native library)

    me = sys.modules[__name__]
    sys.modules["dynlab"] = me
    sys.modules["dynlab.diagnostics"] = diagnostics
    sys.modules["dynlab.flows"] = flows
    sys.modules["dynlab.utils"] = utils
