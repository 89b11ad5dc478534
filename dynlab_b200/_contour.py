"""Zero-contour extraction of a masked grid (last step of RidgeExtractor2D.extract_ridges).

The reference calls ``contourpy.contour_generator(x=, y=, z=masked).lines(0)``
(R:src/dynlab/diagnostics/_base_classes.py:270): default "serial" algorithm, ``corner_mask``
on (z is a masked array), ``LineType.Separate`` -> ``list[np.ndarray[n, 2]]`` of (x, y) vertices.
The product path is ``zero_contour_lines_device``: cell classification on the device
(csrc/contour.cu), polyline linking in C++ (csrc/link.cu), with contourpy's conventions:

* a node is "upper" iff ``z > level``;
* a quad contributes if none of its corners is masked; a quad with exactly one masked corner
  contributes the triangle of the other three (corner masking); two or more: nothing;
* saddle quads are resolved by the mean of the four corners (``zmid > level``);
* lines are oriented with the upper region on their left; an open line starts where it enters
  the contoured region (domain edge, mask edge or triangle diagonal); closed loops repeat
  their first vertex at the end;
* lines are emitted in the row-major scan order (y outer, x inner) of the cell they start in;
* vertices are linear interpolations along cell edges, evaluated with contourpy's own
  expression ``v0 * frac + v1 * (1 - frac)``, ``frac = (z1 - level) / (z1 - z0)`` (point0 = the
  edge's upper node).

These rules reproduce the line sets, vertex order and coordinates of the reference's own
golden tests (R:tests/test_diagnostics/test_lagrangian.py:21-37,
R:tests/test_diagnostics/test_eulerian.py:127-181) from the reference's ridge fields — see
tests/test_contour.py.  contourpy itself is not installed in the build image, so configurations
those goldens do not exercise (several starts in one cell, closed-loop start vertex) follow
this module's own deterministic choice.

The all-numpy twins of the two native stages (``marching_squares``, ``link_segments``) are test
infrastructure and live in ``oracle/contour_oracle.py``.
"""
from __future__ import annotations

import numpy as np


def link_segments_native(cell, start_key, end_key, pts, xdim, ydim):
    """``link_segments`` in C++ (csrc/link.cu, ``dlb_link_segments``): same rules, same output,
    without the per-segment Python loop."""
    from . import _native as nat

    n = len(cell)
    if n == 0:
        return []
    cell = np.ascontiguousarray(cell, dtype=np.int64)
    ka = np.ascontiguousarray(start_key, dtype=np.int64)
    kb = np.ascontiguousarray(end_key, dtype=np.int64)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    start = np.empty(n + 1, dtype=np.int64)
    verts = np.empty((2 * n, 2), dtype=np.float64)
    nl = np.zeros(1, dtype=np.int64)
    nat.check(nat.lib.dlb_link_segments(cell.ctypes.data, ka.ctypes.data, kb.ctypes.data,
                                        pts.ctypes.data, n, int(xdim), int(ydim), start.ctypes.data,
                                        verts.ctypes.data,
                                        nl.ctypes.data))
    b = start[:int(nl[0]) + 1].tolist()
    verts = verts[:b[-1]].copy()  # drop the unused tail; lines are views of this block
    return [verts[b[i]:b[i + 1]] for i in range(len(b) - 1)]


def zero_contour_lines_device(eng, z_dev, mask_dev, x, y, level: float = 0.0):
    """Device case analysis (csrc/contour.cu) + host linking (csrc/link.cu).  ``z_dev`` /
    ``mask_dev`` are the device tensors of the (masked) field; nothing but the O(ridge length)
    segment records crosses PCIe."""
    return contour_lines_finish(eng, eng.contour_launch(z_dev, mask_dev, x, y, level))


def contour_lines_finish(eng, handle):
    """Second half of ``zero_contour_lines_device`` for callers that enqueue more device work
    between the launch (``Engine.contour_launch``) and the linking."""
    cell, ka, kb, pts = eng.contour_fetch(handle)
    return link_segments_native(cell, ka, kb, pts, len(handle["x"]), len(handle["y"]))
