"""Zero-contour extraction of a masked grid (host side of RidgeExtractor2D.extract_ridges).

The reference calls ``contourpy.contour_generator(x=, y=, z=masked).lines(0)``
(R:src/dynlab/diagnostics/_base_classes.py:270).  When contourpy is importable it is used
directly (it is the reference's own dependency); otherwise the marching-squares tracer below
produces the polylines.  Output: ``list[np.ndarray[n, 2]]`` of (x, y) vertices, like
``LineType.Separate``.

Parity note: contourpy is not installed in the build image, so vertex *positions* (linear
interpolation on cell edges) are checked against the reference's golden vectors, but the
ordering of lines and of vertices inside a line follows this tracer's own deterministic rule
(boundary-started lines first, in row-major order of their first cell; then closed loops).
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - not available in the build image
    from contourpy import contour_generator as _contourpy_generator
except Exception:  # noqa: BLE001
    _contourpy_generator = None


def zero_contour_lines(x, y, z, level: float = 0.0):
    if _contourpy_generator is not None:  # pragma: no cover
        return _contourpy_generator(x=x, y=y, z=z).lines(level)
    return marching_squares(np.asarray(x, dtype=float), np.asarray(y, dtype=float), z, level)


def _interp(p0, p1, z0, z1, level):
    f = (level - z0) / (z1 - z0)
    return p0 + f * (p1 - p0)


def marching_squares(x, y, z, level=0.0):
    """Trace the ``level`` contour of ``z[ydim, xdim]`` (masked array or ndarray).

    A quad contributes only if none of its four corners is masked.  Saddle quads are split by
    the sign of the centre value (mean of the corners).  Segments are oriented with the higher
    values on the left and chained through shared cell edges."""
    zd = np.ma.getdata(z).astype(float)
    mask = np.ma.getmaskarray(z) | ~np.isfinite(zd)
    ny, nx = zd.shape
    above = zd > level

    # edge ids: horizontal edge (i, j)-(i, j+1) -> ('h', i, j); vertical (i, j)-(i+1, j) -> ('v', i, j)
    def edge_point(e):
        kind, i, j = e
        if kind == 'h':
            px = _interp(x[j], x[j + 1], zd[i, j], zd[i, j + 1], level)
            return (px, y[i])
        py = _interp(y[i], y[i + 1], zd[i, j], zd[i + 1, j], level)
        return (x[j], py)

    # segments: list of (edge_from, edge_to); next_of[edge_from] -> list of edge_to
    starts = {}
    ends = {}
    segs = []
    for i in range(ny - 1):
        for j in range(nx - 1):
            if mask[i, j] or mask[i, j + 1] or mask[i + 1, j] or mask[i + 1, j + 1]:
                continue
            # corners: 0=(i,j) 1=(i,j+1) 2=(i+1,j+1) 3=(i+1,j); edges between consecutive corners
            a = (above[i, j], above[i, j + 1], above[i + 1, j + 1], above[i + 1, j])
            n_above = sum(a)
            if n_above in (0, 4):
                continue
            e = (('h', i, j), ('v', i, j + 1), ('h', i + 1, j), ('v', i, j))  # bottom,right,top,left
            # walking the quad counter-clockwise (0->1->2->3), a contour enters on an edge where
            # the corners go above->below and leaves where they go below->above: higher on left
            ins = [k for k in range(4) if a[k] and not a[(k + 1) % 4]]
            outs = [k for k in range(4) if not a[k] and a[(k + 1) % 4]]
            if len(ins) == 1:
                pairs = [(ins[0], outs[0])]
            else:  # saddle
                centre_above = 0.25 * (zd[i, j] + zd[i, j + 1] + zd[i + 1, j + 1] + zd[i + 1, j]) > level
                pairs = []
                for k in ins:
                    # the 'in' edge k pairs with the out edge that encloses the above corner k
                    # (centre below) or goes around through the centre (centre above)
                    cand = (k - 1) % 4 if not centre_above else (k + 1) % 4
                    pairs.append((k, cand))
            for ki, ko in pairs:
                s = (e[ki], e[ko])
                segs.append(s)
                starts.setdefault(e[ki], []).append(len(segs) - 1)
                ends.setdefault(e[ko], []).append(len(segs) - 1)

    used = [False] * len(segs)
    lines = []

    def follow(first):
        pts = [edge_point(segs[first][0]), edge_point(segs[first][1])]
        used[first] = True
        cur = segs[first][1]
        while True:
            nxt = [k for k in starts.get(cur, []) if not used[k]]
            if not nxt:
                break
            k = nxt[0]
            used[k] = True
            cur = segs[k][1]
            pts.append(edge_point(cur))
        return np.array(pts)

    # open lines first: segments whose entry edge is not the exit edge of another segment
    for k, s in enumerate(segs):
        if not used[k] and not ends.get(s[0]):
            lines.append(follow(k))
    for k in range(len(segs)):  # closed loops
        if not used[k]:
            lines.append(follow(k))
    return lines
