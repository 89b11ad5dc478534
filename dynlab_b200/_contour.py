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
* vertices are linear interpolations along cell edges.

These rules reproduce the line sets, vertex order and coordinates of the reference's own
golden tests (R:tests/test_diagnostics/test_lagrangian.py:21-37,
R:tests/test_diagnostics/test_eulerian.py:127-181) from the reference's ridge fields — see
tests/test_contour.py.  contourpy itself is not installed in the build image, so configurations
those goldens do not exercise (several starts in one cell, closed-loop start vertex) follow
this module's own deterministic choice.

``marching_squares`` / ``link_segments`` are the all-numpy twins of the two native stages; they
exist as the cross-check the tests compare the native path against and are not called by the
diagnostics.
"""
from __future__ import annotations

import numpy as np


def _cell_segments(zd, mask, above, level):
    """Vectorised per-cell case analysis.  Returns arrays (cell_j, cell_i, a0, a1, b0, b1):
    one row per contour segment, running from the crossing on edge (a0, a1) to the crossing on
    edge (b0, b1); nodes are flat indices j * nx + i."""
    ny, nx = zd.shape
    J, I = np.meshgrid(np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
    # corners counter-clockwise: (j,i) (j,i+1) (j+1,i+1) (j+1,i)
    cj = np.stack([J, J, J + 1, J + 1], axis=-1)
    ci = np.stack([I, I + 1, I + 1, I], axis=-1)
    node = cj * nx + ci
    m = mask[cj, ci]
    up = above[cj, ci]
    nmask = m.sum(axis=-1)
    out = []

    def emit(sel, ka, kb, la, lb):
        """cells `sel` (boolean [ny-1, nx-1]): segment from edge (corner ka -> la) to edge
        (corner kb -> lb); corner indices are per-cell arrays or scalars."""
        if not sel.any():
            return
        jj, ii = J[sel], I[sel]
        nd = node[sel]
        r = np.arange(nd.shape[0])

        def pick(k):
            return nd[r, k[sel]] if isinstance(k, np.ndarray) else nd[:, k]

        out.append(np.stack([jj, ii, pick(ka), pick(la), pick(kb), pick(lb)], axis=1))

    # ---- full quads
    quad = nmask == 0
    for k in range(4):  # a segment starts on the CCW edge k -> k+1 where upper -> lower
        k1 = (k + 1) % 4
        start = quad & up[..., k] & ~up[..., k1]
        if not start.any():
            continue
        n_up = up.sum(axis=-1)
        saddle = start & (n_up == 2) & (up[..., (k + 2) % 4]) & ~up[..., (k + 3) % 4]
        simple = start & ~saddle
        # simple: the end edge is the only lower -> upper edge; find it per cell
        for e in range(4):
            e1 = (e + 1) % 4
            emit(simple & ~up[..., e] & up[..., e1], k, e, k1, e1)
        if saddle.any():
            zmid = 0.25 * zd[cj, ci].sum(axis=-1)
            mid_up = zmid > level
            # centre lower: the upper corner k is cut off -> ends on edge k-1 -> k
            emit(saddle & ~mid_up, k, (k + 3) % 4, k1, k)
            # centre upper: the lower corner k+1 is cut off -> ends on edge k+1 -> k+2
            emit(saddle & mid_up, k, k1, k1, (k + 2) % 4)
    # ---- corner-masked triangles
    for c in range(4):
        tri = (nmask == 1) & m[..., c]
        if not tri.any():
            continue
        v = [(c + 1) % 4, (c + 2) % 4, (c + 3) % 4]  # remaining corners, still CCW
        for s in range(3):
            a, a1 = v[s], v[(s + 1) % 3]
            start = tri & up[..., a] & ~up[..., a1]
            for e in range(3):
                b, b1 = v[e], v[(e + 1) % 3]
                emit(start & ~up[..., b] & up[..., b1], a, b, a1, b1)
    if not out:
        return np.zeros((0, 6), dtype=np.int64)
    return np.concatenate(out, axis=0)


def marching_squares(x, y, z, level=0.0):
    """Trace the ``level`` contour of ``z[ydim, xdim]`` (masked array or ndarray) entirely on the
    host: numpy case analysis of every cell, then polyline linking."""
    rec = host_segments(x, y, z, level)
    return link_segments(*rec) if rec is not None else []


def host_segments(x, y, z, level=0.0):
    """numpy twin of csrc/contour.cu: segment records (cell, start_key, end_key, pts[n, 4]) of
    the ``level`` contour, or None when there are none."""
    zd = np.ma.getdata(z).astype(float)
    mask = np.ma.getmaskarray(z) | ~np.isfinite(zd)
    ny, nx = zd.shape
    if ny < 2 or nx < 2:
        return None
    above = (zd > level) & ~mask
    segs = _cell_segments(zd, mask, above, level)
    if segs.shape[0] == 0:
        return None
    nn = nx * ny

    def edge_key(p, q):
        return np.minimum(p, q) * nn + np.maximum(p, q)

    # vertex on an edge (p, q): linear interpolation to z == level
    def vertex(p, q):
        zp, zq = zd.flat[p], zd.flat[q]
        f = (level - zp) / (zq - zp)
        xp, xq = x[p % nx], x[q % nx]
        yp, yq = y[p // nx], y[q // nx]
        return xp + f * (xq - xp), yp + f * (yq - yp)

    sx, sy = vertex(segs[:, 2], segs[:, 3])
    ex, ey = vertex(segs[:, 4], segs[:, 5])
    cell = segs[:, 0] * (nx - 1) + segs[:, 1]
    return (cell, edge_key(segs[:, 2], segs[:, 3]), edge_key(segs[:, 4], segs[:, 5]),
            np.stack([sx, sy, ex, ey], axis=1))


def link_segments(cell, start_key, end_key, pts):
    """Chain contour segments into polylines.  ``cell`` is the row-major index of the cell a
    segment lies in, ``start_key`` / ``end_key`` identify the cell edges it starts / ends on
    (shared edges have equal keys), ``pts[n, 4]`` holds its two vertices.  Lines come out in
    the scan order of the cell they start in; closed loops repeat their first vertex."""
    n = len(cell)
    if n == 0:
        return []
    order = np.lexsort((start_key, cell))  # scan order of the owning cell, deterministic inside
    cell, start_key, end_key, pts = cell[order], start_key[order], end_key[order], pts[order]
    sorter = np.argsort(start_key, kind="stable")
    sk_sorted = start_key[sorter]
    pos = np.searchsorted(sk_sorted, end_key)
    pos_c = np.minimum(pos, n - 1)
    has_next = (pos < n) & (sk_sorted[pos_c] == end_key)
    nxt = np.where(has_next, sorter[pos_c], -1)
    has_prev = np.zeros(n, dtype=bool)
    has_prev[nxt[has_next]] = True

    nxt_l = nxt.tolist()
    used = np.zeros(n, dtype=bool)
    heads = [int(s) for s in np.flatnonzero(~has_prev)]  # open lines start where nothing leads in
    for s in heads:
        k = s
        while k != -1 and not used[k]:
            used[k] = True
            k = nxt_l[k]
    for k0 in range(n):  # closed loops: whatever is left, entered at their first cell in scan order
        if not used[k0]:
            heads.append(k0)
            k = k0
            while not used[k]:
                used[k] = True
                k = nxt_l[k]
    heads.sort()
    lines = []
    for s in heads:
        idx = [s]
        k = nxt_l[s]
        while k != -1 and k != s:
            idx.append(k)
            k = nxt_l[k]
        idx = np.asarray(idx)
        px = np.concatenate([[pts[s, 0]], pts[idx, 2]])
        py = np.concatenate([[pts[s, 1]], pts[idx, 3]])
        lines.append(np.stack([px, py], axis=1))
    return lines


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
