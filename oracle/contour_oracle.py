"""TEST INFRASTRUCTURE — all-numpy zero-contour tracer (twin of csrc/contour.cu + csrc/link.cu).

The reference calls ``contourpy.contour_generator(x=, y=, z=masked).lines(0)``
(R:src/dynlab/diagnostics/_base_classes.py:270).  contourpy is a third-party dependency of the
reference (unpinned, R:pyproject.toml:21) that is absent from this image, so its published
conventions are restated here (serial algorithm, ``corner_mask`` on, ``LineType.Separate``; see
dynlab_b200/_contour.py for the list) and pinned on the reference's own golden ridge lines
(tests/test_contour.py).  The product path is the device cell pass + C++ linker; this module is
what the tests compare it with.  Only ``tests/`` may import it.

Parity status: pinned by the three golden line sets of the reference's tests (8 decimals);
beyond those (several line starts in one cell, start vertex of closed loops, the last bits of
the interpolated vertices - contourpy's own expression is used, operand order upper node
first) "parity unpinned".
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

    # vertex on an edge: contourpy's interp() (src/base_impl.h, linear z) with point0 = the
    # edge's upper node, point1 = its lower node:
    #   frac = (z1 - level) / (z1 - z0);  v = v0 * frac + v1 * (1 - frac)
    def vertex(p0, p1):
        z0, z1 = zd.flat[p0], zd.flat[p1]
        frac = (z1 - level) / (z1 - z0)
        x0, x1 = x[p0 % nx], x[p1 % nx]
        y0, y1 = y[p0 // nx], y[p1 // nx]
        return x0 * frac + x1 * (1.0 - frac), y0 * frac + y1 * (1.0 - frac)

    sx, sy = vertex(segs[:, 2], segs[:, 3])  # start edge runs upper -> lower
    ex, ey = vertex(segs[:, 5], segs[:, 4])  # end edge runs lower -> upper
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
