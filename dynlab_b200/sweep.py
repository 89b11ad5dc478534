"""Time-sweep driver (BASELINE config 5): FTLE (and optionally the LCS ridge
field) for a list of time intervals on one grid.

The reference has no such entry point — its users loop over ``FTLE().compute(...)``
(R:README.md:19-28).  Slices are independent, so they are sharded round-robin over the ranks of
the process group with *no* data-path communication (each rank keeps whole slices; the row
sharding of ``sharding.py`` is switched off inside the sweep); results stay on the device
unless ``to_host`` is set.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _native as nat
from . import sharding
from ._contour import contour_lines_finish
from ._engine import Engine, as_coord
from ._resolve import resolve_flow
from .utils import ODEINT_DEFAULT_TOL, parse_integrator_kwargs


def handle_cap_grew(eng, cbufs) -> bool:
    """True when a previous slice needed more contour records than the reusable buffers hold."""
    return eng._contour_cap > cbufs[0]


def ftle_time_sweep(x, y, f, t_spans, edge_order: int = 1, percentile=None, ridge: bool = False,
                    to_host: bool = True, **kwargs):
    """FTLE fields for ``t_spans = [(t0, tf), ...]``.

    Returns ``{slice_index: result}`` for the slices this rank owns (all of them without a
    process group).  ``result`` is the FTLE field (``np.ma.MaskedArray`` or a device tensor);
    with ``ridge=True`` it is ``(ftle, directional_derivative, lines)`` (host) or ``(ftle,
    directional_derivative, ridge_mask, lines)`` (device tensors + host polylines): the ridge
    field of ``LCS`` (R:_base_classes.py:216-268) with ``percentile`` applied and its zero
    contour (device cell classification + host linking), i.e. what ``LCS.compute`` returns.
    """
    xs, ys = as_coord(np.asarray(x).squeeze()), as_coord(np.asarray(y).squeeze())
    if xs.ndim != 1 or ys.ndim != 1:
        raise ValueError("x and y must be 1-dimensional arrays.")
    kw = parse_integrator_kwargs(kwargs, ODEINT_DEFAULT_TOL, ODEINT_DEFAULT_TOL)
    _, fid, dim, params = resolve_flow(f)
    if dim != 2:
        raise ValueError("a 2-D diagnostic needs a 2-D flow")
    eng = Engine.get()
    if dist.is_available() and dist.is_initialized() and sharding._mode != "off":
        rank, size = dist.get_rank(), dist.get_world_size()
    else:
        rank, size = 0, 1
    ydim, xdim = len(ys), len(xs)
    out = {}
    fm = eng.empty(ydim, xdim, 2)
    pending = None  # (slice, contour handle, device results): linked while the next slice integrates
    # All device memory of the sweep is taken here, once: the results of every owned slice as
    # stacked tensors (they are handed out as views), the temporaries as one reusable set.  With
    # per-slice allocations torch's caching allocator had to cudaMalloc / cudaFree result-sized
    # blocks in the middle of the sweep (freed xi / scratch blocks get split by the retained
    # results), and cudaFree waits for the running K1: stalls of 80-250 ms on some slices.
    owned = [k for k in range(len(t_spans)) if k % size == rank]
    n_keep = len(owned) if not to_host else 1
    fields = eng.empty(max(n_keep, 1), ydim, xdim)
    if ridge:
        dds = eng.empty(max(n_keep, 1), ydim, xdim)
        masks = eng.empty(max(n_keep, 1), ydim, xdim, dtype=torch.uint8)
        xi_buf, conc_buf, scratch = eng.empty(ydim, xdim, 2), eng.empty(ydim, xdim), eng.empty(ydim, xdim, 2)
        cbufs = eng.contour_buffers(xdim, ydim)

    def finish(p):
        k, handle, field, dd, mask = p
        lines = contour_lines_finish(eng, handle)
        out[k] = (field, dd, lines) if to_host else (field, dd, mask, lines)

    for k, span in enumerate(t_spans):
        if k % size != rank:
            continue
        if len(span) != 2:
            raise ValueError("t must only have 2 values, t_0 and t_final")
        t0, tf = float(span[0]), float(span[1])
        eng.flow_map(fid, params, xs, ys, 0, ydim, t0, tf, kw["rtol"], float(kw["atol"]),
                     kw["max_step"], kw["first_step"] or 0.0, out=fm)
        if pending is not None:  # host linking of the previous slice overlaps this slice's K1
            finish(pending)
            pending = None
        xi_mode = nat.XI_LAPACK if ridge else nat.XI_NONE
        slot = owned.index(k) if not to_host else 0
        field, xi = eng.ftle_stencil(fm, 0, xs, ys, edge_order, abs(tf - t0), 0, ydim, xi_mode,
                                     out=fields[slot], xi_out=xi_buf if ridge else None)
        if not ridge:
            out[k] = np.ma.MaskedArray(eng.to_host(field)) if to_host else field
            continue
        thr = None
        if percentile:
            # R:_base_classes.py:266-268 — np.percentile of the (unmasked) field values
            thr = eng.percentile(field, percentile)
        dd, _, mask = eng.ridge_field(field, xi, xs, ys, edge_order, thr,
                                      bufs=(dds[slot], conc_buf, masks[slot], scratch))
        if handle_cap_grew(eng, cbufs):
            cbufs = eng.contour_buffers(xdim, ydim)
        handle = eng.contour_launch(dd, mask, xs, ys, bufs=cbufs)
        if to_host:
            field = np.ma.MaskedArray(eng.to_host(field))
            dd = np.ma.MaskedArray(eng.to_host(dd), mask=eng.to_host(mask).astype(bool))
        pending = (k, handle, field, dd, mask)
    if pending is not None:
        finish(pending)
    return out
