"""FTLE and LCS (mirror of R:src/dynlab/diagnostics/lagrangian.py).

Pipeline per ``compute``: K1 (flow map of this rank's row slab) -> halo rows -> K2 (np.gradient
stencil + Cauchy-Green tensor + eigen + log, fused) -> gather -> one device-to-host copy of the
field.  ``flow_map`` stays on the device until it is read.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as nat
from .. import sharding
from .._engine import Engine
from ..utils import odeint_wrapper
from ._base_classes import LagrangianDiagnostic2D, RidgeExtractor2D


def _ftle_on_device(diag, local, slab, t, edge_order, xi_mode):
    """K2 on this rank's slab, gathered to the full grid on every rank (device tensors)."""
    eng = Engine.get()
    rank, size = sharding.world()
    if edge_order > 2:
        raise ValueError("'edge_order' greater than 2 not supported")
    t_span = abs(float(t[-1]) - float(t[0]))
    if t_span == 0.0:
        # the reference evaluates 1.0 / (2.0 * abs(t[-1] - t[0])) in Python (R:lagrangian.py:71)
        raise ZeroDivisionError("float division by zero")
    if size == 1:
        return eng.ftle_stencil(local, 0, diag._xf, diag._yf, edge_order, t_span, 0, diag.ydim,
                                xi_mode)
    out = torch.empty(slab.rows_max, diag.xdim, dtype=torch.float64, device=eng.device)
    xi = None
    if slab.rows:
        _, xi = eng.ftle_stencil(local, slab.lo, diag._xf, diag._yf, edge_order, t_span,
                                 slab.row0, slab.rows, xi_mode, out=out[:slab.rows])
    field = sharding.gather_rows(out, diag.ydim, size)
    if xi_mode != nat.XI_NONE:
        pad = torch.empty(slab.rows_max, diag.xdim, 2, dtype=torch.float64, device=eng.device)
        if xi is not None:
            pad[:slab.rows] = xi
        xi = sharding.gather_rows(pad, diag.ydim, size)
    return field, xi


def pipeline_schedule(lo, hi, own0, own1, ydim, edge_order, chunks, last_fraction):
    """Chunk plan of ``FTLE._pipeline``: a list of ``(r0, r1, done, upto)`` - integrate global rows
    [r0, r1), then produce the FTLE rows [done, upto) (empty when ``upto == done``).  The chunks
    tile [lo, hi) in order (equal chunks and a short last one, whose copy is the only one left
    exposed; small slabs: one chunk), the output ranges tile [own0, own1) in order, and every
    row an output row's stencil reads - its neighbours, and rows 0..2 / ydim-3..ydim-1 for the
    one-sided edge_order-2 formulas at the global edges - lies in [lo, r1)."""
    nloc = hi - lo
    if chunks > 1 and nloc >= 64 * chunks:
        last = max(64, nloc // last_fraction)
        bounds = [lo + (nloc - last) * c // (chunks - 1) for c in range(chunks)] + [hi]
    else:
        bounds = [lo, hi]
    steps, done = [], own0
    for r0, r1 in zip(bounds[:-1], bounds[1:]):
        upto = own1 if r1 == hi else min(own1, r1 - 1)  # row r1-1 still needs row r1
        if done == 0 and edge_order == 2 and r1 < min(3, ydim):
            upto = done  # the one-sided formula of row 0 reads rows 0..2
        upto = max(upto, done)
        steps.append((r0, r1, done, upto))
        done = upto
    return steps


class FTLE(LagrangianDiagnostic2D):
    """Finite-time Lyapunov exponent field of a 2-D flow (R:lagrangian.py:9-78)."""

    def __init__(self, integrator=odeint_wrapper, num_threads: int = 1) -> None:
        super().__init__(integrator=integrator, num_threads=num_threads)

    def compute(self, x, y, f, t, edge_order: int = 1, **kwargs):
        """Returns ``np.ma.MaskedArray[ydim, xdim]``; sets ``x, y, xdim, ydim, flow_map,
        field`` like the reference.  ``kwargs`` (rtol, atol, max_step, first_step) go to the
        integrator."""
        if (sharding.world()[1] == 1 and self.pipeline_chunks > 1
                and np.size(y) >= 64 * self.pipeline_chunks and np.size(x) * np.size(y) >= (1 << 22)):
            return self._compute_pipelined(self._plan(x, y, f, t, kwargs), t, edge_order)
        if sharding.host_result_shared():
            return self._compute_shared(x, y, f, t, edge_order, kwargs)
        field = self._run(x, y, f, t, edge_order, self.gather_flow_map, kwargs)
        if sharding.host_result_wanted():
            # several ranks copying at once: the staged path's worker threads and page faults
            # contend for the host cores, one DMA into a pinned buffer per rank does not
            many = sharding.world()[1] > 1
            self.field = np.ma.MaskedArray(Engine.get().to_host(field, pinned=True if many else None))
        else:  # sharding.set_host_result("rank0") on a non-zero rank: placeholder, all masked
            torch.cuda.current_stream(field.device).synchronize()
            self.field = np.ma.MaskedArray(np.broadcast_to(np.nan, tuple(field.shape)), mask=True)
        self._collect_stats()
        return self.field

    # Large grids: K1 runs in row chunks; as soon as a chunk's successor row exists K2 produces
    # that chunk's FTLE rows and a copy stream moves them to pinned host memory while the next
    # chunk integrates, so the device-to-host copy (268 MB at config 3) hides behind K1.  Results
    # are bit-identical to the one-shot path (seeds are independent and the stencil of a row
    # only reads its neighbour rows).
    pipeline_chunks = 4
    pipeline_last_fraction = 32  # the last chunk is 1/32 of the rows (chunks 3-8, fractions 16-64: all within 0.5 %)

    def _pipeline(self, plan, t, edge_order, lo, hi, own0, own1, host):
        """Integrate global rows [lo, hi) in chunks and stream the FTLE rows [own0, own1) into
        ``host[own0:own1]`` (a pinned torch tensor over the whole grid) as they become
        computable.  [lo, hi) must contain every row the stencils of [own0, own1) read."""
        if edge_order > 2:
            raise ValueError("'edge_order' greater than 2 not supported")
        t_span = abs(float(t[-1]) - float(t[0]))
        if t_span == 0.0:
            raise ZeroDivisionError("float division by zero")  # R:lagrangian.py:71
        eng = Engine.get()
        nloc, xdim = hi - lo, self.xdim
        steps = pipeline_schedule(lo, hi, own0, own1, self.ydim, edge_order, self.pipeline_chunks,
                                  self.pipeline_last_fraction)
        fm = eng.empty(nloc, xdim, 2)
        field = eng.empty(max(own1 - own0, 1), xdim)
        counters = torch.zeros(len(steps), nat.CNT_WORDS, dtype=torch.int64, device=eng.device)
        main = torch.cuda.current_stream(eng.device)
        for c, (r0, r1, done, upto) in enumerate(steps):
            self._integrate_rows(plan, r0, r1 - r0, fm[r0 - lo:r1 - lo], counters=counters[c])
            if upto > done:
                eng.ftle_stencil(fm[:r1 - lo], lo, self._xf, self._yf, edge_order, t_span, done,
                                 upto - done, out=field[done - own0:upto - own0])
                ev = torch.cuda.Event()
                ev.record(main)
                eng.copy_stream.wait_event(ev)
                with torch.cuda.stream(eng.copy_stream):
                    host[done:upto].copy_(field[done - own0:upto - own0], non_blocking=True)
        eng.copy_stream.synchronize()
        field.record_stream(eng.copy_stream)
        self._counters_pending, self._counter_tensors = True, counters
        return fm, field

    def _compute_pipelined(self, plan, t, edge_order):
        """Single GPU: the whole grid, result in a fresh pinned buffer."""
        ydim, xdim = self.ydim, self.xdim
        host = torch.empty((ydim, xdim), dtype=torch.float64, pin_memory=True)
        fm, field = self._pipeline(plan, t, edge_order, 0, ydim, 0, ydim, host)
        self.flow_map = fm
        self.field_device = field
        self.field = np.ma.MaskedArray(host.numpy())
        self._collect_stats()
        return self.field

    def _compute_shared(self, x, y, f, t, edge_order, kwargs):
        """Multi-GPU with ``sharding.set_host_result("shared")``: this rank's slab goes through
        the same chunk pipeline straight into the host segment all ranks of the box map - no
        device all-gather, the field crosses PCIe once in total, and the copies (bounded by the
        box's aggregate device-to-host bandwidth, ~100 GB/s for 8 GPUs) hide behind K1.  The
        one or two halo rows are integrated locally (bit-identical to the neighbours' rows), so
        chunks do not wait for other ranks.  ``flow_map`` / ``field_device`` hold this rank's
        rows only."""
        plan = self._plan(x, y, f, t, kwargs)
        rank, size = sharding.world()
        slab = sharding.partition(self.ydim, size, rank)
        host_np, host_t = sharding.shared_host_array((self.ydim, self.xdim))
        own = None
        self._counters_pending, self._counter_tensors = False, None
        if slab.rows:
            fm, own = self._pipeline(plan, t, edge_order, slab.lo, slab.hi, slab.row0,
                                     slab.row0 + slab.rows, host_t)
            self.flow_map = fm[slab.halo_lo:slab.halo_lo + slab.rows]
        else:
            if edge_order > 2:
                raise ValueError("'edge_order' greater than 2 not supported")
            if float(t[-1]) == float(t[0]):
                raise ZeroDivisionError("float division by zero")  # R:lagrangian.py:71
        sharding.host_barrier()  # every rank's rows are in the segment
        self.field_device = own
        self.field = np.ma.MaskedArray(host_np)
        self._collect_stats()
        return self.field

    def compute_device(self, x, y, f, t, edge_order: int = 1, **kwargs):
        """Same computation, result left on the device (``torch.Tensor[ydim, xdim]``, also kept
        as ``field_device``); nothing is synchronised or copied to the host.  Under multi-GPU
        sharding the flow map is not gathered: ``flow_map`` holds this rank's rows only."""
        return self._run(x, y, f, t, edge_order, False, kwargs)

    def _run(self, x, y, f, t, edge_order, gather_flow_map, kwargs):
        keep, self.gather_flow_map = self.gather_flow_map, gather_flow_map
        try:
            local, slab = super().compute(x, y, f, t, **kwargs)
        finally:
            self.gather_flow_map = keep
        field, _ = _ftle_on_device(self, local, slab, t, edge_order, nat.XI_NONE)
        self.field_device = field
        return field


class LCS(LagrangianDiagnostic2D, RidgeExtractor2D):
    """Lagrangian coherent structures: FTLE ridges (R:lagrangian.py:81-182)."""

    def __init__(self, integrator=odeint_wrapper, num_threads: int = 1) -> None:
        super().__init__(integrator=integrator, num_threads=num_threads)

    def compute(self, x, y, f, t, edge_order: int = 1, percentile: float = None,
                force_eigenvectors: bool = False, debug: bool = False, **kwargs):
        """Returns the list of ridge polylines ``[(n, 2) ndarray, ...]``; sets ``ftle`` and
        ``lcs`` (and ``Xi_max``, ``directional_derivative``, ``concavity`` when ``debug``)."""
        local, slab = super().compute(x, y, f, t, **kwargs)
        xi_mode = nat.XI_FORCED if force_eigenvectors else nat.XI_LAPACK
        field, xi = _ftle_on_device(self, local, slab, t, edge_order, xi_mode)
        del self.flow_map  # R:lagrangian.py:137-139
        eng = Engine.get()
        self.ftle = np.ma.MaskedArray(eng.to_host(field))
        self._collect_stats()
        self.lcs = self.extract_ridges(field, xi, percentile, edge_order, debug)
        if debug:
            self.Xi_max = np.ma.MaskedArray(eng.to_host(xi))
        return self.lcs
