"""Host-side bases of the diagnostics (mirror of R:src/dynlab/diagnostics/_base_classes.py).

Same class names, constructor/compute signatures, attribute names and exception types as the
reference; the work itself is enqueued on the GPU through ``dynlab_b200._engine.Engine``:

  LagrangianDiagnostic2D.compute  -> K1 flow-map kernel over this rank's row slab
                                     (reference: seed loop + integrator, :171-189)
  EulerianDiagnostic2D.compute    -> device-resident (u, v) — uploaded, or evaluated by K0 —
                                     handed to the K3 stencil (reference: np.gradient x2, :91-113)
  RidgeExtractor2D.extract_ridges -> ridge-field kernel, then host contouring (:194-277)
"""
from __future__ import annotations

import warnings
from abc import ABC, abstractmethod

import numpy as np
import torch

from .. import _native as nat
from .. import sharding
from .._engine import Engine, as_coord
from .._resolve import resolve_flow
from ..utils import (ODEINT_DEFAULT_TOL, RK45_DEFAULT_ATOL, RK45_DEFAULT_RTOL, odeint_wrapper,
                     parse_integrator_kwargs, rk45_wrapper)


class UnsupportedIntegratorError(NotImplementedError):
    """``integrator=`` is not one of the device integrator sentinels."""


class _Deferred:
    """Device tensor that is only computed if somebody asks for it (``fn`` returns a tuple of
    tensors, ``index`` selects one; the tuple is computed once and shared)."""

    def __init__(self, fn, index, cache):
        self.fn, self.index, self.cache = fn, index, cache

    def resolve(self):
        if "value" not in self.cache:
            self.cache["value"] = self.fn()
        return self.cache["value"][self.index]


class _FlowSource:
    """Stands for ``f(t, np.meshgrid(x, y))`` of a flow with a device twin: the Eulerian kernels
    evaluate it on the fly (csrc/eulerflow.cu) instead of reading ``u, v`` arrays."""

    def __init__(self, fid, params, t):
        self.fid, self.params, self.t = fid, params, float(t)


class _HostSlab:
    """Rows [lo, lo + rows) of a host field, already on the device (multi-GPU Eulerian intake)."""

    def __init__(self, rows: torch.Tensor, lo: int):
        self.rows, self.lo = rows, lo

    def __getitem__(self, key):  # ud[slab.lo:slab.hi] in _euler
        assert key.start == self.lo and key.stop == self.lo + self.rows.shape[0]
        return self.rows


class _LazyHost:
    """Attribute that lives on the device until somebody reads it (then cached on the host)."""

    def __init__(self, name):
        self.dev, self.host = "_" + name + "_dev", "_" + name + "_host"

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        host = obj.__dict__.get(self.host)
        if host is None:
            dev = obj.__dict__.get(self.dev)
            if dev is None:
                raise AttributeError(self.host[1:-5])
            if isinstance(dev, _Deferred):
                dev = obj.__dict__[self.dev] = dev.resolve()
            host = obj.__dict__[self.host] = Engine.get(dev.device).to_host(dev)
        return host

    def __set__(self, obj, value):
        if isinstance(value, _Deferred) or (isinstance(value, torch.Tensor) and value.is_cuda):
            obj.__dict__[self.dev] = value
            obj.__dict__.pop(self.host, None)
        else:
            obj.__dict__[self.host] = value
            obj.__dict__.pop(self.dev, None)

    def __delete__(self, obj):
        obj.__dict__.pop(self.dev, None)
        obj.__dict__.pop(self.host, None)


class Diagnostic2D(ABC):
    """Coordinate handling shared by every 2-D diagnostic (R:_base_classes.py:13-50)."""

    @abstractmethod
    def compute(self, x, y) -> None:
        ys = (y if type(y) is np.ndarray else np.array(y)).squeeze()
        xs = (x if type(x) is np.ndarray else np.array(x)).squeeze()
        if ys.ndim > 1:
            raise ValueError("y must be a 1-dimensional array.")
        if xs.ndim > 1:
            raise ValueError("x must be a 1-dimensional array.")
        self.y, self.x = ys, xs
        self.ydim, self.xdim = len(ys), len(xs)
        # float64 copies for the kernels (np.gradient casts integer coordinates the same way)
        self._yf, self._xf = as_coord(ys), as_coord(xs)

    @staticmethod
    def not_masked(*args) -> bool:
        """True when none of the values is ``np.ma.masked`` (R:_base_classes.py:38-50)."""
        return not any(a is np.ma.masked for a in args)


class EulerianDiagnostic2D(Diagnostic2D, ABC):
    """Velocity-field intake for the Eulerian diagnostics (R:_base_classes.py:53-113)."""

    u = _LazyHost("u")
    v = _LazyHost("v")

    @abstractmethod
    def compute(self, x, y, u, v, f=None, t=None, edge_order=1):
        """Validates the inputs and leaves ``(u, v)`` on the device.  Returns the device
        tensors; the gradients themselves are taken inside the fused K3 kernel."""
        super().compute(x, y)
        eng = Engine.get()
        if u is not None and v is not None:
            if isinstance(u, torch.Tensor) and isinstance(v, torch.Tensor):
                ud, vd = u.squeeze(), v.squeeze()
                if vd.ndim != 2:
                    raise ValueError("v must be a 2-dimensional array.")
                if ud.ndim != 2:
                    raise ValueError("u must be a 2-dimensional array.")
                ud = ud.to(eng.device, torch.float64).contiguous()
                vd = vd.to(eng.device, torch.float64).contiguous()
                self.u, self.v = ud, vd
            else:
                vh = (v if type(v) is np.ndarray else np.array(v)).squeeze()
                uh = (u if type(u) is np.ndarray else np.array(u)).squeeze()
                if vh.ndim != 2:
                    raise ValueError("v must be a 2-dimensional array.")
                if uh.ndim != 2:
                    raise ValueError("u must be a 2-dimensional array.")
                self.u, self.v = uh, vh
                for a in (uh, vh):
                    if tuple(a.shape) != (self.ydim, self.xdim):
                        raise ValueError("when 1d, distances must match the length of the "
                                         "corresponding dimension")
                rank, size = sharding.world()
                if size > 1:
                    # sharded: only this rank's rows (+ halo) cross PCIe, 1/N of the field per
                    # link instead of the whole field on every rank
                    slab = sharding.partition(self.ydim, size, rank)
                    ud = _HostSlab(eng.to_device(uh[slab.lo:slab.hi]), slab.lo)
                    vd = _HostSlab(eng.to_device(vh[slab.lo:slab.hi]), slab.lo)
                    if edge_order > 2:
                        raise ValueError("'edge_order' greater than 2 not supported")
                    return ud, vd
                ud, vd = eng.to_device(uh), eng.to_device(vh)
            for a in (ud, vd):
                if tuple(a.shape) != (self.ydim, self.xdim):
                    raise ValueError("when 1d, distances must match the length of the "
                                     "corresponding dimension")
        elif f is not None and t is not None:
            _, fid, dim, params = resolve_flow(f)
            if dim != 2:
                raise ValueError("a 2-D diagnostic needs a 2-D flow")
            # K0 is fused into the stencil kernel: (u, v) are never written unless the user reads
            # the attributes, in which case dlb_flow_eval_grid produces the same values
            xf, yf, tt = self._xf, self._yf, float(t)
            cache = {}
            self.u = _Deferred(lambda: eng.flow_eval_grid(fid, params, tt, xf, yf), 0, cache)
            self.v = _Deferred(lambda: eng.flow_eval_grid(fid, params, tt, xf, yf), 1, cache)
            ud, vd = _FlowSource(fid, params, tt), None
        else:
            raise RuntimeError(
                "Either a velocity field (u, v) or a function and timestep from which to "
                "calculate a velocity field must be passed to the compute function.")
        if edge_order > 2:
            raise ValueError("'edge_order' greater than 2 not supported")
        return ud, vd

    def _euler(self, mode, ud, vd, edge_order, negate=False, xi_mode=nat.XI_NONE):
        """Run K3 on the whole grid (row-sharded when a process group is active)."""
        eng = Engine.get()
        rank, size = sharding.world()
        fused = isinstance(ud, _FlowSource)
        if size == 1:
            if fused:
                return eng.euler_stencil_flow(mode, ud.fid, ud.params, ud.t, self._xf, self._yf,
                                              edge_order, 0, self.ydim, negate, xi_mode)
            return eng.euler_stencil(mode, ud, vd, 0, self._xf, self._yf, edge_order, 0,
                                     self.ydim, negate, xi_mode)
        slab = sharding.partition(self.ydim, size, rank)
        f = m = xi = None
        if slab.rows and fused:
            f, m, xi = eng.euler_stencil_flow(mode, ud.fid, ud.params, ud.t, self._xf, self._yf,
                                              edge_order, slab.row0, slab.rows, negate, xi_mode)
        elif slab.rows:
            f, m, xi = eng.euler_stencil(mode, ud[slab.lo:slab.hi], vd[slab.lo:slab.hi], slab.lo,
                                         self._xf, self._yf, edge_order, slab.row0, slab.rows,
                                         negate, xi_mode)

        def gather(part, tail, dtype):
            return sharding.gather_slab(part, slab, self.ydim, (self.xdim,) + tail, dtype, eng.device)

        need_mask = mode in (nat.EULER_RATE, nat.EULER_RATIO)
        return (gather(f, (), torch.float64),
                gather(m, (), torch.uint8) if need_mask else None,
                gather(xi, (2,), torch.float64) if xi_mode != nat.XI_NONE else None)


class LagrangianDiagnostic2D(Diagnostic2D, ABC):
    """Flow-map builder (R:_base_classes.py:116-189)."""

    flow_map = _LazyHost("flow_map")

    def __init__(self, integrator=odeint_wrapper, num_threads: int = 1) -> None:
        super().__init__()
        self.num_threads = num_threads
        self.integrator = integrator
        # The reference builds a process pool for num_threads > 1 (:139-146); on the GPU every
        # seed already has its own thread, so the value is validated and otherwise unused.
        if not num_threads >= 1:
            raise ValueError("num_threads must be an integer greater than 0.")
        self.stats = {}
        # Multi-GPU: all-gather the full flow map on every rank after the integration (what the
        # reference's `flow_map` attribute holds).  compute_device() of FTLE switches it off —
        # the stencil only needs the 1-row halo — and leaves `flow_map` as this rank's rows.
        self.gather_flow_map = True

    def _tolerances(self, kwargs):
        if self.integrator is odeint_wrapper:
            return parse_integrator_kwargs(kwargs, ODEINT_DEFAULT_TOL, ODEINT_DEFAULT_TOL)
        if self.integrator is rk45_wrapper:
            return parse_integrator_kwargs(kwargs, RK45_DEFAULT_RTOL, RK45_DEFAULT_ATOL)
        raise UnsupportedIntegratorError(
            f"integrator {self.integrator!r} cannot run on the device: use "
            f"dynlab_b200.utils.odeint_wrapper (default) or rk45_wrapper; there is no CPU "
            f"fallback for Python integrators")

    def _plan(self, x, y, f, t, kwargs) -> dict:
        """Validate the arguments exactly like the reference / scipy would and resolve the flow
        and the integrator options into what K1 needs."""
        Diagnostic2D.compute(self, x, y)
        if len(t) != 2:
            raise ValueError("t must only have 2 values, t_0 and t_final")
        kw = self._tolerances(kwargs)
        _, fid, dim, params = resolve_flow(f)
        if dim != 2:
            raise ValueError("a 2-D diagnostic needs a 2-D flow")
        atol = np.asarray(kw["atol"], dtype=np.float64)
        if atol.ndim > 0:
            if atol.shape != (2,):
                raise ValueError("`atol` has wrong shape.")
            if atol[0] != atol[1]:
                raise NotImplementedError("per-component atol is only available through "
                                          "dynlab_b200.utils.rk45_wrapper")
            atol = atol[0]
        if float(atol) < 0:
            raise ValueError("`atol` must be positive.")
        t0, tf = float(t[0]), float(t[1])
        first = kw["first_step"]
        if first is not None and first > abs(tf - t0):
            raise ValueError("`first_step` exceeds bounds.")
        if kw["rtol"] < 100 * np.finfo(float).eps:
            warnings.warn("At least one element of `rtol` is too small. "
                          f"Setting `rtol = np.maximum(rtol, {100 * np.finfo(float).eps})`.")
        return dict(fid=fid, params=params, t0=t0, tf=tf, rtol=kw["rtol"], atol=float(atol),
                    max_step=kw["max_step"], first_step=first or 0.0)

    def _integrate_rows(self, plan, row0, rows, out, counters=None, eng=None, y=None, nfev=None,
                        ws_slot=None, coords=None, stat_rows=None):
        """Enqueue K1 for rows [row0, row0 + rows) of the grid into ``out`` ([rows, xdim, 2])
        on ``eng``'s device (default: the current one)."""
        eng = eng or Engine.get()
        return eng.flow_map(plan["fid"], plan["params"], self._xf, self._yf if y is None else y,
                            row0, rows, plan["t0"], plan["tf"], plan["rtol"], plan["atol"],
                            plan["max_step"], plan["first_step"], out=out, counters=counters,
                            nfev=nfev, ws_slot=ws_slot, coords=coords, stat_rows=stat_rows)

    # -- cost-weighted row slabs ------------------------------------------------------------
    _row_cost_cache: dict = {}
    PILOT_ROWS = 64          # rows sampled by the pilot run
    PILOT_MIN_SEEDS = 1 << 20

    def _row_costs(self, plan, eng):
        """Per-row work estimate for ``sharding.weighted_bounds``, measured once per (flow, t,
        tolerances, grid) by a pilot run of K1 with per-seed ``nfev`` output over 1/64 of the
        seeds (~1 ms at config 3): in EVERY row, a few of its 32-seed groups (columns staggered
        from row to row).  A warp advances until its slowest lane is done, so a group costs its
        largest nfev; a row costs the mean over its sampled groups.  Sampling every row matters:
        the double gyre's cost per row rises by 5 % over the last 200 rows of config 3, which a
        every-64th-row pilot misses (it left the last of 8 slabs 1.9 % heavy; this one: 0.1 %).
        Removes the flow's own imbalance (4 % between equal slabs at 8 GPUs); returns None for
        grids too small to matter."""
        ydim, xdim = self.ydim, self.xdim
        if ydim * xdim < self.PILOT_MIN_SEEDS or ydim < 4 * self.PILOT_ROWS:
            return None
        key = (plan["fid"], tuple(plan["params"]), plan["t0"], plan["tf"], plan["rtol"],
               plan["atol"], plan["max_step"], plan["first_step"], self._xf.tobytes(),
               self._yf.tobytes())
        cost = self._row_cost_cache.get(key)
        if cost is None:
            groups = -(-xdim // 32)
            per_row = max(1, min(groups, round(groups / 64)))
            rows = np.arange(ydim)[:, None, None]
            g = (7 * rows + np.arange(per_row)[None, :, None] * (groups // per_row)) % groups
            cols = np.minimum(g * 32 + np.arange(32)[None, None, :], xdim - 1)
            seeds = np.stack([self._xf[cols], np.broadcast_to(self._yf[rows], cols.shape)], axis=-1)
            with torch.cuda.device(eng.device):
                sd = torch.from_numpy(np.ascontiguousarray(seeds.reshape(-1, 2))).to(eng.device)
                nfev = torch.zeros(sd.shape[0], dtype=torch.int32, device=eng.device)
                eng.rk45_endpoints(plan["fid"], plan["params"], sd, plan["t0"], plan["tf"],
                                   plan["rtol"], plan["atol"], plan["max_step"],
                                   plan["first_step"], nfev=nfev)
                cost = nfev.view(ydim, per_row, 32).max(dim=2).values.double().mean(dim=1)
                cost = cost.cpu().numpy()
            if len(self._row_cost_cache) > 32:
                self._row_cost_cache.clear()
            self._row_cost_cache[key] = cost
        return cost

    @abstractmethod
    def compute(self, x, y, f, t, **kwargs):
        """Integrate this rank's row slab.  Returns ``(local, slab)`` where ``local`` is the
        device tensor ``[slab.nloc, xdim, 2]`` with the own rows filled and the halo rows
        exchanged; attribute ``flow_map`` (host ``[ydim, xdim, 2]``) is materialised lazily."""
        plan = self._plan(x, y, f, t, kwargs)
        eng = Engine.get()
        rank, size = sharding.world()
        slab = sharding.partition(self.ydim, size, rank)
        local = eng.empty(max(slab.nloc, 1), self.xdim, 2)
        redundant = size > 1 and sharding.halo_mode() == "redundant"
        if slab.rows and redundant:
            # integrate the (at most two) halo rows here as well: 2 / rows extra seeds, no
            # neighbour exchange and no waiting for a slower neighbour before the stencil; the
            # per-seed integration is deterministic, so the rows equal the neighbours' bit for bit
            self._integrate_rows(plan, slab.lo, slab.nloc, local[:slab.nloc],
                                 stat_rows=(slab.row0, slab.rows))
        elif slab.rows:
            own = local[slab.halo_lo:slab.halo_lo + slab.rows]
            self._integrate_rows(plan, slab.row0, slab.rows, own)
        self._counters_pending = slab.rows > 0
        self._counter_tensors = None
        if not slab.rows:  # idle rank (more ranks than row pairs): nothing integrated here
            self.stats = {"nfev": 0, "accepted": 0, "rejected": 0, "failed": 0, "seeds": 0}
        own = local[slab.halo_lo:slab.halo_lo + slab.rows]
        if size > 1:
            if not redundant:
                sharding.exchange_halo(local, slab, self.ydim, rank, size)
            if self.gather_flow_map:
                if slab.rows == slab.rows_max:
                    pad = own  # contiguous row block: gather straight from it
                else:
                    pad = torch.empty(slab.rows_max, self.xdim, 2, dtype=torch.float64,
                                      device=eng.device)
                    pad[:slab.rows] = own
                self.flow_map = sharding.gather_rows(pad, self.ydim, size)
            else:
                self.flow_map = own
        else:
            self.flow_map = local[:slab.nloc]
        return local, slab

    def _collect_stats(self):
        """Read the integrator counters (after the results were synchronised) and warn about
        seeds scipy would have ended with status -1; the reference keeps their last state
        silently (R:_base_classes.py:180-182), so the values are kept."""
        if getattr(self, "_counters_pending", False):
            tensors = getattr(self, "_counter_tensors", None)
            if isinstance(tensors, list):  # one [chunks, 8] tensor per local group member
                parts = [Engine.get(t.device).read_counters(t) for t in tensors]
                self.stats = {k: sum(p[k] for p in parts) for k in parts[0]} if parts else {
                    "nfev": 0, "accepted": 0, "rejected": 0, "failed": 0, "seeds": 0}
            else:
                self.stats = Engine.get().read_counters(tensors)
            self._counters_pending = False
            if self.stats["failed"]:
                warnings.warn(f"{self.stats['failed']} trajectories stopped early: required step "
                              "size is less than spacing between numbers (scipy status -1)",
                              RuntimeWarning)


class RidgeExtractor2D(Diagnostic2D, ABC):
    """Ridge extraction shared by LCS and iLES (R:_base_classes.py:192-277)."""

    def extract_ridges(self, field, Xi, percentile, edge_order, debug=False):
        """``field`` / ``Xi``: device tensors ``[ydim, xdim]`` / ``[ydim, xdim, 2]``.
        Gradient / Hessian / masks, the percentile threshold and the contour cell pass run on
        the device; only the O(ridge length) segment records come back to be linked."""
        from .. import _contour

        eng = Engine.get()
        threshold = None
        if percentile:
            # R:_base_classes.py:266-268 — np.percentile of the (unmasked) field values, selected
            # on the device (csrc/select.cu) and interpolated like numpy
            threshold = eng.percentile(field, percentile)
        dd, conc, mask = eng.ridge_field(field, Xi, self._xf, self._yf, edge_order, threshold)
        # cell classification on the device, polyline linking in C++ (host, O(ridge length))
        ridge_lines = _contour.zero_contour_lines_device(eng, dd, mask, self._xf, self._yf)
        if debug:
            self.directional_derivative = np.ma.MaskedArray(
                eng.to_host(dd), mask=eng.to_host(mask).astype(bool))
            self.concavity = np.ma.MaskedArray(eng.to_host(conc))
        return ridge_lines
