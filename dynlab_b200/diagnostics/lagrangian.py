"""FTLE and LCS (mirror of R:src/dynlab/diagnostics/lagrangian.py).

Pipeline per ``compute``: K1 (flow map of this rank's row slab) -> halo rows -> K2 (np.gradient
stencil + Cauchy-Green tensor + eigen + log, fused) -> gather -> one device-to-host copy of the
field.  ``flow_map`` stays on the device until it is read.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as nat
from .. import peer, sharding
from .._engine import Engine
from ..utils import odeint_wrapper
from ._base_classes import LagrangianDiagnostic2D, RidgeExtractor2D, _Deferred


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
    out = xi = None
    if slab.rows:
        out, xi = eng.ftle_stencil(local, slab.lo, diag._xf, diag._yf, edge_order, t_span,
                                   slab.row0, slab.rows, xi_mode)
    field = sharding.gather_slab(out, slab, diag.ydim, (diag.xdim,), torch.float64, eng.device)
    if xi_mode != nat.XI_NONE:
        xi = sharding.gather_slab(xi, slab, diag.ydim, (diag.xdim, 2), torch.float64, eng.device)
    return field, xi


def pipeline_schedule(lo, hi, own0, own1, ydim, edge_order, chunks, last_fraction):
    """Chunk plan of ``FTLE._pipeline``: a list of ``(r0, r1, done, upto)`` - integrate global rows
    [r0, r1), then produce the FTLE rows [done, upto) (empty when ``upto == done``).  The chunks
    tile [lo, hi) in order (equal chunks and a short last one, whose copy is the only one left
    exposed; small slabs: one chunk), the output ranges tile [own0, own1) in order, and every
    row an output row's stencil reads - its neighbours, and rows 0..2 / ydim-3..ydim-1 for the
    one-sided edge_order-2 formulas at the global edges - lies in [lo, r1)."""
    nloc = hi - lo
    if isinstance(chunks, (tuple, list)):
        # explicit chunk sizes as fractions of the rows (overlapped pipeline: large chunks
        # first, the last two small); chunks below 32 rows are merged into their predecessor
        cum, bounds = 0.0, [lo]
        for fr in chunks[:-1]:
            cum += fr / float(sum(chunks))
            b = lo + int(round(nloc * cum))
            if b - bounds[-1] >= 32 and hi - b >= 32:
                bounds.append(b)
        bounds.append(hi)
    elif chunks > 1 and nloc >= 64 * chunks:
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


def member_schedule(slab, ydim, edge_order, exchange, chunks, last_fraction):
    """Chunk plan of one group member: ``(steps, tail)``.  ``steps`` as in
    ``pipeline_schedule``; ``tail`` lists the output row ranges that have to wait for the halo
    exchange (``exchange`` mode: K1 covers the own rows only, the rows next to a neighbour are
    produced after ``dlb_halo_exchange``).  Without exchange the halo rows are integrated
    locally and ``tail`` is empty."""
    s = slab
    if not s.rows:
        return [], []
    if exchange and (s.halo_lo or s.halo_hi):
        own0, own1 = s.row0 + s.halo_lo, s.row0 + s.rows - s.halo_hi
        if edge_order == 2 and own1 == ydim and ydim - 3 < s.row0:
            own1 = max(own0, ydim - 1)  # the one-sided formula of the last row reads the halo row
        if own1 > own0:
            steps = pipeline_schedule(s.row0, s.row0 + s.rows, own0, own1, ydim, edge_order,
                                      chunks, last_fraction)
        else:
            steps = [(s.row0, s.row0 + s.rows, own0, own0)]
        done = steps[-1][3]
        tail = [r for r in ((s.row0, own0), (done, s.row0 + s.rows)) if r[1] > r[0]]
        return steps, tail
    return pipeline_schedule(s.lo, s.hi, s.row0, s.row0 + s.rows, ydim, edge_order, chunks,
                             last_fraction), []


def counted_rows(r0, r1, slab):
    """(first row, number of rows) of the K1 chunk [r0, r1) whose seeds enter the integrator
    counters: the chunk's intersection with the member's OWN rows (redundantly integrated halo
    rows are computed but not counted, so a sharded run reports the single-device totals)."""
    a, b = max(r0, slab.row0), min(r1, slab.row0 + slab.rows)
    return (a, b - a) if b > a else (r0, 0)


class FTLE(LagrangianDiagnostic2D):
    """Finite-time Lyapunov exponent field of a 2-D flow (R:lagrangian.py:9-78)."""

    def __init__(self, integrator=odeint_wrapper, num_threads: int = 1) -> None:
        super().__init__(integrator=integrator, num_threads=num_threads)

    def compute(self, x, y, f, t, edge_order: int = 1, **kwargs):
        """Returns ``np.ma.MaskedArray[ydim, xdim]``; sets ``x, y, xdim, ydim, flow_map,
        field`` like the reference.  ``kwargs`` (rtol, atol, max_step, first_step) go to the
        integrator."""
        group = self._group(np.size(x) * np.size(y))
        if group is not None:  # several GPUs: row slabs, peer-memory data path (csrc/peer.cu)
            return self._compute_group(group, self._plan(x, y, f, t, kwargs), t, edge_order, "host")
        n_chunks = len(self.pipeline_chunks) if isinstance(self.pipeline_chunks, (tuple, list)) \
            else self.pipeline_chunks
        if (sharding.world()[1] == 1 and n_chunks > 1
                and np.size(y) >= 64 * n_chunks and np.size(x) * np.size(y) >= (1 << 22)):
            return self._compute_pipelined(self._plan(x, y, f, t, kwargs), t, edge_order)
        # one device, or sharding.set_transport("nccl"): kernels, then NCCL halo + all-gathers,
        # then one device-to-host copy of the gathered field per rank
        field = self._run(x, y, f, t, edge_order, self.gather_flow_map, kwargs)
        many = sharding.world()[1] > 1
        self.field = np.ma.MaskedArray(Engine.get().to_host(field, pinned=True if many else None))
        self._collect_stats()
        return self.field

    @staticmethod
    def _group(n_seeds):
        if sharding.world()[1] > 1 and sharding.transport() != "peer":
            return None
        return peer.current_group(n_seeds)

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

    # ---------------------------------------------------------------------------------------
    # Several GPUs (one process driving all of them, or one torchrun rank per GPU): every
    # member integrates its row slab in chunks; finished FTLE rows leave the device on the copy
    # stream while the next chunk integrates — into pinned host memory ("host": what compute()
    # returns) or, pushed by a copy engine over NVLink, into every member's copy of the field
    # ("device": what compute_device() returns).  Nothing is gathered afterwards and no
    # collective library is on the data path; `flow_map` is collected from the members' buffers
    # only if it is read.  Results are bit-identical to the single-device ones: seeds are
    # independent and a row's stencil reads its neighbour rows only.
    def _compute_group(self, G, plan, t, edge_order, target):
        if edge_order > 2:
            raise ValueError("'edge_order' greater than 2 not supported")
        t_span = abs(float(t[-1]) - float(t[0]))
        if t_span == 0.0:
            raise ZeroDivisionError("float division by zero")  # R:lagrangian.py:71
        ydim, xdim = self.ydim, self.xdim
        G.check_timeouts()
        bounds = self._slab_bounds(G, plan)
        slabs = [sharding.slab_from_bounds(bounds, ydim, r) for r in range(G.size)]
        exchange = sharding.halo_mode() == "exchange"
        gen = G.next_generation()
        row_b = xdim * 16
        # rotating buffers are all created by the first call that needs them (allocation, IPC
        # mapping and page-locking cost ~0.1 s per buffer: never inside a later, timed call)
        fm_ring = [G.symmetric(("flow_map", k), max(max(s.nloc for s in slabs), 1) * row_b)
                   for k in range(2)]
        fm_sym = fm_ring[gen % 2]
        field_sym = host_np = host_t = None
        if target == "device":
            ring = [G.symmetric(("field", k), ydim * xdim * 8) for k in range(peer.RESULT_DEPTH)]
            field_sym = ring[gen % peer.RESULT_DEPTH]
        elif G.multiprocess:
            # a shared host segment no rank still holds a result in (SharedResultPool)
            host_np, host_t = sharding.SharedResultPool.get((ydim, xdim)).acquire()
        else:
            host_t = torch.empty((ydim, xdim), dtype=torch.float64, pin_memory=True)
            host_np = host_t.numpy()

        # one process per GPU: every rank ends up with the whole field; one process driving all
        # of them: only the member whose tensor is returned needs it
        dests = None if G.multiprocess else [G.local[0].rank]
        runs = []
        for m in G.local:
            s = slabs[m.rank]
            eng = m.engine
            with torch.cuda.device(m.device):
                main = torch.cuda.current_stream(m.device)
                main.wait_stream(eng.copy_stream)  # pushes of earlier calls out of reused buffers
                steps, tail = member_schedule(
                    s, ydim, edge_order, exchange,
                    self.device_chunks if target == "device" else self.pipeline_chunks,
                    self.pipeline_last_fraction)
                fm = fm_sym.view(m, torch.float64, (max(s.nloc, 1), xdim, 2))
                if target == "device":
                    out = field_sym.view(m, torch.float64, (ydim, xdim))  # global row index
                    out_row0 = 0
                else:
                    out = eng.empty(max(s.rows, 1), xdim)
                    out_row0 = s.row0
                counters = torch.zeros(max(len(steps), 1), nat.CNT_WORDS, dtype=torch.int64,
                                       device=m.device)
            runs.append(dict(m=m, s=s, steps=steps, tail=tail, fm=fm, out=out, out_row0=out_row0,
                             counters=counters, main=main, pushes=0,
                             coords=(eng.device_coord(self._xf), eng.device_coord(self._yf))))

        trace = self.trace if isinstance(self.trace, list) else None

        def mark(run, label, stream):
            """Debug timeline (``FTLE.trace = []``): a timed CUDA event on ``stream``."""
            if trace is not None:
                with torch.cuda.device(run["m"].device):
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record(stream)
                trace.append((run["m"].rank, label, ev))

        for run in runs:
            mark(run, "start", run["main"])

        def ship(run, a, b, last, after=None):
            """FTLE rows [a, b) of this member were just enqueued on stream ``after`` (default:
            its main stream): send them on their way on the copy stream."""
            m, eng = run["m"], run["m"].engine
            with torch.cuda.device(m.device):
                ev = torch.cuda.Event()
                ev.record(after or run["main"])
                eng.copy_stream.wait_event(ev)
                if target == "device":
                    lanes = int(self.push_lanes)
                    everyone = list(range(G.size)) if dests is None else list(dests)
                    if lanes > 1 and not last and len(everyone) > 2:
                        # peers dealt over several copy streams: one engine does not fill the
                        # NVLink ports of seven receivers
                        for k, st in enumerate(eng.get_push_streams(lanes)):
                            st.wait_event(ev)
                            G.push_rows(m, field_sym, a * xdim * 8, (b - a) * xdim * 8, None,
                                        st.cuda_stream, everyone[k::lanes])
                            eng.copy_stream.wait_stream(st)
                    else:
                        G.push_rows(m, field_sym, a * xdim * 8, (b - a) * xdim * 8,
                                    gen * peer.GEN_STRIDE + peer.DONE if last else None,
                                    eng.copy_stream.cuda_stream, dests)
                    run["pushed"] = torch.cuda.Event()
                    run["pushed"].record(eng.copy_stream)
                    mark(run, f"dma-push rows {a}:{b} done", eng.copy_stream)
                elif b > a:
                    with torch.cuda.stream(eng.copy_stream):
                        r0 = run["out_row0"]
                        host_t[a:b].copy_(run["out"][a - r0:b - r0], non_blocking=True)
                    mark(run, f"d2h rows {a}:{b} done", eng.copy_stream)

        def stencil(run, g0, g1, a, b):
            """K2: FTLE rows [a, b) from the member's flow-map rows [g0, g1) (global indices)."""
            eng, r0, lo = run["m"].engine, run["out_row0"], run["s"].lo
            with torch.cuda.device(run["m"].device):
                eng.ftle_stencil(run["fm"][g0 - lo:g1 - lo], g0, self._xf, self._yf, edge_order,
                                 t_span, a, b - a, out=run["out"][a - r0:b - r0])

        def produce(run, g0, g1, a, b, last, fused, stream):
            """K2 for FTLE rows [a, b) on ``stream`` and their departure.  ``fused`` (device
            target, the last rows of a slab): dlb_ftle_stencil_gather - the epilogue itself
            stores into the other members' fields over NVLink and the flags follow the kernel,
            no copy-engine pass after it."""
            m, eng = run["m"], run["m"].engine
            if not (fused and target == "device"):
                with torch.cuda.device(m.device), torch.cuda.stream(stream):
                    stencil(run, g0, g1, a, b)
                mark(run, f"K2 rows {a}:{b} done", stream)
                ship(run, a, b, last, stream)
                return
            peers, flags = G.push_targets(m, field_sym, a * xdim * 8, dests)
            value = gen * peer.GEN_STRIDE + peer.DONE
            if not last:
                flags = []
            lo = run["s"].lo
            with torch.cuda.device(m.device), torch.cuda.stream(stream):
                if run.get("pushed") is not None:  # the flag must not overtake earlier DMA pushes
                    stream.wait_event(run["pushed"])
                eng.ftle_stencil_gather(run["fm"][g0 - lo:g1 - lo], g0, self._xf, self._yf,
                                        edge_order, t_span, a, b - a, run["out"][a:b], peers,
                                        flags, value)
            mark(run, f"K2+gather rows {a}:{b} done", stream)

        # chunk-major over the local members, so that every device has work queued before the
        # host moves on to the next chunk
        overlap = bool(self.overlap_chunks)
        side = bool(self.side_stencil) and not overlap
        for run in runs:
            eng = run["m"].engine
            run["k1_streams"] = [run["main"], eng.alt_stream] if overlap else [run["main"]]
            run["k2_stream"] = eng.stencil_stream if overlap else run["main"]
            run["k1_events"] = []
            if overlap:
                eng.alt_stream.wait_stream(run["main"])
            if overlap or side:
                eng.stencil_stream.wait_stream(run["main"])
        for c in range(max([len(r["steps"]) for r in runs] + [0])):
            for run in runs:
                if c >= len(run["steps"]):
                    continue
                m, s = run["m"], run["s"]
                r0, r1, done, upto = run["steps"][c]
                lo = s.lo  # fm row 0 is global row s.lo (halo slot included)
                k1s = run["k1_streams"][c % len(run["k1_streams"])]
                with torch.cuda.device(m.device), torch.cuda.stream(k1s):
                    # overlap: K1 of chunk c+1 (other stream, private workspace) takes over the
                    # SMs CTA by CTA as chunk c drains - no tail bubble between chunks
                    self._integrate_rows(plan, r0, r1 - r0, run["fm"][r0 - lo:r1 - lo],
                                         counters=run["counters"][c], eng=m.engine,
                                         ws_slot=(c % 2) if overlap else None,
                                         coords=run["coords"],
                                         stat_rows=counted_rows(r0, r1, s))
                    if overlap:
                        ev = torch.cuda.Event()
                        ev.record(k1s)
                        run["k1_events"].append(ev)
                mark(run, f"K1 rows {r0}:{r1} done", k1s)
                if upto > done:
                    k2s = run["k2_stream"]
                    left = len(run["steps"]) - 1 - c  # chunks still to come
                    if overlap:
                        for ev in run["k1_events"][-2:]:
                            k2s.wait_event(ev)
                    elif side and left > 0:
                        # K2 of a chunk that is not the last: on the priority stencil stream,
                        # enqueued before the next K1 - both become ready when this K1 ends, K2's
                        # CTAs are picked first and the next K1 fills the SMs around them, so K2
                        # costs the main stream nothing
                        with torch.cuda.device(m.device):
                            ev = torch.cuda.Event()
                            ev.record(run["main"])
                            m.engine.stencil_stream.wait_event(ev)
                        k2s = m.engine.stencil_stream
                    last = left == 0 and not run["tail"]
                    produce(run, run["steps"][0][0], r1, done, upto, last,
                            left < self.fused_gather_chunks, k2s)
        if overlap or side:
            for run in runs:
                eng = run["m"].engine
                if overlap:
                    run["main"].wait_stream(eng.alt_stream)
                run["main"].wait_stream(eng.stencil_stream)
        # 1-row halo exchange (dlb_halo_exchange: every member pushes first, then waits), then the
        # rows that were waiting for it
        for run in runs:
            if run["tail"]:
                G.halo_push(run["m"], fm_sym, row_b, slabs, gen, run["main"].cuda_stream)
        for run in runs:
            m, s = run["m"], run["s"]
            if run["tail"]:
                mark(run, "halo pushed", run["main"])
                if not G.multiprocess:  # see wait_gathered below: stream dependency, no spinning
                    for other in runs:
                        if abs(other["m"].rank - m.rank) == 1:
                            run["main"].wait_stream(other["main"])
                G.halo_wait(m, slabs, gen, run["main"].cuda_stream)
                mark(run, "halo arrived", run["main"])
                for k, (a, b) in enumerate(run["tail"]):
                    produce(run, s.lo, s.hi, a, b, k == len(run["tail"]) - 1, True, run["main"])
            elif not s.rows:
                ship(run, 0, 0, True)  # nothing to send: still tell the others we are done

        self._counters_pending = True
        self._counter_tensors = [r["counters"] for r in runs if r["s"].rows]
        first = runs[0]
        cache = {}
        self.flow_map = _Deferred(lambda: self._gather_flow_map(G, fm_sym, slabs, first["m"]), 0,
                                  cache)
        if target == "device":
            for run in runs:
                if dests is None or run["m"].rank in dests:
                    if not G.multiprocess:
                        # one process: every member's departures are also stream dependencies
                        # (cross-device events), so the flag wait below never has to spin
                        for other in runs:
                            run["main"].wait_stream(other["m"].engine.copy_stream)
                            run["main"].wait_stream(other["main"])
                    G.wait_gathered(run["m"], gen, run["main"].cuda_stream)
                    mark(run, "all rows gathered", run["main"])
            self.field_device = first["out"]
            return first["out"]
        for run in runs:
            run["m"].engine.copy_stream.synchronize()
            if run["s"].rows:
                run["out"].record_stream(run["m"].engine.copy_stream)
        G.check_timeouts()
        if G.multiprocess:
            sharding.host_barrier()  # every rank's rows are in the segment
            if sharding.host_result_mode() == "private":
                host_np = host_np.copy()
        self.field_device = first["out"]
        self.field = np.ma.MaskedArray(host_np)
        self._collect_stats()
        return self.field

    def _slab_bounds(self, G, plan):
        """Row boundaries of the members' slabs: equal work (pilot run, `_row_costs`) when the
        grid is large enough to care, equal rows otherwise.  With one process per GPU rank 0's
        estimate is broadcast once per (flow, t, grid) so that all ranks cut at the same rows."""
        key = ("bounds", G.size, bool(self.weighted_slabs), plan["fid"], tuple(plan["params"]),
               plan["t0"], plan["tf"], plan["rtol"], plan["atol"], self._xf.tobytes(),
               self._yf.tobytes())
        # its own dict: under torchrun every rank must hit or miss together (a miss is a
        # collective), and only rank 0 also stores pilot results in _row_cost_cache
        b = self._bounds_cache.get(key)
        if b is None:
            b = sharding.uniform_bounds(self.ydim, G.size)
            if self.weighted_slabs and G.size > 1:
                if G.multiprocess:
                    import torch.distributed as dist

                    box = [None]
                    if G.local[0].rank == 0:
                        cost = self._row_costs(plan, G.local[0].engine)
                        box[0] = None if cost is None else sharding.weighted_bounds(self.ydim, G.size, cost)
                    dist.broadcast_object_list(box, src=0)
                    b = box[0] or b
                else:
                    cost = self._row_costs(plan, G.local[0].engine)
                    if cost is not None:
                        b = sharding.weighted_bounds(self.ydim, G.size, cost)
            if len(self._bounds_cache) > 32:  # keys carry the coordinate bytes: keep it small
                self._bounds_cache.clear()
            self._bounds_cache[key] = b
        return b

    _bounds_cache: dict = {}

    weighted_slabs = True  # cut the slabs by integrator work instead of by rows (large grids)
    # K1 chunks on alternating streams (chunk c+1 fills the SMs as chunk c drains), K2 on its own
    # high-priority stream; False: everything in order on one stream (K2 between the chunks)
    overlap_chunks = False
    # K2 of every chunk but the last on the priority stencil stream, enqueued before the next K1.
    # Measured WORSE at 8 GPUs (9.82 vs 9.05 ms): the next K1's persistent CTAs are dispatched
    # first, K2 gets no SM slot until that chunk drains, and the copy of the first chunk's rows
    # starts a chunk late.  Kept as a switch for the record.
    side_stencil = False
    trace = None  # set to [] to collect (rank, label, timed CUDA event) of the next group call
    # device result: the K2 launches of this many last chunks (and of the rows that wait for a
    # halo exchange) store straight into the other GPUs' fields (dlb_ftle_stencil_gather);
    # earlier chunks are pushed by a copy engine behind the integrator
    fused_gather_chunks = 1
    # copy streams the peer pushes of a chunk are dealt over (device result): one per receiver at
    # 8 GPUs - a single stream's back-to-back copies finished just after the last K1 chunk and
    # held up the final fused K2 (8.784 -> 8.740 ms, profiles/r02_variants_n8_push_lanes.jsonl)
    push_lanes = 7
    # device result on several GPUs: chunk sizes (fractions of a slab's rows).  One large chunk
    # whose rows a copy engine spreads while the small last one integrates; the last one's K2 is
    # the fused gather.  (The host result keeps `pipeline_chunks`: equal chunks, because there
    # the box-wide device-to-host bandwidth is the limit.)
    device_chunks = (0.875, 0.125)

    def _gather_flow_map(self, G, fm_sym, slabs, m):
        """The reference's `flow_map` attribute (R:_base_classes.py:187): collected on demand by
        one-sided reads of the members' buffers (`dlb_peer_copy`), valid until the next-but-one
        compute of this group."""
        ydim, xdim = self.ydim, self.xdim
        with torch.cuda.device(m.device):
            full = m.engine.empty(ydim, xdim, 2)
            stream = torch.cuda.current_stream(m.device).cuda_stream
            for r, s in enumerate(slabs):
                if s.rows:
                    G.pull(m, full[s.row0:s.row0 + s.rows], fm_sym, r, s.halo_lo * xdim * 16, stream)
        return (full,)

    def compute_device(self, x, y, f, t, edge_order: int = 1, **kwargs):
        """Same computation, result left on the device (``torch.Tensor[ydim, xdim]``, also kept
        as ``field_device``); nothing is synchronised or copied to the host.  On several GPUs
        every member ends up with the whole field (the tensor returned is the first local
        member's); it lives in one of ``peer.RESULT_DEPTH`` rotating buffers, i.e. it keeps its
        values through the next call and is overwritten by the one after."""
        group = self._group(np.size(x) * np.size(y))
        if group is not None:
            return self._compute_group(group, self._plan(x, y, f, t, kwargs), t, edge_order, "device")
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
        self.gather_flow_map = False  # deleted right after the stencil (R:lagrangian.py:137-139)

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
