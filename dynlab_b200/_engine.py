"""Host side of the C-ABI: device buffers (owned by torch), streams, transfers.

PyTorch is plumbing here — it allocates device/pinned memory and provides the CUDA stream;
every floating-point operation of the path happens inside ``libdynlab_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _native as nat

_F64 = torch.float64


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "dynlab_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def as_coord(a) -> np.ndarray:
    """Coordinate vector as contiguous float64 (numpy casts integer coordinates to float64
    in np.gradient: numpy/lib/_function_base_impl.py:1242-1245)."""
    return np.ascontiguousarray(a, dtype=np.float64)


class Engine:
    """Per-device launcher.  Buffers that do not leave the engine (workspace, counters) are
    cached; result tensors are allocated per call through torch's caching allocator."""

    _by_device: dict[int, "Engine"] = {}

    @classmethod
    def get(cls, device=None) -> "Engine":
        _require_cuda()
        idx = torch.cuda.current_device() if device is None else torch.device(device).index
        if idx is None:
            idx = torch.cuda.current_device()
        eng = cls._by_device.get(idx)
        if eng is None:
            eng = cls._by_device[idx] = Engine(idx)
        return eng

    def __init__(self, index: int):
        self.device = torch.device("cuda", index)
        self._workspace = None
        self.counters = torch.zeros(nat.CNT_WORDS, dtype=torch.int64, device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)  # D2H overlapped with compute
        # chunk pipeline of the Lagrangian path: K1 chunks alternate between the caller's stream
        # and `alt_stream` (chunk c+1 fills the SMs as chunk c drains), K2 of a finished chunk
        # runs on the high-priority `stencil_stream` as soon as CTA slots free up
        self.push_streams = []  # extra copy streams: peer pushes to several GPUs side by side
        self.alt_stream = torch.cuda.Stream(device=self.device)
        self.stencil_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self._k1_ws = {}      # slot -> private K1 workspace (work counter + x, y rows)
        self._coords = {}     # coordinate bytes -> device-resident copy
        self.profile = None  # set to {} to collect (start, end) CUDA events per kernel name
        self.launches = 0    # kernels of libdynlab_b200.so launched through this engine

    def _timed(self, name, n_kernels=1):
        """Context manager: count launches; record CUDA events around them when profiling."""
        eng = self

        class _T:
            def __enter__(self_t):
                if eng.profile is not None:
                    self_t.e0 = torch.cuda.Event(enable_timing=True)
                    self_t.e1 = torch.cuda.Event(enable_timing=True)
                    self_t.e0.record(torch.cuda.current_stream(eng.device))

            def __exit__(self_t, *exc):
                eng.launches += n_kernels
                if eng.profile is not None:
                    self_t.e1.record(torch.cuda.current_stream(eng.device))
                    eng.profile.setdefault(name, []).append((self_t.e0, self_t.e1))
                return False

        return _T()

    # -- plumbing ---------------------------------------------------------------------
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def workspace(self, xdim: int, ydim: int):
        need = int(nat.lib.dlb_workspace_bytes(xdim, ydim))
        if self._workspace is None or self._workspace.numel() < need:
            # allocate the new block while the old one is still alive, so that the address
            # changes (the library caches stencil coefficients per workspace address)
            grown = torch.empty(max(need, 1 << 22), dtype=torch.uint8, device=self.device)
            self._workspace = grown
        return self._workspace.data_ptr(), self._workspace.numel()

    def get_push_streams(self, n: int):
        while len(self.push_streams) < n:
            self.push_streams.append(torch.cuda.Stream(device=self.device))
        return self.push_streams[:n]

    def k1_workspace(self, slot: int, xdim: int, rows: int):
        """Private workspace of K1 launches that may overlap other launches (chunk pipeline)."""
        need = int(nat.lib.dlb_workspace_bytes(xdim, rows))
        ws = self._k1_ws.get(slot)
        if ws is None or ws.numel() < need:
            ws = self._k1_ws[slot] = torch.empty(max(need, 1 << 20), dtype=torch.uint8,
                                                 device=self.device)
        return ws.data_ptr(), ws.numel()

    def device_coord(self, a: np.ndarray) -> torch.Tensor:
        """Device-resident copy of a coordinate vector, cached by value: K1 launches read x and
        y from it (``dlb_flowmap_rk45_resident``) instead of uploading them every time."""
        key = a.tobytes()
        hit = self._coords.pop(key, None)
        if hit is None:
            if len(self._coords) >= 64:  # drop the least recently used (long idle by then)
                self._coords.pop(next(iter(self._coords)))
            with torch.cuda.device(self.device):
                hit = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)
        self._coords[key] = hit  # most recently used last
        return hit

    def empty(self, *shape, dtype=_F64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    H2D_CHUNK = 16 << 20   # bytes per staged chunk
    H2D_WORKERS = 6

    def to_device(self, host: np.ndarray):
        """Host float64 array -> device tensor.  Large arrays (the 537 MB velocity fields of
        BASELINE config 2 arrive as pageable numpy memory) go through ``_to_device_staged``;
        a plain ``tensor.to(device)`` from pageable memory ran at ~5 GB/s and was 90 % of
        ``AttractionRate.compute(x, y, u=u, v=v)``."""
        a = np.ascontiguousarray(host, dtype=np.float64)
        if a.nbytes < 4 * self.H2D_CHUNK:
            return torch.from_numpy(a).to(self.device, non_blocking=False)
        return self._to_device_staged(a)

    def _to_device_staged(self, a: np.ndarray):
        """Pageable host -> device through pinned staging buffers: worker threads copy chunks
        into their own double-buffered pinned blocks (numpy releases the GIL for the memcpy) and
        enqueue the DMA on their own stream, so host memcpy bandwidth adds up over the workers
        and overlaps the PCIe transfers (537 MB: 16 ms = 33 GB/s with 6 workers; 48 ms plain)."""
        flat = a.reshape(-1)
        n = flat.size
        per = self.H2D_CHUNK // 8
        nchunks = (n + per - 1) // per
        workers = min(self.H2D_WORKERS, nchunks)
        with torch.cuda.device(self.device):
            dst = torch.empty(n, dtype=_F64, device=self.device)
            main = torch.cuda.current_stream(self.device)
            st_all = self._staging()

            def work(w):
                st = st_all[w]
                st["stream"].wait_stream(main)  # dst's memory may still be in use upstream
                for k, c in enumerate(range(w, nchunks, workers)):
                    buf, ev = st["bufs"][k & 1], st["evs"][k & 1]
                    ev.synchronize()  # the previous DMA out of this buffer is done
                    lo, hi = c * per, min(n, (c + 1) * per)
                    np.copyto(buf.numpy()[:hi - lo], flat[lo:hi])
                    with torch.cuda.stream(st["stream"]):
                        dst[lo:hi].copy_(buf[:hi - lo], non_blocking=True)
                        ev.record(st["stream"])

            self._run_workers(work, workers)
            for w in range(workers):
                dst.record_stream(st_all[w]["stream"])
                main.wait_stream(st_all[w]["stream"])
        return dst.view(a.shape)

    # Large results: False = staged copy into ordinary (pageable) memory, the same cost on every
    # call; True = one DMA into a result-sized pinned buffer from torch's caching host
    # allocator - 2.5x faster once the cache is warm (9.5 vs 24.5 ms per 537 MB), but
    # page-locking that buffer costs 200-300 ms on each of the first two calls of a given size
    # and the pinned blocks stay allocated.  Loops over many fields may prefer True.
    PINNED_RESULTS = False

    def to_host(self, t: torch.Tensor, pinned=None) -> np.ndarray:
        """Device tensor -> numpy array (see ``PINNED_RESULTS`` for large ones; ``pinned``
        overrides it for one call)."""
        if pinned is None:
            pinned = self.PINNED_RESULTS
        if (not pinned and t.dtype == _F64
                and t.numel() * t.element_size() >= 4 * self.H2D_CHUNK):
            return self._to_host_staged(t)
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return h.numpy()

    def _staging(self):
        """Per-worker stream + two pinned chunk buffers + their events (rebuilt when the chunk
        size or the worker count changes)."""
        per = self.H2D_CHUNK // 8
        cur = getattr(self, "_h2d", None)
        if cur is None or len(cur) < self.H2D_WORKERS or cur[0]["bufs"][0].numel() != per:
            self._h2d = [dict(stream=torch.cuda.Stream(device=self.device),
                              bufs=[torch.empty(per, dtype=_F64, pin_memory=True) for _ in range(2)],
                              evs=[torch.cuda.Event(), torch.cuda.Event()])
                         for _ in range(self.H2D_WORKERS)]
        return self._h2d

    def _run_workers(self, work, workers):
        import threading

        errors = []

        def guarded(w):
            try:
                torch.cuda.set_device(self.device)
                work(w)
            except Exception as exc:  # noqa: BLE001 - re-raised on the caller's thread
                errors.append(exc)

        threads = [threading.Thread(target=guarded, args=(w,)) for w in range(workers)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]

    def _to_host_staged(self, t: torch.Tensor) -> np.ndarray:
        """Device -> pageable host through the pinned staging buffers: every worker keeps one
        DMA in flight while it copies the previous chunk out of its other buffer."""
        src = t.contiguous().view(-1)
        n = src.numel()
        out = np.empty(n, dtype=np.float64)
        per = self.H2D_CHUNK // 8
        nchunks = (n + per - 1) // per
        workers = min(self.H2D_WORKERS, nchunks)
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            st_all = self._staging()

            def work(w):
                st = st_all[w]
                st["stream"].wait_stream(main)  # the producer kernels are upstream on `main`
                mine = list(range(w, nchunks, workers))

                def issue(k):
                    lo, hi = mine[k] * per, min(n, (mine[k] + 1) * per)
                    with torch.cuda.stream(st["stream"]):
                        st["bufs"][k & 1][:hi - lo].copy_(src[lo:hi], non_blocking=True)
                        st["evs"][k & 1].record(st["stream"])

                issue(0)
                for k in range(len(mine)):
                    if k + 1 < len(mine):
                        issue(k + 1)
                    st["evs"][k & 1].synchronize()
                    lo, hi = mine[k] * per, min(n, (mine[k] + 1) * per)
                    np.copyto(out[lo:hi], st["bufs"][k & 1].numpy()[:hi - lo])

            self._run_workers(work, workers)
            for w in range(workers):
                src.record_stream(st_all[w]["stream"])
        return out.reshape(tuple(t.shape))

    def read_counters(self, tensors=None) -> dict:
        """Integrator counters of the last K1 launch, or summed over ``tensors`` ([n, 8])."""
        c = self.counters.cpu().numpy() if tensors is None else tensors.cpu().numpy().sum(axis=0)
        return {"nfev": int(c[nat.CNT_NFEV]), "accepted": int(c[nat.CNT_ACCEPTED]),
                "rejected": int(c[nat.CNT_REJECTED]), "failed": int(c[nat.CNT_FAILED]),
                "seeds": int(c[nat.CNT_SEEDS])}

    # -- K1 ---------------------------------------------------------------------------
    def flow_map(self, flow_id, params, x, y, row0, rows, t0, tf, rtol, atol,
                 max_step=math.inf, first_step=0.0, out=None, nfev=None, status=None,
                 counters=None, ws_slot=None, coords=None, stat_rows=None):
        """Enqueue K1 for rows [row0, row0+rows) of the grid; returns the [rows, xdim, 2] tensor.
        ``ws_slot``: use that private K1 workspace instead of the engine's shared one (launches
        that run concurrently with other launches of this engine).  ``stat_rows = (row0, rows)``:
        only those rows enter the counters (default: all integrated rows)."""
        xdim = len(x)
        with torch.cuda.device(self.device):
            if out is None:
                out = self.empty(rows, xdim, 2)
            if ws_slot is None:
                ws, wsz = self.workspace(xdim, max(len(y), rows))
            else:
                ws, wsz = self.k1_workspace(ws_slot, xdim, rows)
            p = nat.darray(params)
            # coords: (device x, device y) looked up once by a caller that launches many chunks
            dx, dy = coords if coords is not None else (self.device_coord(x), self.device_coord(y))
            with self._timed("rk45_flowmap"):
              nat.check(nat.lib.dlb_flowmap_rk45_resident(
                flow_id, p, len(params), dx.data_ptr(), xdim, dy.data_ptr(), row0, rows,
                float(t0), float(tf), float(rtol), float(atol), float(max_step), float(first_step),
                out.data_ptr(), nfev.data_ptr() if nfev is not None else None,
                status.data_ptr() if status is not None else None,
                (self.counters if counters is None else counters).data_ptr(),
                row0 if stat_rows is None else int(stat_rows[0]),
                rows if stat_rows is None else int(stat_rows[1]), ws, wsz, self.stream()))
        return out

    def rk45_endpoints(self, flow_id, params, seeds, t0, tf, rtol, atol, max_step=math.inf,
                       first_step=0.0, nfev=None, status=None):
        m, n = seeds.shape
        with torch.cuda.device(self.device):
            out = self.empty(m, n)
            ws, wsz = self.workspace(0, 0)
            p = nat.darray(params)
            a = nat.darray(np.broadcast_to(np.asarray(atol, dtype=np.float64), (n,)))
            with self._timed("rk45_endpoints"):
              nat.check(nat.lib.dlb_rk45_endpoints(
                flow_id, p, len(params), n, seeds.data_ptr(), m, float(t0), float(tf),
                float(rtol), a, float(max_step), float(first_step), out.data_ptr(),
                nfev.data_ptr() if nfev is not None else None,
                status.data_ptr() if status is not None else None,
                self.counters.data_ptr(), ws, wsz, self.stream()))
        return out

    # -- K2 ---------------------------------------------------------------------------
    def ftle_stencil(self, fm_local, grow0, x, y, edge_order, t_span, out_row0, out_rows,
                     xi_mode=nat.XI_NONE, out=None, xi_out=None):
        xdim, ydim = len(x), len(y)
        nloc = fm_local.shape[0]
        with torch.cuda.device(self.device):
            if out is None:
                out = self.empty(out_rows, xdim)
            xi = None
            if xi_mode != nat.XI_NONE:
                xi = xi_out if xi_out is not None else self.empty(out_rows, xdim, 2)
            ws, wsz = self.workspace(xdim, ydim)
            with self._timed("ftle_stencil", 2):
              nat.check(nat.lib.dlb_ftle_stencil(
                fm_local.data_ptr(), nloc, grow0, nat.dptr(x), xdim, nat.dptr(y), ydim,
                int(edge_order), float(t_span), out_row0, out_rows, out.data_ptr(),
                xi.data_ptr() if xi is not None else None, xi_mode, ws, wsz, self.stream()))
        return out, xi

    def ftle_stencil_gather(self, fm_local, grow0, x, y, edge_order, t_span, out_row0, out_rows,
                            out, peer_ptrs, flag_ptrs, flag_value):
        """K2 fused with the row gather (``dlb_ftle_stencil_gather``): values go to ``out`` and to
        the raw device addresses ``peer_ptrs`` (other GPUs' buffers, same row offset), then
        ``flag_value`` is released at ``flag_ptrs``."""
        xdim, ydim = len(x), len(y)
        nloc = fm_local.shape[0]
        peers = (C.c_void_p * max(len(peer_ptrs), 1))(*peer_ptrs)
        flags = (C.c_void_p * max(len(flag_ptrs), 1))(*flag_ptrs)
        with torch.cuda.device(self.device):
            ws, wsz = self.workspace(xdim, ydim)
            with self._timed("ftle_stencil", 3):
              nat.check(nat.lib.dlb_ftle_stencil_gather(
                fm_local.data_ptr(), nloc, grow0, nat.dptr(x), xdim, nat.dptr(y), ydim,
                int(edge_order), float(t_span), out_row0, out_rows,
                out.data_ptr() if out_rows else fm_local.data_ptr(), peers, len(peer_ptrs), flags,
                len(flag_ptrs), int(flag_value), ws, wsz, self.stream()))
        return out

    # -- K3 ---------------------------------------------------------------------------
    def euler_stencil(self, mode, u_local, v_local, grow0, x, y, edge_order, out_row0, out_rows,
                      negate=False, xi_mode=nat.XI_NONE):
        xdim, ydim = len(x), len(y)
        nloc = u_local.shape[0]
        with torch.cuda.device(self.device):
            out = self.empty(out_rows, xdim)
            need_mask = mode in (nat.EULER_RATE, nat.EULER_RATIO)
            mask = self.empty(out_rows, xdim, dtype=torch.uint8) if need_mask else None
            xi = self.empty(out_rows, xdim, 2) if xi_mode != nat.XI_NONE else None
            ws, wsz = self.workspace(xdim, ydim)
            with self._timed("euler_stencil", 2):
              nat.check(nat.lib.dlb_euler_stencil(
                mode, u_local.data_ptr(), v_local.data_ptr(), nloc, grow0, nat.dptr(x), xdim,
                nat.dptr(y), ydim, int(edge_order), out_row0, out_rows, int(bool(negate)),
                out.data_ptr(), mask.data_ptr() if mask is not None else None,
                xi.data_ptr() if xi is not None else None, xi_mode, ws, wsz, self.stream()))
        return out, mask, xi

    # -- K0 ---------------------------------------------------------------------------
    def euler_stencil_flow(self, mode, flow_id, params, t, x, y, edge_order, out_row0, out_rows,
                           negate=False, xi_mode=nat.XI_NONE):
        """K0 fused into K3: the Eulerian field of an analytic flow at time ``t`` on
        meshgrid(x, y) without materialising (u, v)."""
        xdim, ydim = len(x), len(y)
        with torch.cuda.device(self.device):
            out = self.empty(out_rows, xdim)
            need_mask = mode in (nat.EULER_RATE, nat.EULER_RATIO)
            mask = self.empty(out_rows, xdim, dtype=torch.uint8) if need_mask else None
            xi = self.empty(out_rows, xdim, 2) if xi_mode != nat.XI_NONE else None
            ws, wsz = self.workspace(xdim, ydim)
            p = nat.darray(params)
            with self._timed("euler_stencil_flow", 2):
              nat.check(nat.lib.dlb_euler_stencil_flow(
                mode, flow_id, p, len(params), float(t), nat.dptr(x), xdim, nat.dptr(y), ydim,
                int(edge_order), out_row0, out_rows, int(bool(negate)), out.data_ptr(),
                mask.data_ptr() if mask is not None else None,
                xi.data_ptr() if xi is not None else None, xi_mode, ws, wsz, self.stream()))
        return out, mask, xi

    def flow_eval_grid(self, flow_id, params, t, x, y):
        xdim, ydim = len(x), len(y)
        with torch.cuda.device(self.device):
            u, v = self.empty(ydim, xdim), self.empty(ydim, xdim)
            ws, wsz = self.workspace(xdim, ydim)
            p = nat.darray(params)
            with self._timed("flow_eval"):
              nat.check(nat.lib.dlb_flow_eval_grid(
                flow_id, p, len(params), float(t), nat.dptr(x), xdim, nat.dptr(y), ydim,
                u.data_ptr(), v.data_ptr(), ws, wsz, self.stream()))
        return u, v

    def flow_eval_points(self, flow_id, dim, params, t, pts):
        """pts: [dim, n] device tensor -> [dim, n] velocities."""
        n = pts.shape[1]
        with torch.cuda.device(self.device):
            out = self.empty(dim, n)
            p = nat.darray(params)
            with self._timed("flow_eval"):
              nat.check(nat.lib.dlb_flow_eval_points(
                flow_id, p, len(params), float(t), pts[0].data_ptr(), pts[1].data_ptr(),
                pts[2].data_ptr() if dim == 3 else None, n, out[0].data_ptr(), out[1].data_ptr(),
                out[2].data_ptr() if dim == 3 else None, self.stream()))
        return out

    # -- ridge field ------------------------------------------------------------------
    def ridge_buffers(self, xdim, ydim):
        """(dd, conc, mask, scratch) for ``ridge_field`` - callers that sit behind a host sync
        allocate them earlier, while the device is still busy (cudaMalloc of result-sized blocks
        that cannot come from the cache costs milliseconds)."""
        with torch.cuda.device(self.device):
            return (self.empty(ydim, xdim), self.empty(ydim, xdim),
                    self.empty(ydim, xdim, dtype=torch.uint8), self.empty(ydim, xdim, 2))

    def ridge_field(self, field, xi, x, y, edge_order, threshold=None, bufs=None):
        xdim, ydim = len(x), len(y)
        with torch.cuda.device(self.device):
            dd, conc, mask, scratch = bufs if bufs is not None else self.ridge_buffers(xdim, ydim)
            ws, wsz = self.workspace(xdim, ydim)
            with self._timed("ridge_field", 2):
              nat.check(nat.lib.dlb_ridge_field(
                field.data_ptr(), xi.data_ptr(), nat.dptr(x), xdim, nat.dptr(y), ydim,
                int(edge_order), int(threshold is not None),
                float(threshold if threshold is not None else 0.0), dd.data_ptr(),
                conc.data_ptr(), mask.data_ptr(), scratch.data_ptr(), ws, wsz, self.stream()))
        return dd, conc, mask

    # -- contour segments -------------------------------------------------------------
    _contour_cap = 0  # largest segment count seen so far (+ head room): avoids re-launches

    def contour_buffers(self, xdim, ydim, cap=None):
        """(cap, count, cell, key_a, key_b, pts) record buffers for ``contour_launch``."""
        with torch.cuda.device(self.device):
            cap = int(cap or max(1 << 16, 8 * (xdim + ydim), self._contour_cap))
            return (cap, torch.zeros(1, dtype=torch.int64, device=self.device),
                    self.empty(cap, dtype=torch.int64), self.empty(cap, dtype=torch.int64),
                    self.empty(cap, dtype=torch.int64), self.empty(cap, 4))

    def contour_launch(self, z, mask, x, y, level=0.0, cap=None, bufs=None):
        """Enqueue the device pass of the contour extraction on the current stream; returns a
        handle for ``contour_fetch``.  Nothing is synchronised."""
        xdim, ydim = len(x), len(y)
        with torch.cuda.device(self.device):
            ws, wsz = self.workspace(xdim, ydim)
            if bufs is None:
                bufs = self.contour_buffers(xdim, ydim, cap)
            cap, count, cell, ka, kb, pts = bufs
            with self._timed("contour_segments"):
                nat.check(nat.lib.dlb_contour_segments(
                    z.data_ptr(), mask.data_ptr() if mask is not None else None, nat.dptr(x),
                    xdim, nat.dptr(y), ydim, float(level), cell.data_ptr(), ka.data_ptr(),
                    kb.data_ptr(), pts.data_ptr(), cap, count.data_ptr(), ws, wsz,
                    self.stream()))
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
        return dict(z=z, mask=mask, x=x, y=y, level=level, cap=cap, count=count, cell=cell,
                    ka=ka, kb=kb, pts=pts, done=done)

    def contour_fetch(self, h):
        """Host arrays (cell, key_a, key_b, pts[n, 4]) of a ``contour_launch``.  The records are
        read on the copy stream, so kernels enqueued on the main stream after the launch (the
        next slice of a sweep) keep running."""
        with torch.cuda.device(self.device):
            cs = self.copy_stream
            cs.wait_event(h["done"])
            with torch.cuda.stream(cs):
                n_host = torch.empty(1, dtype=torch.int64, pin_memory=True)
                h["count"].record_stream(cs)
                n_host.copy_(h["count"], non_blocking=True)
            cs.synchronize()
            n = int(n_host[0])
            if n > h["cap"]:
                # did not fit: run again with room.  Outside the copy-stream context, i.e. on
                # the caller's (main) stream like every other launch that touches the workspace
                Engine._contour_cap = n + n // 4
                return self.contour_fetch(self.contour_launch(
                    h["z"], h["mask"], h["x"], h["y"], h["level"], cap=n + n // 8))
            Engine._contour_cap = max(Engine._contour_cap, n + n // 4)
            out = []
            with torch.cuda.stream(cs):
                for key in ("cell", "ka", "kb", "pts"):
                    t = h[key][:n]
                    t.record_stream(cs)
                    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=n > 4096)
                    host.copy_(t, non_blocking=True)
                    out.append(host)
            cs.synchronize()
        return tuple(t.numpy() for t in out)

    def contour_segments(self, z, mask, x, y, level=0.0):
        """Device pass of the contour extraction: returns host arrays (cell, key_a, key_b,
        pts[n, 4]) of the segments of the ``level`` contour of ``z`` (device tensor [ydim, xdim];
        ``mask`` uint8 device tensor or None)."""
        return self.contour_fetch(self.contour_launch(z, mask, x, y, level))

    # -- percentile -------------------------------------------------------------------
    def percentile(self, values, q):
        """``np.percentile(values, q)`` (default 'linear' method) of a device tensor without
        moving it: the two neighbouring order statistics come from the radix select in
        csrc/select.cu, the interpolation below follows numpy's ``_quantile`` / ``_lerp``
        expression by expression so the result is bit-identical
        (numpy/lib/_function_base_impl.py: virtual index ``(n - 1) * q``, ``gamma = virtual -
        floor(virtual)``, ``a + (b - a) * gamma`` or ``b - (b - a) * (1 - gamma)`` for
        ``gamma >= 0.5``)."""
        qf = np.true_divide(np.float64(q), 100)
        if not (0.0 <= qf <= 1.0):
            raise ValueError("Percentiles must be in the range [0, 100]")
        n = values.numel()
        if n == 0:
            return float("nan")
        virtual = (n - 1) * qf
        prev = np.floor(virtual)
        if virtual >= n - 1:
            lo = hi = n - 1
        else:
            lo, hi = int(prev), int(prev) + 1
        gamma = np.float64(virtual - prev)
        with torch.cuda.device(self.device):
            flat = values.reshape(-1)
            if not flat.is_contiguous():
                flat = flat.contiguous()
            if getattr(self, "_select_bufs", None) is None:  # reused: a sweep calls this per slice
                nbytes = int(nat.lib.dlb_select_scratch_bytes())
                self._select_bufs = (self.empty(3), torch.empty(nbytes, dtype=torch.uint8,
                                                                device=self.device),
                                     torch.empty(3, dtype=_F64, pin_memory=True))
            res, scratch, res_host = self._select_bufs  # res: [a_lo, a_hi, nan count (uint64 bits)]
            nbytes = scratch.numel()
            ranks = (C.c_int64 * 2)(lo, hi)
            with self._timed("order_statistics", 17):
                nat.check(nat.lib.dlb_order_statistics(
                    flat.data_ptr(), n, ranks, 2, res.data_ptr(), res[2:].data_ptr(),
                    scratch.data_ptr(), nbytes, self.stream()))
            res_host.copy_(res, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            host = res_host.clone()
        if int(host[2:].view(torch.int64)[0]) != 0:
            return float("nan")
        a, b = np.float64(host[0].item()), np.float64(host[1].item())
        diff = np.subtract(b, a)
        if gamma >= 0.5:
            return float(np.subtract(b, diff * (1 - gamma)))
        return float(np.add(a, diff * gamma))

    # -- FP64 peak --------------------------------------------------------------------
    def dfma_peak(self, iters: int = 4096):
        fl, ms = C.c_double(0.0), C.c_double(0.0)
        with torch.cuda.device(self.device):
            nat.check(nat.lib.dlb_dfma_peak(iters, C.byref(fl), C.byref(ms)))
        return fl.value, ms.value
