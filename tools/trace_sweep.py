"""Host/device timeline of ftle_time_sweep(ridge=True): wall-clock stamps of the host stages and
CUDA-event times of the device stages for a few slices (where does the device sit idle?)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dynlab_b200.flows as flows
from dynlab_b200 import _native as nat
from dynlab_b200._contour import contour_lines_finish
from dynlab_b200._engine import Engine
from dynlab_b200._resolve import resolve_flow

eng = Engine.get()
xs, ys = np.linspace(0, 2, 4096), np.linspace(0, 1, 2048)
xdim, ydim = len(xs), len(ys)
_, fid, _, params = resolve_flow(flows.double_gyre)
spans = [(k * 10.0 / 64 + 20.0, k * 10.0 / 64) for k in range(10)]
fm = eng.empty(ydim, xdim, 2)
keep = []
pending = None
T0 = time.perf_counter()
def now(): return (time.perf_counter() - T0) * 1e3
base = torch.cuda.Event(enable_timing=True); base.record()
rows = []
for k, (t0, tf) in enumerate(spans):
    st = {"k": k, "h_start": now()}
    e = {n: torch.cuda.Event(enable_timing=True) for n in ("k1a", "k1b", "rf", "ct")}
    e["k1a"].record()
    eng.flow_map(fid, params, xs, ys, 0, ydim, t0, tf, 1e-8, 1e-8, out=fm)
    e["k1b"].record()
    st["h_k1_enq"] = now()
    if pending is not None:
        keep.append(contour_lines_finish(eng, pending))
    st["h_finish"] = now()
    f_out, xi_out = eng.empty(ydim, xdim), eng.empty(ydim, xdim, 2)
    rb, cb = eng.ridge_buffers(xdim, ydim), eng.contour_buffers(xdim, ydim)
    st["h_alloc"] = now()
    field, xi = eng.ftle_stencil(fm, 0, xs, ys, 2, 20.0, 0, ydim, nat.XI_LAPACK, out=f_out, xi_out=xi_out)
    thr = eng.percentile(field, 80)
    st["h_pct"] = now()
    dd, conc, mask = eng.ridge_field(field, xi, xs, ys, 2, thr, bufs=rb)
    e["rf"].record()
    pending = eng.contour_launch(dd, mask, xs, ys, bufs=cb)
    e["ct"].record()
    st["h_launch"] = now()
    keep.append((field, dd, mask))
    rows.append((st, e))
torch.cuda.synchronize()
for st, e in rows[4:9]:
    d = {n: base.elapsed_time(ev) for n, ev in e.items()}
    print(f"slice {st['k']}: host start {st['h_start']:8.2f} k1_enq {st['h_k1_enq']:8.2f} finish {st['h_finish']:8.2f} alloc {st['h_alloc']:8.2f} "
          f"pct {st['h_pct']:8.2f} launch {st['h_launch']:8.2f} | dev K1 {d['k1a']:8.2f}..{d['k1b']:8.2f} ridge_end {d['rf']:8.2f} contour_end {d['ct']:8.2f}")
