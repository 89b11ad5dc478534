"""K1 throughput on BASELINE config 4 (divergent step counts) for refill settings."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from dynlab_b200._engine import Engine
from dynlab_b200._resolve import resolve_flow
from dynlab_b200 import flows
eng = Engine.get()
def run(fn, n=3):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
cases = [('bead_on_a_rotating_hoop', np.linspace(-1, 1, 4096) * 3, np.linspace(-1, 1, 4096) * 2.5, (0.0, 2.0)),
         ('duffing_oscillator', np.linspace(-2, 2, 4096), np.linspace(-2, 2, 4096), (0.0, 10.0)),
         ('duffing_oscillator', np.linspace(-2, 2, 4096), np.linspace(-2, 2, 4096), (10.0, 0.0)),
         ('double_gyre', np.linspace(0, 2, 4096), np.linspace(0, 1, 2048), (20.0, 0.0))]
for name, x, y, t in cases:
    _, fid, _, params = resolve_flow(getattr(flows, name))
    out = eng.empty(len(y), len(x), 2)
    k1 = lambda: eng.flow_map(fid, params, x, y, 0, len(y), t[0], t[1], 1e-8, 1e-8, out=out)
    for refill in (32, 16, 8):
        os.environ['DLB_RK45_REFILL'] = str(refill)
        ms = run(k1)
        c = eng.read_counters()
        att = (c['nfev'] - 2 * c['seeds']) / 6
        print(f"{name:26s} t={t} refill={refill:2d}: {ms:8.2f} ms  {len(x)*len(y)/ms/1e3:8.1f} Mseeds/s  nfev/seed {c['nfev']/c['seeds']:7.1f}  attempts/s {att/ms/1e6:6.2f} G  failed {c['failed']}", flush=True)
