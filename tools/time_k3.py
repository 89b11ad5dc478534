"""Time K0 + K3 on config 2 (8192x8192 bickley_jet, edge_order 2), device-resident."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from dynlab_b200._engine import Engine
from dynlab_b200._resolve import resolve_flow
from dynlab_b200 import flows, _native as nat
eng = Engine.get()
_, fid, _, params = resolve_flow(flows.bickley_jet)
x, y = np.linspace(0, 20000, 8192), np.linspace(-4000, 4000, 8192)
def run(fn, n=10):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))
u, v = eng.flow_eval_grid(fid, params, 0.0, x, y)
print('K0 ms', run(lambda: eng.flow_eval_grid(fid, params, 0.0, x, y)), 'GB/s (16 B/pt)', 16 * 8192 * 8192 / run(lambda: eng.flow_eval_grid(fid, params, 0.0, x, y))[0] / 1e6)
for name, mode, xi in (('s_min', nat.EULER_S_MIN, 0), ('s_max', nat.EULER_S_MAX, 0), ('rate', nat.EULER_RATE, 0), ('ratio', nat.EULER_RATIO, 0), ('s_min+xi', nat.EULER_S_MIN, 2)):
    t = run(lambda: eng.euler_stencil(mode, u, v, 0, x, y, 2, 0, 8192, False, xi))
    b = (40 if xi else 24) + (1 if mode >= 2 else 0)
    print(f'K3 {name:9s} ms {t[0]:.4f} / {t[1]:.4f}   GB/s {b * 8192 * 8192 / t[0] / 1e6:.0f}', flush=True)
