"""Time K1 on config 3 for refill-threshold / chunk settings (env read per launch)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from dynlab_b200._engine import Engine
from dynlab_b200._resolve import resolve_flow
from dynlab_b200 import flows
eng = Engine.get()
_, fid, _, params = resolve_flow(flows.double_gyre)
x, y = np.linspace(0, 2, 8192), np.linspace(0, 1, 4096)
out = eng.empty(4096, 8192, 2)
def run(n=3):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.flow_map(fid, params, x, y, 0, 4096, 10.0, 0.0, 1e-8, 1e-8, out=out); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
run(2)
for refill in (1, 2, 4, 8, 16, 32):
    for chunk in (32, 128, 512):
        os.environ['DLB_RK45_REFILL'] = str(refill); os.environ['DLB_RK45_CHUNK'] = str(chunk)
        print(f'refill {refill:2d} chunk {chunk:4d}: {run():8.3f} ms', flush=True)
c = eng.read_counters(); print(c, 'attempts/seed', (c['nfev'] - 2 * c['seeds']) / 6 / c['seeds'])
