"""Time K1 (and K2) on config 3."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from dynlab_b200._engine import Engine
from dynlab_b200._resolve import resolve_flow
from dynlab_b200 import flows
eng = Engine.get()
_, fid, _, params = resolve_flow(flows.double_gyre)
x, y = np.linspace(0, 2, 8192), np.linspace(0, 1, 4096)
out = eng.empty(4096, 8192, 2)
def run(fn, n=5):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))
k1 = lambda: eng.flow_map(fid, params, x, y, 0, 4096, 10.0, 0.0, 1e-8, 1e-8, out=out)
run(k1, 2)
for env in sys.argv[1:] or ['']:
    for kv in env.split(','):
        if '=' in kv:
            k, v = kv.split('='); os.environ[k] = v
    print(env, 'K1 min/median ms', run(k1), flush=True)
c = eng.read_counters(); print(c)
fo = eng.empty(4096, 8192)
k2 = lambda: eng.ftle_stencil(out, 0, x, y, 2, 10.0, 0, 4096, out=fo)
print('K2 min/median ms', run(k2, 10), ' GB/s', 24 * 8192 * 4096 / run(k2, 10)[0] / 1e6)
