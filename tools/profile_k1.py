"""One warm-up + two timed launches of K1 (and K2) on config 3, for ncu."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from dynlab_b200._engine import Engine
from dynlab_b200._resolve import resolve_flow
from dynlab_b200 import flows
eng = Engine.get()
_, fid, _, params = resolve_flow(flows.double_gyre)
x, y = np.linspace(0, 2, 8192), np.linspace(0, 1, 4096)
out = eng.empty(4096, 8192, 2); fo = eng.empty(4096, 8192)
for _ in range(3):
    eng.flow_map(fid, params, x, y, 0, 4096, 10.0, 0.0, 1e-8, 1e-8, out=out)
    eng.ftle_stencil(out, 0, x, y, 2, 10.0, 0, 4096, out=fo)
torch.cuda.synchronize()
print('ok', eng.read_counters())
