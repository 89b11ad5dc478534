import sys
import numpy as np, torch
sys.path.insert(0, '.')
from dynlab_b200._engine import Engine
from dynlab_b200._resolve import resolve_flow
from dynlab_b200 import flows, _native as nat
eng = Engine.get()
_, fid, _, params = resolve_flow(flows.bickley_jet)
x, y = np.linspace(0, 20000, 8192), np.linspace(-4000, 4000, 8192)
u, v = eng.flow_eval_grid(fid, params, 0.0, x, y)
for _ in range(3):
    eng.euler_stencil(nat.EULER_S_MIN, u, v, 0, x, y, 2, 0, 8192, False, 0)
torch.cuda.synchronize(); print('ok')
