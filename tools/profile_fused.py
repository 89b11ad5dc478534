"""A few launches of the fused flow + stencil kernel (bickley_jet 8192x8192, S_MIN) for ncu."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynlab_b200 import flows, _native as nat
from dynlab_b200._engine import Engine
from dynlab_b200._resolve import resolve_flow
eng = Engine.get()
N = 8192
x, y = np.linspace(0, 20000, N), np.linspace(-4000, 4000, N)
_, fid, _, params = resolve_flow(flows.bickley_jet)
for _ in range(3):
    out = eng.euler_stencil_flow(nat.EULER_S_MIN, fid, params, 0.0, x, y, 2, 0, N)
torch.cuda.synchronize()
print("ok", float(out[0][5, 5]))
