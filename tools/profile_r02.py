"""Round-2 ncu workload: after warm-up, ONE launch each of K1 (config 3), K2 (config 3), K3 S_MIN
and K3 S_MIN + Xi (config 2).
    python tools/profile_r02.py && ncu --set full --clock-control none --import-source on \
        -k regex:"rk45_kernel|stencil_tma_kernel" -s <warm-up launches> -c 4 -o gpurun_out/r02 \
        python tools/profile_r02.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynlab_b200 import _native as nat  # noqa: E402
from dynlab_b200 import flows  # noqa: E402
from dynlab_b200._engine import Engine  # noqa: E402
from dynlab_b200._resolve import resolve_flow  # noqa: E402

eng = Engine.get()
_, fid, _, params = resolve_flow(flows.double_gyre)
x, y = np.linspace(0, 2, 8192), np.linspace(0, 1, 4096)
fm, fo = eng.empty(4096, 8192, 2), eng.empty(4096, 8192)
_, bid, _, bparams = resolve_flow(flows.bickley_jet)
xb, yb = np.linspace(0, 20000, 8192), np.linspace(-4000, 4000, 8192)
u, v = eng.flow_eval_grid(bid, bparams, 0.0, xb, yb)
reps = int(os.environ.get("REPS", "2"))  # reps - 1 warm-up rounds: 4 matching launches each
for _ in range(reps):
    eng.flow_map(fid, params, x, y, 0, 4096, 10.0, 0.0, 1e-8, 1e-8, out=fm)
    eng.ftle_stencil(fm, 0, x, y, 2, 10.0, 0, 4096, out=fo)
    eng.euler_stencil(nat.EULER_S_MIN, u, v, 0, xb, yb, 2, 0, 8192, False, 0)
    eng.euler_stencil(nat.EULER_S_MIN, u, v, 0, xb, yb, 2, 0, 8192, True, 2)
torch.cuda.synchronize()
print("ok", eng.read_counters())
