import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynlab_b200 import _native as nat
from dynlab_b200._engine import Engine
from dynlab_b200.flows import FLOW_REGISTRY
from oracle import py_oracle as po
eng = Engine.get()
rng = np.random.default_rng(8)
grids = [(np.linspace(0.1, 1.9, 97), np.linspace(0.05, 0.95, 71)),
         (np.sort(rng.uniform(0.1, 1.9, 64)), np.linspace(0.05, 0.95, 130)),
         (np.linspace(0.1, 1.9, 33), np.sort(rng.uniform(0.05, 0.95, 67))),
         (np.sort(rng.uniform(0.1, 1.9, 5)), np.sort(rng.uniform(0.05, 0.95, 3)))]
modes = [(nat.EULER_S_MIN, nat.XI_NONE), (nat.EULER_S_MAX, nat.XI_LAPACK), (nat.EULER_S_MIN, nat.XI_FORCED),
         (nat.EULER_RATE, nat.XI_NONE), (nat.EULER_RATIO, nat.XI_NONE)]
for name, (fid, dim, _) in FLOW_REGISTRY.items():
    if dim != 2: continue
    params = po.flow_params(name)
    sc = 1e4 if name == "bickley_jet" else 1.0
    for gi, (x0, y0) in enumerate(grids):
        x, y = x0 * sc, (y0 - 0.5) * sc
        ydim = len(y); eo = 1 + gi % 2
        for mode, xim in modes:
            neg = mode == nat.EULER_S_MIN and xim == nat.XI_FORCED
            got = eng.euler_stencil_flow(mode, fid, params, 0.37, x, y, eo, 0, ydim, neg, xim)
            if ydim >= 8:
                for r0, rows in ((0, 3), (2, ydim - 5), (ydim - 2, 2)):
                    part = eng.euler_stencil_flow(mode, fid, params, 0.37, x, y, eo, r0, rows, neg, xim)
                    d = (part[0].view(torch.int64) - got[0][r0:r0 + rows].view(torch.int64)).abs()
                    if int(d.max()):
                        print(name, "grid", gi, "eo", eo, "mode", mode, xim, "slab", r0, rows, "ulp", int(d.max()), "at", np.unravel_index(int(d.argmax()), d.shape))
print("done")
