"""Small end-to-end run of every kernel, for compute-sanitizer."""
import os, sys
import numpy as np
sys.path.insert(0, '.')
from dynlab_b200.diagnostics import FTLE, LCS, AttractionRate, TrajectoryRepulsionRate, iLES
from dynlab_b200.flows import double_gyre, bickley_jet, lorenz
from dynlab_b200.utils import rk45_wrapper
for bulk in ("1", "0"):
    os.environ["DLB_STENCIL_BULK"] = bulk
    for (nx, ny, eo) in ((130, 70, 2), (131, 67, 1), (257, 200, 2), (3, 3, 2), (2, 2, 1)):
        x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
        f = FTLE().compute(x, y, double_gyre, (3, 0), edge_order=eo, rtol=1e-6, atol=1e-6)
        l = LCS().compute(x, y, f=double_gyre, t=(3, 0), edge_order=eo, percentile=50, rtol=1e-6, atol=1e-6)
        a = AttractionRate().compute(x * 1e4, y * 8e3 - 4e3, f=bickley_jet, t=0.1, edge_order=eo)
        u, v = double_gyre(0.2, np.meshgrid(x, y))
        r = TrajectoryRepulsionRate().compute(x, y, u=u, v=v, edge_order=eo)
        i = iLES().compute(x, y, u=u, v=v, kind='repelling', edge_order=eo, percentile=50)
        print(bulk, nx, ny, eo, float(f.max()), len(l), float(a.min()), float(r.max()), len(i), flush=True)
print(rk45_wrapper(lorenz, (0, 0.5), (1.0, 2.0, 3.0), rtol=1e-8, atol=1e-8)[-1])
print("SANITIZE_SMALL_OK")
