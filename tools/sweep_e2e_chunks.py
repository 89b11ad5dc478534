import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from dynlab_b200.diagnostics import FTLE
from dynlab_b200.flows import double_gyre
x, y = np.linspace(0, 2, 8192), np.linspace(0, 1, 4096)
for chunks, frac in ((4, 16), (4, 32), (4, 64), (6, 32), (8, 32), (3, 16), (5, 64)):
    FTLE.pipeline_chunks, FTLE.pipeline_last_fraction = chunks, frac
    d = FTLE()
    for _ in range(3): d.compute(x, y, double_gyre, (10, 0), edge_order=2, rtol=1e-8, atol=1e-8)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(8): d.compute(x, y, double_gyre, (10, 0), edge_order=2, rtol=1e-8, atol=1e-8)
    torch.cuda.synchronize(); print(chunks, frac, round((time.perf_counter() - t0) / 8 * 1e3, 3), 'ms', flush=True)
