"""cProfile of the host side of ftle_time_sweep(ridge=True) on a few config-5 slices."""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dynlab_b200.flows as flows  # noqa: E402
from dynlab_b200.sweep import ftle_time_sweep  # noqa: E402

x, y = np.linspace(0, 2, 4096), np.linspace(0, 1, 2048)
spans = [(k * 10.0 / 64 + 20.0, k * 10.0 / 64) for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8)]
kw = dict(edge_order=2, percentile=80, ridge=True, to_host=False, rtol=1e-8, atol=1e-8)
ftle_time_sweep(x, y, flows.double_gyre, spans[:1], **kw)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
ftle_time_sweep(x, y, flows.double_gyre, spans, **kw)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
