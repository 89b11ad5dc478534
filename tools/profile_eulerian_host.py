"""cProfile of AttractionRate.compute(x, y, u=, v=) with 8192x8192 host arrays (config 2)."""
import cProfile, os, pstats, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynlab_b200.diagnostics import AttractionRate
N = 8192
x, y = np.linspace(0, 20000, N), np.linspace(-4000, 4000, N)
rng = np.random.default_rng(0)
u, v = rng.normal(size=(N, N)), rng.normal(size=(N, N))
d = AttractionRate()
d.compute(x, y, u=u, v=v, edge_order=2)
d.compute(x, y, u=u, v=v, edge_order=2)
pr = cProfile.Profile(); pr.enable()
d.compute(x, y, u=u, v=v, edge_order=2)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
