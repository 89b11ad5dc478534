import sys, numpy as np
sys.path.insert(0, '.')
from tests.test_oracle_golden import EUL_GRIDS
import dynlab_b200.diagnostics as D
g = np.load('tests/golden/eulerian.npz')
for name in ['aduffing_eo2', 'dg_eo2', 'bickley_eo2']:
    x, y, eo = EUL_GRIDS[name]
    u, v = g[name + '__u'], g[name + '__v']
    dudy, dudx = np.gradient(u, y, x, edge_order=eo); dvdy, dvdx = np.gradient(v, y, x, edge_order=eo)
    for kind in ('attracting', 'repelling'):
        for forced in (True, False):
            tag = f"{name}__iles_{kind}_{'forced' if forced else 'raw'}"
            il = D.iLES(); il.compute(x, y, u=u, v=v, kind=kind, edge_order=eo, force_eigenvectors=forced, debug=True)
            xi = np.ma.getdata(il.Xi_max); ref = g[tag + '__xi']
            diff = np.abs(xi - ref).max(-1)
            bad = np.argwhere(diff > 1e-7)
            print(tag, 'bad', len(bad), 'of', diff.size)
            for i, j in bad[:6]:
                print('   ', i, j, 'mine', xi[i, j], 'ref', ref[i, j], 'a,b,d', dudx[i, j], 0.5 * (dudy[i, j] + dvdx[i, j]), dvdy[i, j])
