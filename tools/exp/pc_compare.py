import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); os.environ['CSLVR_B200_QUIET']='1'
import numpy as np, bench, time
from cases import product_context
nxy = int(os.environ.get('NXY', '260'))
bench.NX = bench.NY = nxy
bench.L = 20000.0 * nxy / 260 if os.environ.get('SCALE_L', '1') == '1' else 20000.0
m = bench.build_model(1)
print('mesh', nxy, 'tets', m.num_cells)
for pc, kw in (('chebyshev',{}),('mg',{})):
    c = product_context(m, preconditioner=pc, relative_tolerance=1e-9, maximum_iterations=25, krylov_rtol=1e-5, **kw)
    u = np.full(2*m.n_nodes, 3e-16)
    t0=time.time(); st = c.newton_solve(u); t=time.time()-t0
    print(pc, kw, 'newton', st['iterations'], 'krylov', st['krylov_iterations'], 'launches', st['gpu_launches'], 'krylov_ms %.1f total %.3f s'%(st['t_krylov_ms'], t), st['converged'])
    c.close()
