import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); os.environ['CSLVR_B200_QUIET']='1'
import numpy as np, bench, time
from cases import product_context
m = bench.build_model(1)
for pc, kw in (('chebyshev',dict(pc_degree=2)),('chebyshev',dict(pc_degree=4)),('chebyshev',dict(pc_degree=8, pc_eig_ratio=100.))):
    c = product_context(m, preconditioner=pc, relative_tolerance=1e-9, maximum_iterations=25, krylov_rtol=1e-5, **kw)
    u = np.full(2*m.n_nodes, 3e-16)
    t0=time.time(); st = c.newton_solve(u); t=time.time()-t0
    print(pc, kw, 'newton', st['iterations'], 'krylov', st['krylov_iterations'], 'launches', st['gpu_launches'], 'krylov_ms %.1f total %.3f s'%(st['t_krylov_ms'], t), st['converged'])
    c.close()
