import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); os.environ['CSLVR_B200_QUIET']='1'
import numpy as np
from cases import ismip_model, product_context
from cslvr_b200 import capi
from cslvr_b200.partition import partition_model, rank_context, gather_owned
nxy = int(os.environ.get('NXY', '408'))
nslab = int(os.environ.get('NSLAB', '1'))
m = ismip_model('A', L=float(os.environ.get('LKM', '80')) * 1000.0, n=(nxy, nxy, 5))
print('tets', m.num_cells, 'slabs', nslab)
for hf, a in [tuple(float(x) for x in s.split(',')) for s in sys.argv[1:]]:
    os.environ['BP_MG_HORIZONTAL_FIRST'] = str(hf); os.environ['BP_MG_ALPHA0'] = str(a); os.environ['BP_MG_ALPHA'] = str(a)
    kw = dict(preconditioner='mg', relative_tolerance=1e-9, maximum_iterations=25, krylov_rtol=1e-5)
    if nslab == 1:
        c = product_context(m, **kw)
        u = np.full(2*m.n_nodes, 3e-16)
        t0=time.time(); st = c.newton_solve(u); t=time.time()-t0
        c.close()
    else:
        rms = partition_model(m, nslab)
        ctxs = [rank_context(m, rm, **kw) for rm in rms]
        us = [np.full(2 * rm.n_nodes, 3e-16) for rm in rms]
        t0=time.time(); st = capi.newton_solve_group(ctxs, us); t=time.time()-t0
        for c in ctxs: c.close()
    print('hfirst %.1f alpha %.1f' % (hf, a), 'newton', st['iterations'], 'krylov', st['krylov_iterations'], 'krylov_ms %.1f total %.3f s'%(st['t_krylov_ms'], t), st['converged'], flush=True)
