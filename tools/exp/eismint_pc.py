import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); os.environ['CSLVR_B200_QUIET']='1'
import numpy as np
from cases import eismint_dome_model, product_context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
m = eismint_dome_model(n=(n, n, 5))
for pc, kw in (('mg', {}), ('mg', dict(krylov_rtol=1e-8)), ('chebyshev', {}), ('column', {}), ('block_jacobi', {})):
    p = dict(use_pressure_bc=False, relative_tolerance=1e-9, relaxation_parameter=0.7, maximum_iterations=40, krylov_rtol=1e-5, preconditioner=pc)
    p.update(kw)
    c = product_context(m, **p)
    u = np.full(2*m.n_nodes, 3e-16)
    t0=time.time(); st = c.newton_solve(u); t=time.time()-t0
    r = st['residuals']
    print(pc, kw, 'conv', st['converged'], 'its', st['iterations'], 'krylov', st['krylov_iterations'], '%.2fs'%t, 'rel res every 5:', ['%.1e'%(x/r[0]) for x in r[::5]])
    c.close()
