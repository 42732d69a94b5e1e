import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); os.environ['CSLVR_B200_QUIET']='1'
import numpy as np, bench
from cases import eismint_dome_model, product_context
def run(m, label, **kw):
    for sm in ('auto', 'block', 'column'):
        c = product_context(m, preconditioner='mg', mg_smoother=sm, relative_tolerance=1e-9, maximum_iterations=40, krylov_rtol=1e-5, **kw)
        u = np.full(2*m.n_nodes, 3e-16)
        t0=time.time(); st = c.newton_solve(u); t=time.time()-t0
        print(label, sm, 'conv', st['converged'], 'newton', st['iterations'], 'krylov', st['krylov_iterations'], '%.3fs'%t)
        c.close()
run(eismint_dome_model(n=(200,200,5)), 'eismint200', use_pressure_bc=False, relaxation_parameter=0.7)
run(bench.build_model(1), 'ismipC260')
