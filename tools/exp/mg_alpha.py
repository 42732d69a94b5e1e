import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); os.environ['CSLVR_B200_QUIET']='1'
import numpy as np, bench, time
from cases import product_context
m = bench.build_model(1)
print('tets', m.num_cells)
combos = [tuple(float(x) for x in a.split(",")) for a in sys.argv[1:]]
for hf, a0, a, gm, dg, rt, f32 in combos:
    os.environ['BP_MG_HORIZONTAL_FIRST'] = str(hf); os.environ['BP_MG_ALPHA0'] = str(a0); os.environ["BP_MG_ALPHA"] = str(a); os.environ["BP_MG_GAMMA"] = str(int(gm)); os.environ["BP_MG_FP32"] = str(int(f32))
    c = product_context(m, preconditioner='mg', relative_tolerance=1e-9, maximum_iterations=25, krylov_rtol=1e-5, pc_degree=int(dg), pc_eig_ratio=rt, verbose=(os.environ.get('V') == '1'))
    u = np.full(2*m.n_nodes, 3e-16)
    try:
        c.set_params(maximum_iterations=1, krylov_maxit=2); c.newton_solve(u.copy()); c.set_params(maximum_iterations=25, krylov_maxit=10000)
        t0=time.time(); st = c.newton_solve(u); t=time.time()-t0
        print('hfirst %.1f alpha0 %.1f alpha %.1f gamma %d deg %d ratio %g f32 %d' % (hf, a0, a, gm, dg, rt, f32), 'newton', st['iterations'], 'krylov', st['krylov_iterations'], 'krylov_ms %.1f total %.3f s'%(st['t_krylov_ms'], t), st['converged'])
    except Exception as e:
        print('hfirst %.1f alpha0 %.1f alpha %.1f' % (hf, a0, a), 'FAILED', str(e)[:100])
    c.close()
