import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); os.environ['CSLVR_B200_QUIET']='1'
import numpy as np, bench
from cases import product_context
from cslvr_b200 import capi
m = bench.build_model(1)
c = product_context(m, relative_tolerance=1e-9, maximum_iterations=25, krylov_rtol=1e-5)
c.set_field(capi.FIELD_B, m.vertex_field(m.B))
u = np.full(2*m.n_nodes, 3e-16)
c.newton_solve(u.copy())
t0=time.time(); st=c.newton_solve(u); t1=time.time()
print('newton %.3f s its %d/%d' % (t1-t0, st['iterations'], st['krylov_iterations']))
for rep in range(2):
    t0=time.time(); uz=c.solve_vert_velocity(u); t1=time.time()
    print('u_z %.3f s, bicgstab its %d' % (t1-t0, c.last_vert_iterations))
    t0=time.time(); p=c.solve_pressure(u); t1=time.time()
    print('p   %.3f s, pcg its %d' % (t1-t0, c.last_projection_iterations))
