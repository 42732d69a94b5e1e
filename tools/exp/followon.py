import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); os.environ['CSLVR_B200_QUIET']='1'
import numpy as np, bench, torch
from cases import product_context
from cslvr_b200 import capi
m = bench.build_model(1); c = product_context(m)
c.set_field(capi.FIELD_B, m.vertex_field(m.B))
u = np.full(2*m.n_nodes, 3e-16); st = c.newton_solve(u); print('newton', st['iterations'], st['krylov_iterations'], st['t_total_ms'])
ud = torch.from_numpy(u).cuda()
for name, f in (('pressure', c.solve_pressure), ('vert_velocity', c.solve_vert_velocity)):
    for rep in range(2):
        torch.cuda.synchronize(); t0=time.perf_counter(); out = f(ud); torch.cuda.synchronize(); t=time.perf_counter()-t0
    print(name, '%.2f ms'%(t*1e3), 'iters', getattr(c,'last_projection_iterations',None), getattr(c,'last_vert_iterations',None), float(out.min()), float(out.max()))
