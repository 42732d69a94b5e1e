import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); os.environ['CSLVR_B200_QUIET']='1'
import numpy as np, bench
from cases import product_context
m = bench.build_model(1)
c = product_context(m, preconditioner='mg', relative_tolerance=1e-9, maximum_iterations=1, krylov_rtol=1e-5, krylov_maxit=8)
u = np.full(2*m.n_nodes, 3e-16)
c.newton_solve(u)
