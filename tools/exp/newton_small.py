import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); os.environ['CSLVR_B200_QUIET']='1'
import numpy as np, bench, time
from cases import product_context
m = bench.build_model(1); c = product_context(m)
c.set_params(relative_tolerance=1e-9, maximum_iterations=int(sys.argv[1]) if len(sys.argv)>1 else 25, krylov_rtol=1e-5, krylov_maxit=int(sys.argv[2]) if len(sys.argv)>2 else 10000)
u = np.full(2*m.n_nodes, 3e-16)
t0=time.time(); st = c.newton_solve(u); print({k:v for k,v in st.items() if k!='residuals'}, time.time()-t0)
