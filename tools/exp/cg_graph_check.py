"""does the CG-iteration graph get built?  python tools/exp/cg_graph_check.py [nxy]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from cslvr_b200 import capi
m = bench.build_model(1, nxy=int(sys.argv[1]) if len(sys.argv) > 1 else 100)
ctx = capi.context_from_model(m, verbose=True)
for it in range(2):
	u = np.full(2 * m.n_nodes, 3e-16)
	st = ctx.newton_solve(u)
	print({k: st[k] for k in ('converged', 'iterations', 'krylov_iterations', 'gpu_launches', 'host_launches', 't_krylov_ms')})
