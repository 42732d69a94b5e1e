"""Time (and give ncu something to capture) the cell-assembly kernels on an ISMIP-HOM A mesh.

    python tools/prof_assemble.py [nxy=260] [modes=tiled,gather] [repeats=20]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('CSLVR_B200_QUIET', '1')
import bench
from cslvr_b200 import capi

nxy = int(sys.argv[1]) if len(sys.argv) > 1 else 260
modes = (sys.argv[2] if len(sys.argv) > 2 else 'tiled,gather').split(',')
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
t0 = time.time()
model = bench.build_model(1, nxy=nxy)
s = bench.slab_from_model(model)
u = bench.synthetic_state(s)
print("model %.1f s, %d tets" % (time.time() - t0, model.num_cells), flush=True)
for mode in modes:
	t0 = time.time()
	ctx = capi.context_from_model(model, assembly=mode)
	t_setup = time.time() - t0
	ctx.assemble(u, want_F=True, want_J=True, J=False)
	what = capi.ASSEMBLE_F | capi.ASSEMBLE_J
	t_cell = ctx.time_assemble(what | 4, repeats=reps, flush_l2=True)
	t_pass = ctx.time_assemble(what, repeats=reps, flush_l2=True)
	nt, nv, nn = model.num_cells, model.num_vertices, model.n_nodes
	b = 16 * nt + 32 * nv + 24 * nn + 8 * ctx.nnz + 16 * nn
	print("%-8s setup %.1f s  cell %.4f ms  pass %.4f ms  %.1f Mtets/s  roofline %.3f of 6545.6 GB/s" % (
		mode, t_setup, t_cell, t_pass, nt / t_cell / 1e3, b / t_cell / 1e6 / 6545.6), flush=True)
	ctx.close()
