"""Newton / CG iteration counts of the multigrid-preconditioned solve on ISMIP-HOM A for a list of environment
settings:   python tools/exp/mg_sweep.py nxy L_km "BP_MG_GAMMA=2" "BP_MG_ALPHA=1.5 BP_MG_TAIL=16384" ...   ("" = defaults)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault('CSLVR_B200_QUIET', '1')
import bench
from cslvr_b200 import capi

nxy, Lkm = int(sys.argv[1]), float(sys.argv[2])
bench.L = Lkm * 1e3
model = bench.build_model(1, nxy=nxy)
print("ISMIP-HOM A  L = %g km  %dx%dx5  = %d tets,  dx = %.0f m" % (Lkm, nxy, nxy, model.num_cells, Lkm * 1e3 / nxy), flush=True)
for spec in sys.argv[3:] or [""]:
	kw = {}
	for tok in spec.split():
		k, v = tok.split('=')
		if k.startswith('BP_'):
			os.environ[k] = v
		else:
			kw[k] = float(v) if '.' in v else int(v)
	slabs = int(kw.pop('slabs', 1))
	if slabs > 1:          # the partitioned path on ONE device: k contexts in lock-step (coupled hierarchy, halo by device copies)
		from cslvr_b200.partition import partition_model, rank_context
		rms = partition_model(model, slabs)
		ctxs = [rank_context(model, rm, relative_tolerance=1e-9, maximum_iterations=25, krylov_rtol=1e-5, **kw) for rm in rms]
		us = [np.full(2 * rm.n_nodes, 3e-16) for rm in rms]
		for c in ctxs:
			c.set_params(maximum_iterations=1, krylov_maxit=4, krylov_error_on_nonconvergence=False)
		capi.newton_solve_group(ctxs, [x.copy() for x in us])
		for c in ctxs:
			c.set_params(maximum_iterations=25, krylov_maxit=10000, krylov_error_on_nonconvergence=True)
		t0 = time.perf_counter()
		st = capi.newton_solve_group(ctxs, us)
		t = time.perf_counter() - t0
		print("%-60s newton %2d  cg %5d  %.3f s  (%d slabs on one device)  conv %s" % (spec, st['iterations'], st['krylov_iterations'], t, slabs, st['converged']), flush=True)
		for c in ctxs:
			c.close()
		continue
	ctx = capi.context_from_model(model, relative_tolerance=1e-9, maximum_iterations=25, krylov_rtol=1e-5, **kw)
	u = np.full(2 * model.n_nodes, 3e-16)
	ctx.set_params(maximum_iterations=1, krylov_maxit=4, krylov_error_on_nonconvergence=False)
	ctx.newton_solve(u.copy())
	ctx.set_params(maximum_iterations=25, krylov_maxit=10000, krylov_error_on_nonconvergence=True)
	t0 = time.perf_counter()
	st = ctx.newton_solve(u)
	t = time.perf_counter() - t0
	print("%-60s newton %2d  cg %5d  %.3f s  %.3f ms/it  conv %s" % (spec or "(defaults)", st['iterations'], st['krylov_iterations'], t, 1e3 * st['t_krylov_ms'] / 1e3 / max(st['krylov_iterations'], 1), st['converged']), flush=True)
	ctx.close()
	for tok in spec.split():
		if tok.startswith('BP_'):
			os.environ.pop(tok.split('=')[0], None)
