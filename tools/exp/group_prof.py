"""Kernel-time breakdown of the Krylov loop, one context vs an in-process group of slabs on ONE GPU (run under
ncu --metrics gpu__time_duration.sum for the launch list).   python tools/exp/group_prof.py nslabs [nxy] [krylov_maxit]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nxy = int(sys.argv[2]) if len(sys.argv) > 2 else 577
kmax = int(sys.argv[3]) if len(sys.argv) > 3 else 24
os.environ['BENCH_NXY'] = str(nxy)
import bench
from cslvr_b200 import capi
kw = dict(relative_tolerance=1e-9, maximum_iterations=1, relaxation_parameter=1.0, krylov_rtol=1e-5, krylov_maxit=kmax, krylov_error_on_nonconvergence=False)
if ns == 1:
	m = bench.build_model(1, nxy=nxy)
	ctx = capi.context_from_model(m, verbose=True)
	ctx.set_params(**kw)
	for _ in range(2):
		u = np.full(2 * m.n_nodes, 3e-16)
		st = ctx.newton_solve(u)
		print(st['krylov_iterations'], st['t_krylov_ms'], st['gpu_launches'], st['host_launches'])
else:
	ctxs, us = [], []
	for r in range(ns):
		s = bench.slab_of_rank(r, ns)
		c = capi.BPContext(s.coords, s.cells, 2 * s.n_nodes, s.facets, dict(bench.CONSTS, periodic=True, use_pressure_bc=True, verbose=(r == 0)),
		                   vertex_to_node=s.vertex_to_node, partition=s.partition, device=0)
		for f, k in ((capi.FIELD_S, 'S'), (capi.FIELD_A, 'A'), (capi.FIELD_BETA, 'beta')):
			c.set_field(f, s.fields[k])
		c.set_params(**kw)
		ctxs.append(c); us.append(np.full(2 * s.n_nodes, 3e-16))
	for _ in range(2):
		for u in us: u[:] = 3e-16
		st = capi.newton_solve_group(ctxs, us)
		print(st['krylov_iterations'], st['t_krylov_ms'], st['gpu_launches'], st['host_launches'])
