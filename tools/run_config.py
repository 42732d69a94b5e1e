#!/usr/bin/env python
"""Solve one BASELINE.json configuration at a chosen size and print the solver statistics as one JSON line
(not a bench line: set-up is numpy, the solve is the library).  One GPU, or one rank per GPU under torchrun:
every rank builds the synthetic model, keeps its slab of ice columns (cslvr_b200.partition) and the ranks
solve together (peer-memory / NCCL halo exchange, coupled multigrid).

    python tools/run_config.py eismint 577 577 5                      # config 3, 9.99 M tets
    python tools/run_config.py greenland 816 816 5                    # config 4, 19.98 M tets
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \\
        tools/run_config.py greenland 816 816 5
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ.setdefault('CSLVR_B200_QUIET', '1')


def main():
	from cases import eismint_dome_model, greenland_like_model, product_context
	from cslvr_b200 import capi
	name = sys.argv[1]
	n = tuple(int(a) for a in sys.argv[2:5])
	world = int(os.environ.get('WORLD_SIZE', '1'))
	rank = int(os.environ.get('RANK', '0'))
	local = int(os.environ.get('LOCAL_RANK', '0'))
	kw = dict(use_pressure_bc=(name != 'eismint'), relative_tolerance=1e-9, relaxation_parameter=0.7,
	          maximum_iterations=60, krylov_rtol=1e-5)
	t0 = time.time()
	m = eismint_dome_model(n=n) if name == 'eismint' else greenland_like_model(n=n)
	t_model = time.time() - t0
	t0 = time.time()
	if world == 1:
		c = product_context(m, **kw)
		nd, own = 2 * m.n_nodes, slice(None)
		allmax = lambda *v: list(v)
		allsum = allmax
	else:
		import torch
		import torch.distributed as dist
		from cslvr_b200.partition import partition_model, rank_context
		torch.cuda.set_device(local)
		dev = torch.device('cuda', local)
		dist.init_process_group('nccl', device_id=dev)
		rm = partition_model(m, world, rank)
		c = rank_context(m, rm, device=local, **kw)
		ids = [capi.comm_unique_id() if rank == 0 else None]
		dist.broadcast_object_list(ids, src=0, device=dev)
		c.comm_init(ids[0], rank, world)
		nd, own = 2 * rm.n_nodes, slice(0, 2 * rm.n_owned)

		def allmax(*v):
			t = torch.tensor(list(v), device=dev, dtype=torch.float64)
			dist.all_reduce(t, op=dist.ReduceOp.MAX)
			return t.tolist()

		def allsum(*v):
			t = torch.tensor(list(v), device=dev, dtype=torch.float64)
			dist.all_reduce(t, op=dist.ReduceOp.SUM)
			return t.tolist()
	t_ctx = time.time() - t0
	u = np.full(nd, 3e-16)
	allmax(0.0)                                          # every rank is ready
	t0 = time.time()
	st = c.newton_solve(u)
	t_solve, = allmax(time.time() - t0)
	ms, = allmax(c.time_assemble(3, repeats=10, flush_l2=True))
	umax, = allmax(float(np.abs(u[own]).max()))
	# parity at full size (untimed): F and J x of the CUDA path at the converged velocity against the oracle's C twin on
	# this rank's arrays (owned rows are complete locally), and the oracle's own residual there
	parity = None
	if '--no-parity' not in sys.argv:
		import scipy.sparse as sp
		from oracle import bp_oracle as bo
		from oracle.bp_oracle_cwrap import COracle
		if world == 1:
			coords, cells, v2n, nn = m.mesh.coordinates(), m.mesh.cells(), np.asarray(m.vertex_to_node), m.n_nodes
			facets = (m.facet_cell, m.facet_local, m.ff)
			fS, fA, fb = m.vertex_field(m.S), m.vertex_field(m.A), m.vertex_field(m.beta)
			gid = np.arange(nn)
		else:
			coords, cells, v2n, nn = rm.coords, rm.cells, rm.vertex_to_node, rm.n_nodes
			facets, fS, fA, fb, gid = rm.facets, rm.fields['S'], rm.fields['A'], rm.fields['beta'], rm.node_global
		n2 = 2 * np.asarray(v2n)[cells]
		cd = np.concatenate([n2, n2 + 1], axis=1).astype(np.int32)
		P = bo.BPProblem(coords, cells, cd, 2 * nn, facets[0], facets[1], facets[2], fS, fA, fb, periodic=m.use_periodic,
		                 use_pressure_bc=kw['use_pressure_bc'], consts=dict(n=m.n, eps_reg=m.eps_reg, rho_i=m.rho_i, rhosw=m.rhosw, g=m.g))
		co = COracle(P, threads=max(1, (os.cpu_count() or 1) // world))
		# (F itself is ~1e-9 of its terms at the solution: it is compared at a perturbed state, the residual at the solution separately)
		up = u * (1.0 + 1e-2 * np.sin(0.37 * np.repeat(gid, 2) + np.tile([0.0, 0.7], nn)))
		r_sol = co.assemble(u, want_J=False)[0]
		Fo, Jo, _ = co.assemble(up)
		Fp, _ = c.assemble(up, want_F=True, want_J=True, J=False)
		x = np.sin(0.61 * np.repeat(gid, 2) + np.tile([0.3, 1.1], nn))
		Jx_o = sp.csr_matrix((Jo, co.indices, co.indptr), shape=(2 * nn, 2 * nn)) @ x
		Jx_p = c.spmv(x)
		scale = lambda a: max(float(np.abs(a).max()), 1e-300)
		eF, eJ = allmax(float(np.abs(Fp[own] - Fo[own]).max()), float(np.abs(Jx_p[own] - Jx_o[own]).max()))
		sF, sJ = allmax(scale(Fo[own]), scale(Jx_o[own]))
		a, = allsum(float(r_sol[own] @ r_sol[own]))
		parity = dict(F_rel=eF / sF, Jx_rel=eJ / sJ, oracle_residual_at_solution=float(np.sqrt(a)), oracle_residual_rel=float(np.sqrt(a)) / st['residual_norm0'],
		              checked="F and J x at the converged velocity perturbed by 1 % vs oracle/bp_oracle_c.c on every rank's arrays (owned rows); the oracle's residual at the converged velocity", tolerance=1e-12)
		parity['ok'] = bool(parity['F_rel'] < 1e-12 and parity['Jx_rel'] < 1e-12)
	if rank == 0:
		out = dict(config=name, mesh=n, n_gpus=world, n_tets=m.num_cells, n_dofs=2 * m.n_nodes, host_model_s=t_model, context_setup_s=t_ctx,
		           newton_solve_s=t_solve, converged=st['converged'], newton_iterations=st['iterations'],
		           krylov_iterations=st['krylov_iterations'], assemble_ms_in_solve=st['t_assemble_ms'], krylov_ms=st['t_krylov_ms'],
		           assembly_pass_ms=ms, assembly_mtets_s=m.num_cells / ms / 1e3, preconditioner=c.pc_label(),
		           gpu_launches=st['gpu_launches'], host_launches=st.get('host_launches'),
		           u_max=umax, residual0=st['residual_norm0'], residual=st['residual_norm'], parity=parity)
		print(json.dumps(out), flush=True)
	c.close()
	if world > 1:
		dist.barrier()
		dist.destroy_process_group()


if __name__ == '__main__':
	main()
