"""The partitioned path on ONE GPU: k contexts solved in lock-step (halo by device
copies, reductions summed on the device) must reproduce the single-context result
(k-rank = 1-rank to ~1e-13 in F, same Newton counts +-1)."""
import numpy as np
import pytest

from cases import ismip_model, marine_model, oracle_problem, product_context, state, relerr
from cslvr_b200 import capi
from cslvr_b200.partition import partition_model, rank_context, local_vector, gather_owned

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('builder,nr', [(lambda: ismip_model('A', n=(12, 6, 4)), 2),
                                        (lambda: ismip_model('C', L=20000.0, n=(9, 7, 3)), 3),
                                        (lambda: marine_model(n=(10, 5, 3)), 4)])
def test_group_assembly_and_newton_match_single_context(builder, nr):
	m = builder()
	relax = 1.0 if m.use_periodic else 0.7
	kw = dict(relative_tolerance=1e-10, maximum_iterations=60, krylov_rtol=1e-12, relaxation_parameter=relax)
	single = product_context(m, **kw)
	rms = partition_model(m, nr)
	ctxs = [rank_context(m, rm, **kw) for rm in rms]
	u = state(2 * m.n_nodes, 21)
	F1, _ = single.assemble(u, want_F=True, want_J=False)
	Fs = capi.assemble_group(ctxs, [local_vector(rm, u) for rm in rms])
	Fg = gather_owned(rms, Fs, m.n_nodes)
	assert relerr(Fg, F1) < 1e-13
	assert relerr(Fg, oracle_problem(m).residual(u)) < 1e-12
	u1 = np.full(2 * m.n_nodes, 3e-16)
	s1 = single.newton_solve(u1)
	us = [np.full(2 * rm.n_nodes, 3e-16) for rm in rms]
	sg = capi.newton_solve_group(ctxs, us)
	ug = gather_owned(rms, us, m.n_nodes)
	assert s1['converged'] and sg['converged']
	assert abs(s1['iterations'] - sg['iterations']) <= 1
	assert relerr(ug, u1) < 1e-8
	# ghost copies agree with their owners after the solve
	for rm, ul in zip(rms, us):
		assert relerr(ul, local_vector(rm, ug)) < 1e-8
	for c in ctxs + [single]:
		c.close()


def test_coupled_multigrid_keeps_the_single_context_iteration_count():
	"""The multigrid hierarchy of a partitioned mesh keeps the couplings across the slab interfaces on
	every level (ghost nodes + halo exchange per level, global dense coarsest solve): the CG iteration
	count stays near the single-context one, where the rank-local hierarchy (BP_MG_COUPLED=0) needs
	clearly more; the solution is the same either way."""
	import os
	m = ismip_model('C', L=20000.0, n=(64, 24, 4))
	kw = dict(preconditioner='mg', relative_tolerance=1e-11, maximum_iterations=25, krylov_rtol=1e-12)
	single = product_context(m, **kw)
	u1 = np.full(2 * m.n_nodes, 3e-16)
	s1 = single.newton_solve(u1)
	single.close()
	res = {}
	for mode in ('1', '0'):
		os.environ['BP_MG_COUPLED'] = mode
		try:
			rms = partition_model(m, 4)
			ctxs = [rank_context(m, rm, **kw) for rm in rms]
			us = [np.full(2 * rm.n_nodes, 3e-16) for rm in rms]
			sg = capi.newton_solve_group(ctxs, us)
			ug = gather_owned(rms, us, m.n_nodes)
			for c in ctxs:
				c.close()
		finally:
			os.environ.pop('BP_MG_COUPLED', None)
		assert sg['converged'] and abs(sg['iterations'] - s1['iterations']) <= 1
		assert relerr(ug, u1) < 1e-8
		res[mode] = sg['krylov_iterations']
	assert res['1'] <= 1.35 * s1['krylov_iterations'] + 8, (res, s1['krylov_iterations'])
	assert res['0'] >= res['1'], (res, s1['krylov_iterations'])
