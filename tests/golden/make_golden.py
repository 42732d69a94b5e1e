"""Writes tests/golden/*.npz from the ORACLE (python tests/golden/make_golden.py).

PARITY UNPINNED: the reference cannot run here and ships no golden vectors for this
path, so these files are oracle outputs -- they pin the oracle and the
CUDA path to each other and guard against regressions; they are not reference data."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def build(name):
	from cases import ismip_model, marine_model
	if name == 'ismip_a_6x5x3':
		return ismip_model('A', n=(6, 5, 3))
	if name == 'marine_6x5x3':
		return marine_model(n=(6, 5, 3))
	raise KeyError(name)


def main():
	from cases import oracle_problem, state
	os.environ['CSLVR_B200_QUIET'] = '1'
	for name, relax in (('ismip_a_6x5x3', 1.0), ('marine_6x5x3', 0.7)):
		P = oracle_problem(build(name))
		u = state(P.n_dofs, 99)
		ip, ix = P.pattern()
		un, info = P.newton(rtol=1e-11, relax=relax, maxit=80)
		assert info['converged']
		np.savez_compressed(os.path.join(HERE, name + '.npz'), u=u, F=P.residual(u), J=P.jacobian(u),
		                    indptr=ip, indices=ix, u_newton=un, relax=relax,
		                    newton_iterations=info['iterations'])
		print(name, P.n_dofs, ix.size, info['iterations'])


def main_rs():
	"""MomentumDukowiczStokesReduced (oracle/rs_oracle.py): residual and Jacobian at a random state with a
	random frozen u_z and B_ring, and the home-rolled Newton solution with the u_z callback."""
	from cases import rs_problem, state
	os.environ['CSLVR_B200_QUIET'] = '1'
	for base in ('ismip_a_6x5x3', 'marine_6x5x3'):
		m = build(base)
		rng = np.random.default_rng(7)
		Br = rng.normal(size=m.num_vertices) * 0.2
		uz = rng.normal(size=m.n_nodes)[m.vertex_to_node] * 3.0
		P = rs_problem(m, B_ring=Br)
		P.set_uz(uz)
		u = state(P.n_dofs, 98)
		F, J = P.residual(u), P.jacobian(u)
		P.set_uz(np.zeros(m.num_vertices))
		un, info = P.newton_rs(rtol=1e-9, atol=1e-6, maxit=50, relax=0.8)
		assert info['converged']
		np.savez_compressed(os.path.join(HERE, 'rs_' + base + '.npz'), u=u, uz=uz, B_ring=Br, F=F, J=J, u_newton=un,
		                    uz_newton=P.uz, newton_iterations=info['iterations'], residuals=np.array(info['residuals']))
		print('rs_' + base, P.n_dofs, info['iterations'])


if __name__ == '__main__':
	main()
	main_rs()
