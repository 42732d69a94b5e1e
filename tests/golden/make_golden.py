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


if __name__ == '__main__':
	main()
