"""MMS convergence of the ORACLE (reference: cslvr/verification.py:329-340 with the
compensatory forcing of cslvr/momentum.py:261-300): on three refinements of a periodic box with an undulating
bed the P1 Galerkin solution of the modified BP problem converges to the manufactured velocity with order ~2 in
L2.  This pins the oracle's forms, quadrature, dof map and Newton loop against something nobody in this repo
wrote down as code: an analytic solution."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

import mms
from cases import oracle_problem


def test_oracle_converges_to_the_manufactured_solution_with_order_two():
	errs = []
	for n, nz in ((4, 2), (8, 4), (16, 8)):
		m = mms.mms_model(n, nz)
		P = oracle_problem(m)
		b = mms.forcing_vector(P.coords, P.cells, P.cell_dofs, P.n_dofs, P.fc, P.fl, P.ff, P.beta)
		ip, ix = P.pattern()
		u = np.full(P.n_dofs, 1.0)
		for it in range(30):
			F = P.residual(u) - b
			if np.linalg.norm(F) < 1e-10 * np.linalg.norm(b):
				break
			J = sp.csr_matrix((P.jacobian(u), ix, ip), shape=(P.n_dofs, P.n_dofs))
			u -= spl.spsolve(J.tocsc(), F)
		assert it < 29
		e, nrm = mms.l2_error(P.coords, P.cells, P.cell_dofs, u)
		errs.append(e / nrm)
	orders = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
	assert errs[-1] < 7e-3 and min(orders) > 1.7, (errs, orders)
