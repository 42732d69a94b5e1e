"""Size-independent properties at BASELINE.json's full single-GPU size (configs[1]: ISMIP-HOM C,
260 x 260 x 5 = 2.03 M tets), where the oracle no longer finishes in seconds:

* the two assembly mappings (row gather / atomic scatter) agree to 1e-12;
* J is symmetric:  x.(J y) == y.(J x);
* J is the derivative of F:  J v == (F(u + h v) - F(u - h v)) / 2h;
* partition of unity: the cell part of J annihilates rigid translations, so
  1_x . J 1_x == int beta dGamma_b (computed here from the mesh in numpy);
* the Picard operator of linear=True is linear in its argument and SPD along random directions;
* Newton converges to rtol 1e-9 and the solution is reproduced by a 4-slab partitioned solve;
* F, the CSR pattern and EVERY entry of J against the oracle's C twin (oracle/bp_oracle_c.c), which
  assembles 2 M tets in a fraction of a second: 1e-12 in the max norm and entry by entry.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from cases import product_context, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def big():
	from cases import ismip_model
	m = ismip_model('C', L=20000.0, n=(260, 260, 5))
	assert m.num_cells == 2028000
	rng = np.random.default_rng(99)
	x, y, z = m.mesh.coordinates()[m.node_vertex].T
	# a smooth sheared flow plus noise: every strain-rate component is exercised
	u = np.empty(2 * m.n_nodes)
	u[0::2] = 20.0 + 15.0 * (z - z.min()) / 1000.0 + 3.0 * np.sin(2 * np.pi * x / 20000.0) + rng.normal(size=m.n_nodes) * 0.05
	u[1::2] = 2.0 * np.cos(2 * np.pi * y / 20000.0) + rng.normal(size=m.n_nodes) * 0.05
	return m, u, rng


def test_full_size_parity_with_the_c_twin(big):
	from cases import oracle_problem, relerr_entrywise
	from oracle.bp_oracle_cwrap import COracle
	m, u, rng = big
	co = COracle(oracle_problem(m))
	Fo, Jo, _ = co.assemble(u)
	c = product_context(m)
	ip, ix = c.csr_pattern()
	assert np.array_equal(ip, co.indptr) and np.array_equal(ix, co.indices)          # bit-exact sparsity
	F, J = c.assemble(u, want_F=True, want_J=True)
	c.close()
	assert relerr(F, Fo) < 1e-12 and relerr(J, Jo) < 1e-12
	# entry by entry: an entry is a sum of up to 24 element contributions of the size of the largest entries, so
	# its own relative error is eps * max|J| / |entry|: 1e-11 above 1e-4 max|J|, 1e-5 down to 1e-10 max|J|
	assert relerr_entrywise(J, Jo, floor=1e-4) < 1e-11 and relerr_entrywise(J, Jo, floor=1e-10) < 1e-5
	assert relerr_entrywise(F, Fo, floor=1e-4) < 1e-10


def test_gather_and_scatter_assembly_agree_at_full_size(big):
	m, u, rng = big
	c = product_context(m, assembly='gather')
	Fg, _ = c.assemble(u, want_F=True, want_J=True, J=False)
	x = rng.normal(size=u.size)
	yg = c.spmv(x)
	c.close()
	c = product_context(m, assembly='scatter')
	Fs, _ = c.assemble(u, want_F=True, want_J=True, J=False)
	ys = c.spmv(x)
	c.close()
	assert relerr(Fg, Fs) < 1e-12 and relerr(yg, ys) < 1e-12


def test_jacobian_symmetry_derivative_and_partition_of_unity(big):
	m, u, rng = big
	c = product_context(m)
	F, _ = c.assemble(u, want_F=True, want_J=True, J=False)
	x, y = rng.normal(size=u.size), rng.normal(size=u.size)
	Jx, Jy = c.spmv(x), c.spmv(y)
	assert abs(x @ Jy - y @ Jx) < 1e-12 * max(abs(x @ Jy), abs(np.abs(x) @ np.abs(Jy)))
	assert x @ Jx > 0 and y @ Jy > 0                                  # SPD along random directions
	# rigid translations: only the sliding term sees them
	X = m.mesh.coordinates()
	cells = m.mesh.cells()
	basal = np.isin(m.ff, (3, 5))
	loc = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
	fv = cells[m.facet_cell[basal]][np.arange(basal.sum())[:, None], loc[m.facet_local[basal]]]
	P = X[fv]
	area = 0.5 * np.linalg.norm(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), axis=1)
	beta_v = m.vertex_field(m.beta)
	int_beta = (area * beta_v[fv].mean(axis=1)).sum()                 # exact for P1 beta
	ex = np.zeros(u.size); ex[0::2] = 1.0
	ey = np.zeros(u.size); ey[1::2] = 1.0
	Jex, Jey = c.spmv(ex), c.spmv(ey)
	assert abs(ex @ Jex - int_beta) < 1e-9 * int_beta and abs(ey @ Jey - int_beta) < 1e-9 * int_beta
	assert abs(ex @ Jey) < 1e-9 * int_beta
	# directional derivative of the residual
	v = rng.normal(size=u.size)
	h = 1e-4
	Fp, _ = c.assemble(u + h * v, want_F=True, want_J=False)
	Fm, _ = c.assemble(u - h * v, want_F=True, want_J=False)
	c.assemble(u, want_F=False, want_J=True, J=False)
	assert relerr(c.spmv(v), (Fp - Fm) / (2 * h)) < 1e-6
	c.close()


def test_picard_operator_is_linear_at_full_size(big):
	m, u, rng = big
	c = product_context(m, krylov_rtol=1e-10)
	out = c.linear_solve(u)                     # K(u) w = -f
	# the resident matrix is now K(u): linear, symmetric
	x, y = rng.normal(size=u.size), rng.normal(size=u.size)
	a, b = 0.7, -1.3
	assert relerr(c.spmv(a * x + b * y), a * c.spmv(x) + b * c.spmv(y)) < 1e-12
	assert abs(x @ c.spmv(y) - y @ c.spmv(x)) < 1e-11 * abs(np.abs(x) @ np.abs(c.spmv(y)))
	assert np.all(np.isfinite(out)) and c.last_linear_iterations > 0
	c.close()


def test_newton_at_full_size_and_partitioned(big):
	m, u, rng = big
	c = product_context(m, relative_tolerance=1e-11, maximum_iterations=25, krylov_rtol=1e-11)
	un = np.full(u.size, 3e-16)
	st = c.newton_solve(un)
	assert st['converged'] and st['iterations'] <= 12
	assert st['residuals'][-1] < 1e-11 * st['residuals'][0]
	F, _ = c.assemble(un, want_F=True, want_J=False)
	assert np.linalg.norm(F) < 1e-10 * st['residuals'][0]
	c.close()
	# the same problem as four x-slabs on this GPU (halo by device copies): same velocity
	from cslvr_b200 import capi
	from cslvr_b200.partition import partition_model, rank_context, gather_owned
	rms = partition_model(m, 4)
	ctxs = [rank_context(m, rm, relative_tolerance=1e-11, maximum_iterations=25, krylov_rtol=1e-11) for rm in rms]
	us = [np.full(2 * rm.n_nodes, 3e-16) for rm in rms]
	sg = capi.newton_solve_group(ctxs, us)
	ug = gather_owned(rms, us, m.n_nodes)
	assert sg['converged'] and abs(sg['iterations'] - st['iterations']) <= 1
	assert relerr(ug, un) < 1e-8           # north_star's velocity bar (both sides solved to 1e-11)
	for cc in ctxs:
		cc.close()
