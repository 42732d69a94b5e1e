"""Pinning the oracle without reference goldens ("parity unpinned"):

* mesh counts quoted by the reference (simulations/verification/ismip_hom_a.py:5);
* a sympy re-derivation of the UFL forms of cslvr/momentumbp.py (action -> derivative
  -> derivative, exactly what ufl.derivative does symbolically) on single tetrahedra;
* residual = grad(action), Jacobian = grad(residual) by finite differences, globally;
* MomentumBP's classic residual == MomentumDukowiczBP's variational residual;
* the Jacobian is symmetric positive definite with beta > 0;
* the C twin equals the numpy oracle; the committed golden vectors still reproduce.
"""
import os

import numpy as np
import pytest

from cases import ismip_model, marine_model, oracle_problem, state, relerr
from oracle import bp_oracle as bo

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_boxmesh_counts_quoted_by_the_reference():
	c, t = bo.box_mesh((0, 0, 0), (1, 1, 1), 10, 10, 2)
	fc, fl = bo.exterior_facets(t)
	assert (c.shape[0], t.shape[0], (4 * t.shape[0] + fc.size) // 2) == (363, 1200, 2680)


def test_sizes_of_config_1():
	# 15x15x5 periodic -> N_t 6750, N_v 1536, N_d 2700, nnz 73 800; non-periodic 3072 / 78 424
	P = oracle_problem(ismip_model('A'))
	assert (P.cells.shape[0], P.coords.shape[0], P.n_dofs, P.pattern()[1].size) == (6750, 1536, 2700, 73800)
	Q = oracle_problem(ismip_model('A', periodic=False))
	assert (Q.n_dofs, Q.pattern()[1].size) == (3072, 78424)


def _sympy_element(coords, S, A, beta, u, basal_facet, press_facet, consts, uz=None, B_ring=None):
	"""F and J of ONE tetrahedron by symbolic differentiation of the action
	(cslvr/momentumbp.py:460-500), the cell integral with an exact degree-1 rule since A
	is taken constant here, facet integrals by exact polynomial integration."""
	import sympy as sp
	n, e0, rho, rsw, g = (sp.Float(consts[k], 30) for k in ('n', 'eps_reg', 'rho_i', 'rhosw', 'g'))
	U = sp.symbols('u0:8')
	X = sp.Matrix(coords)
	M = sp.Matrix([[1, X[a, 0], X[a, 1], X[a, 2]] for a in range(4)])
	Minv = M.inv()                       # column a = coefficients of lambda_a = c0 + c.x
	G = Minv[1:, :].T                    # grad lambda_a (rows)
	vol = abs(M.det()) / 6
	gu = [[sum(U[4 * c + a] * G[a, j] for a in range(4)) for j in range(3)] for c in range(2)]
	if uz is not None:
		# MomentumDukowiczStokesReduced.strain_rate_tensor (cslvr/momentumstokes.py:406-418):
		# eps_02 = (u_x,z + u_z,x)/2, eps_12 = (u_y,z + u_z,y)/2 with the frozen P1 u_z
		for c in range(2):
			gu[c][2] = gu[c][2] + sum(sp.Float(uz[a], 30) * G[a, c] for a in range(4))
	eps = sp.Matrix([[gu[0][0], (gu[0][1] + gu[1][0]) / 2, gu[0][2] / 2],
	                 [(gu[0][1] + gu[1][0]) / 2, gu[1][1], gu[1][2] / 2],
	                 [gu[0][2] / 2, gu[1][2] / 2, -gu[0][0] - gu[1][1]]])       # momentumbp.py:122-135
	epsdot = (eps * eps).trace() / 2                                           # momentum.py:245-251
	Vd = (2 * n) / (n + 1) * sp.Float(A, 30) ** (-1 / n) * (epsdot + e0) ** ((n + 1) / (2 * n))
	gS = [sum(S[a] * G[a, j] for a in range(4)) for j in range(2)]
	ubar = [sum(U[4 * c + a] for a in range(4)) / 4 for c in range(2)]          # cell mean of P1 = exact integral / vol
	action = vol * (Vd + rho * g * (ubar[0] * gS[0] + ubar[1] * gS[1]))
	s, t = sp.symbols('s t')
	for lf, kind in ((basal_facet, 'b'), (press_facet, 'p')):
		if lf is None:
			continue
		la = [a for a in range(4) if a != lf]
		lam = [1 - s - t, s, t]
		P0, P1, P2 = (X[la[k], :] for k in range(3))
		cr = (P1 - P0).cross(P2 - P0)
		area2 = sp.sqrt(cr.dot(cr))
		uq = [sum(lam[k] * U[4 * c + la[k]] for k in range(3)) for c in range(2)]
		integ = lambda f: sp.integrate(sp.integrate(f, (t, 0, 1 - s)), (s, 0, 1)) * area2
		if kind == 'b':
			bq = sum(lam[k] * beta[la[k]] for k in range(3))
			if uz is None:
				action += integ(bq * (uq[0] ** 2 + uq[1] ** 2) / 2)             # -Sl dGamma_b
			else:
				# -Sl_gnd dGamma_bg with u_z_b = (-B_ring - u_h.n_h) / n_z, momentumstokes.py:369-370
				nrm = cr / area2
				if nrm.dot((X[lf, :] - P0)) > 0:
					nrm = -nrm
				Brq = sum(lam[k] * sp.Float(B_ring[la[k]], 30) for k in range(3))
				uzb = (-Brq - uq[0] * nrm[0] - uq[1] * nrm[1]) / nrm[2]
				action += integ(bq * (uq[0] ** 2 + uq[1] ** 2 + uzb ** 2) / 2)
		else:
			nrm = cr / area2
			if nrm.dot((X[lf, :] - P0)) > 0:
				nrm = -nrm
			fw = sum(lam[k] * (rho * g * (S[la[k]] - X[la[k], 2]) - rsw * g * (-min(0.0, float(X[la[k], 2])))) for k in range(3))
			action -= integ(fw * (uq[0] * nrm[0] + uq[1] * nrm[1]))             # -Pb dGamma_bw
	F = [sp.diff(action, U[i]) for i in range(8)]
	J = [[sp.diff(F[i], U[j]) for j in range(8)] for i in range(8)]
	sub = {U[i]: sp.Float(u[i], 30) for i in range(8)}
	return (np.array([float(f.subs(sub)) for f in F]),
	        np.array([[float(J[i][j].subs(sub)) for j in range(8)] for i in range(8)]))


@pytest.mark.parametrize('seed', [0, 1])
def test_single_tet_against_symbolic_differentiation(seed):
	rng = np.random.default_rng(seed)
	coords = np.array([[0, 0, 0], [1200.0, 30, -40], [100, 900.0, 60], [50, -80, 350.0]]) + rng.normal(size=(4, 3)) * 20
	cells = np.array([[0, 1, 2, 3]])
	S = coords[:, 2] + 500 + rng.normal(size=4) * 30
	beta = 800 + rng.uniform(size=4) * 400
	A = 3.3e-17
	coords[:, 2] -= 300.0                                  # some vertices below sea level
	cd = bo.canonical_cell_dofs(cells, np.arange(4))
	fc, fl = np.array([0, 0]), np.array([3, 1])
	ff = np.array([bo.GAMMA_B_GND, bo.GAMMA_L_UDR])
	P = bo.BPProblem(coords, cells, cd, 8, fc, fl, ff, S, A, beta, periodic=False, use_pressure_bc=True)
	u8 = rng.normal(size=8) * 40
	u = np.empty(8); u[0::2] = u8[:4]; u[1::2] = u8[4:]    # dof = 2*node + c  <->  local 4c + a
	Fs, Js = _sympy_element(coords.tolist(), S.tolist(), A, beta.tolist(), u8.tolist(), 3, 1, P.c)
	perm = np.array([0, 2, 4, 6, 1, 3, 5, 7])               # local (4c+a) -> global dof
	F, J = P.residual(u), P.jacobian_matrix(u).toarray()
	assert relerr(F[perm], Fs) < 1e-12
	assert relerr(J[np.ix_(perm, perm)], Js) < 1e-12


@pytest.mark.parametrize('seed', [0, 1])
def test_reduced_stokes_single_tet_against_symbolic_differentiation(seed):
	"""MomentumDukowiczStokesReduced (cslvr/momentumstokes.py:300-400): the same symbolic
	pipeline with u_z in the strain-rate tensor and u_z_b in the sliding term."""
	from oracle.rs_oracle import RSProblem
	rng = np.random.default_rng(10 + seed)
	coords = np.array([[0, 0, 0], [1200.0, 30, -40], [100, 900.0, 60], [50, -80, 350.0]]) + rng.normal(size=(4, 3)) * 20
	cells = np.array([[0, 1, 2, 3]])
	S = coords[:, 2] + 500 + rng.normal(size=4) * 30
	beta = 800 + rng.uniform(size=4) * 400
	A = 3.3e-17
	coords[:, 2] -= 300.0
	uz = rng.normal(size=4) * 5
	Br = rng.normal(size=4) * 0.3
	cd = bo.canonical_cell_dofs(cells, np.arange(4))
	fc, fl = np.array([0, 0]), np.array([3, 1])
	ff = np.array([bo.GAMMA_B_GND, bo.GAMMA_L_UDR])
	P = RSProblem(coords, cells, cd, 8, fc, fl, ff, S, A, beta, periodic=False, use_pressure_bc=True, B=coords[:, 2], B_ring=Br)
	P.set_uz(uz)
	u8 = rng.normal(size=8) * 40
	u = np.empty(8); u[0::2] = u8[:4]; u[1::2] = u8[4:]
	Fs, Js = _sympy_element(coords.tolist(), S.tolist(), A, beta.tolist(), u8.tolist(), 3, 1, P.c, uz=uz.tolist(), B_ring=Br.tolist())
	perm = np.array([0, 2, 4, 6, 1, 3, 5, 7])
	F, J = P.residual(u), P.jacobian_matrix(u).toarray()
	assert relerr(F[perm], Fs) < 1e-12
	assert relerr(J[np.ix_(perm, perm)], Js) < 1e-12
	# a floating bed carries no sliding term in this model (dGamma_bg only)
	P2 = RSProblem(coords, cells, cd, 8, fc, fl, np.array([bo.GAMMA_B_FLT, bo.GAMMA_L_UDR]), S, A, beta, periodic=False, B=coords[:, 2])
	_, _, Jf = P2.facet_tensors(u)
	assert np.all(Jf == 0.0)


def test_reduced_stokes_residual_and_jacobian_are_derivatives_of_the_action():
	from cases import rs_problem
	P = rs_problem(marine_model(n=(5, 4, 3)))
	rng = np.random.default_rng(4)
	P.set_uz(rng.normal(size=P.coords.shape[0]) * 3)
	u = rng.normal(size=P.n_dofs) * 20
	F, J = P.residual(u), P.jacobian_matrix(u)
	for k in range(4):
		v = rng.normal(size=P.n_dofs)
		h = 1e-4
		dA = (P.action(u + h * v) - P.action(u - h * v)) / (2 * h)
		assert abs(dA - F @ v) < 1e-7 * max(abs(dA), abs(F @ v))
		dF = (P.residual(u + h * v) - P.residual(u - h * v)) / (2 * h)
		assert relerr(J @ v, dF) < 1e-6
	assert abs(J - J.T).max() < 1e-12 * abs(J).max()
	# with u_z = 0 the cell integrals are Blatter-Pattyn's; the sliding term is not (u_z_b)
	Q = rs_problem(ismip_model('A', n=(5, 4, 3)))
	Pb = oracle_problem(ismip_model('A', n=(5, 4, 3)))
	ub = state(Q.n_dofs, 2)
	Q.set_uz(np.zeros(Q.coords.shape[0]))
	fq, fb = Q.residual(ub), Pb.residual(ub)
	assert relerr(fq, fb) > 1e-6          # the bed of ISMIP-HOM A is bumpy: u_z_b couples u_x and u_y
	assert relerr(Q.cell_tensors(ub)[1], Pb.cell_tensors(ub)[1]) == 0.0


@pytest.mark.parametrize('builder', [lambda: ismip_model('A', n=(5, 4, 3)), lambda: marine_model(n=(5, 4, 3))])
def test_residual_and_jacobian_are_derivatives_of_the_action(builder):
	P = oracle_problem(builder())
	rng = np.random.default_rng(3)
	u = rng.normal(size=P.n_dofs) * 20
	F, J = P.residual(u), P.jacobian_matrix(u)
	for k in range(5):
		v = rng.normal(size=P.n_dofs)
		h = 1e-4
		dA = (P.action(u + h * v) - P.action(u - h * v)) / (2 * h)
		assert abs(dA - F @ v) < 1e-7 * max(abs(dA), abs(F @ v))
		dF = (P.residual(u + h * v) - P.residual(u - h * v)) / (2 * h)
		assert relerr(J @ v, dF) < 1e-6
	assert abs(J - J.T).max() < 1e-12 * abs(J).max()
	assert np.linalg.eigvalsh(J.toarray()).min() > 0                      # SPD (beta > 0)


def test_classic_and_dukowicz_residuals_coincide():
	for m in (ismip_model('C', L=20000.0, n=(6, 5, 3)), marine_model()):
		P = oracle_problem(m)
		u = state(P.n_dofs, 5)
		assert relerr(P.residual(u, form='classic'), P.residual(u, form='dukowicz')) < 1e-14


def test_c_twin_matches_numpy_oracle():
	from oracle.bp_oracle_cwrap import COracle
	for m in (ismip_model('A', n=(7, 6, 3)), marine_model()):
		P = oracle_problem(m)
		co = COracle(P)
		u = state(P.n_dofs, 6)
		F, J, _ = co.assemble(u)
		assert relerr(F, P.residual(u)) < 1e-13 and relerr(J, P.jacobian(u)) < 1e-13
	un, info_n = P.newton(rtol=1e-10, relax=0.7, maxit=80)
	uc, info_c = co.newton(rtol=1e-10, relax=0.7, maxit=80, krylov_rtol=1e-12)
	assert info_n['converged'] and info_c['converged'] and relerr(uc, un) < 1e-7


def test_ismip_hom_sanity_ranges():
	"""loose physical pins: docs/get_started.rst:102-104 (A, L = 8 km, beta = 1000: surface
	speed ~87-92 m/a for full Stokes) and ismip_hom_c_inverse.py:99-101 hints."""
	m = ismip_model('A', L=8000.0)
	P = oracle_problem(m)
	u, info = P.newton()
	assert info['converged'] and info['iterations'] <= 12
	top = np.isclose(P.coords[:, 2], P.S)
	ux = u[0::2][m.vertex_to_node[top]]
	assert 60.0 < ux.min() and ux.max() < 120.0


@pytest.mark.parametrize('name', ['ismip_a_6x5x3', 'marine_6x5x3'])
def test_golden_vectors_reproduce(name):
	"""tests/golden/*.npz were written by tests/golden/make_golden.py from the oracle;
	they pin it against regressions (they are NOT reference outputs)."""
	g = np.load(os.path.join(GOLD, name + '.npz'))
	from golden.make_golden import build
	P = oracle_problem(build(name))
	assert np.array_equal(P.pattern()[0], g['indptr']) and np.array_equal(P.pattern()[1], g['indices'])
	assert relerr(P.residual(g['u']), g['F']) < 1e-13
	assert relerr(P.jacobian(g['u']), g['J']) < 1e-13
	un, info = P.newton(rtol=1e-11, relax=float(g['relax']), maxit=80)
	assert relerr(un, g['u_newton']) < 1e-9


@pytest.mark.parametrize('name', ['rs_ismip_a_6x5x3', 'rs_marine_6x5x3'])
def test_reduced_stokes_golden_vectors_reproduce(name):
	"""tests/golden/rs_*.npz (make_golden.py::main_rs, oracle outputs): residual / Jacobian with a
	frozen random u_z and B_ring, and the home-rolled Newton solution with the u_z callback."""
	from cases import rs_problem
	from golden.make_golden import build
	g = np.load(os.path.join(GOLD, name + '.npz'))
	P = rs_problem(build(name[3:]), B_ring=g['B_ring'])
	P.set_uz(g['uz'])
	assert relerr(P.residual(g['u']), g['F']) < 1e-13
	assert relerr(P.jacobian(g['u']), g['J']) < 1e-13
	P.set_uz(np.zeros_like(g['uz']))
	un, info = P.newton_rs()
	assert info['iterations'] == int(g['newton_iterations']) and relerr(un, g['u_newton']) < 1e-9
	assert relerr(P.uz, g['uz_newton']) < 1e-8


def test_symmetric_elimination_of_dirichlet_rows_equals_dolfins_row_replacement():
	"""dolfin's NewtonSolver replaces the constrained rows of J by identity rows and sets b_i = u_i - g_i (SURVEY A-N:
	bc.apply, non-symmetric).  The library eliminates the constrained dofs symmetrically instead -- free rows lifted by
	J d, rows AND columns replaced -- so that CG keeps an SPD matrix (bp_set_dirichlet): the same Newton
	update, and the eliminated matrix is symmetric positive definite."""
	import scipy.sparse as sp
	import scipy.sparse.linalg as spl
	from cases import marine_model, oracle_problem
	m = marine_model()
	P = oracle_problem(m)
	rng = np.random.default_rng(11)
	u = rng.normal(size=P.n_dofs) * 20.0
	J = P.jacobian_matrix(u).tocsr()
	F = P.residual(u)
	bd = np.sort(rng.choice(P.n_dofs, size=P.n_dofs // 9, replace=False))
	g = rng.normal(size=bd.size) * 5.0
	free = np.ones(P.n_dofs); free[bd] = 0.0
	d = np.zeros(P.n_dofs); d[bd] = u[bd] - g
	# dolfin: rows only
	Jn = (sp.diags(free) @ J + sp.diags(1.0 - free)).tocsc()
	bn = F.copy(); bn[bd] = d[bd]
	dx_n = spl.spsolve(Jn, bn)
	# the library: b_free -= J d, rows and columns -> identity
	bs = F - J @ d; bs[bd] = d[bd]
	Js = (sp.diags(free) @ J @ sp.diags(free) + sp.diags(1.0 - free)).tocsc()
	dx_s = spl.spsolve(Js, bs)
	# (J is badly scaled at a random state: the two LU solutions agree to its conditioning, the residual of dolfin's system is exact)
	assert np.abs(Jn @ dx_s - bn).max() < 1e-9 * np.abs(bn).max()
	assert np.abs(dx_s - dx_n).max() < 1e-4 * np.abs(dx_n).max()
	assert np.allclose(dx_s[bd], u[bd] - g)
	assert abs(Js - Js.T).max() < 1e-12 * abs(Js).max()
	assert spl.eigsh(Js, k=1, which='SA', return_eigenvectors=False, tol=1e-6)[0] > 0.0
