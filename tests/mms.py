"""Method of manufactured solutions for the Blatter-Pattyn branch -- the external truth the oracle and the
CUDA path are measured against (the reference ships no goldens; its own check is MMS:
cslvr/verification.py:329-340, 452-476 + cslvr/momentum.py:261-300, driven by
simulations/verification/ismip_hom_a.py:166-167).

A smooth periodic velocity u*(x, y, z) is written down in sympy.  The reference turns it into "compensatory
forcing terms" that are subtracted from the residual (momentum.py:293-301); here the same modified problem
is posed in its weak form, which needs no hand-derived stress divergence:

    find u_h :  F(u_h) - b = 0,      b_i = a(u*; phi_i)  (the BP form at the ANALYTIC u*, by degree-8 quadrature)

with a(u; v) = int 2 eta(u) epsdot_BP(u) : grad v dOmega + int beta u . v dGamma_b (+ the gravity term, which is
zero because the surface is flat).  u* solves the continuous modified problem, F is monotone, so the P1 Galerkin
solution converges to u* with order 2 in L2 and order 1 in H1.  Everything here is numpy + sympy and touches
neither the oracle's nor the product's element code (only the mesh arrays and the P1 basis gradients)."""
import numpy as np
import sympy as sp

from oracle.quadrature import tetrahedron_rule, triangle_rule

L = 40000.0
A0, N_GLEN, EPS_REG, BETA0 = 1e-16, 3.0, 1e-15, 1000.0


def S_fun(x, y, z=None):
	return 0.0 * x


def B_fun(x, y, z=None):
	return -1000.0 + 300.0 * np.sin(2 * np.pi * x / L) * np.sin(2 * np.pi * y / L)


def beta_fun(x, y, z=None):
	return BETA0 * (1.0 + 0.5 * np.cos(2 * np.pi * x / L))


def _exact():
	x, y, z = sp.symbols('x y z', real=True)
	k = 2 * sp.pi / L
	B = -1000 + 300 * sp.sin(k * x) * sp.sin(k * y)
	s = (z - B) / (0 - B)                                  # 0 at the bed, 1 at the surface
	ux = 40 + 15 * sp.sin(k * x) * sp.cos(k * y) + 60 * s ** 2 * (1 + sp.Rational(1, 4) * sp.cos(k * y))
	uy = 8 * sp.cos(k * x) * sp.sin(k * y) * (sp.Rational(1, 2) + s)
	fn = {}
	for nm, e in (('ux', ux), ('uy', uy)):
		fn[nm] = sp.lambdify((x, y, z), e, 'numpy')
		for d, v in (('x', x), ('y', y), ('z', z)):
			fn[nm + '_' + d] = sp.lambdify((x, y, z), sp.diff(e, v), 'numpy')
	return fn


EXACT = _exact()


def exact_velocity(X):
	return EXACT['ux'](X[:, 0], X[:, 1], X[:, 2]), EXACT['uy'](X[:, 0], X[:, 1], X[:, 2])


def _p1_gradients(P):
	"""grad lambda_a [nt, 4, 3] and volumes [nt] of the tets P [nt, 4, 3]."""
	M = np.concatenate([np.ones(P.shape[:2] + (1,)), P], axis=2)           # rows (1, x, y, z)
	Minv = np.linalg.inv(M)                                                 # columns = coefficients of lambda_a
	G = np.transpose(Minv[:, 1:4, :], (0, 2, 1))
	vol = np.abs(np.linalg.det(M)) / 6.0
	return G, vol


def forcing_vector(coords, cells, cell_dofs, n_dofs, facet_cell, facet_local, facet_marker, beta_vertex, degree=8):
	"""b_i = a(u*; phi_i): cell integral of 2 eta(u*) d(epsdot)(u*; phi_i) and the sliding integral of beta_h u* . phi_i
	(beta_h: the P1 interpolant the discrete problem uses)."""
	P = coords[cells]
	G, vol = _p1_gradients(P)
	pts, wts = tetrahedron_rule(degree)
	b = np.zeros(n_dofs)
	Fx = np.zeros((cells.shape[0], 4))
	Fy = np.zeros((cells.shape[0], 4))
	for lam, w in zip(np.asarray(pts), np.asarray(wts)):
		Xq = np.einsum('a,tad->td', lam, P)
		a = (Xq[:, 0], Xq[:, 1], Xq[:, 2])
		uxx, uxy, uxz = EXACT['ux_x'](*a), EXACT['ux_y'](*a), EXACT['ux_z'](*a)
		uyx, uyy, uyz = EXACT['uy_x'](*a), EXACT['uy_y'](*a), EXACT['uy_z'](*a)
		ed = uxx ** 2 + uyy ** 2 + uxx * uyy + 0.25 * (uxy + uyx) ** 2 + 0.25 * uxz ** 2 + 0.25 * uyz ** 2 + EPS_REG
		two_eta = A0 ** (-1.0 / N_GLEN) * ed ** ((1.0 - N_GLEN) / (2.0 * N_GLEN))
		sh = 0.5 * (uxy + uyx)
		cx = np.stack([2 * uxx + uyy, sh, 0.5 * uxz], axis=1)                 # d(epsdot)[(0, a)] = cx . grad lambda_a
		cy = np.stack([sh, 2 * uyy + uxx, 0.5 * uyz], axis=1)
		Fx += (w * vol * two_eta)[:, None] * np.einsum('td,tad->ta', cx, G)
		Fy += (w * vol * two_eta)[:, None] * np.einsum('td,tad->ta', cy, G)
	np.add.at(b, cell_dofs[:, 0:4].ravel(), Fx.ravel())
	np.add.at(b, cell_dofs[:, 4:8].ravel(), Fy.ravel())
	# sliding on dGamma_b (markers 3, 5)
	basal = np.isin(facet_marker, (3, 5))
	fc, fl = facet_cell[basal], facet_local[basal]
	loc = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])[fl]
	fv = cells[fc][np.arange(fc.size)[:, None], loc]
	Pf = coords[fv]
	area = 0.5 * np.linalg.norm(np.cross(Pf[:, 1] - Pf[:, 0], Pf[:, 2] - Pf[:, 0]), axis=1)
	tp, tw = triangle_rule(6)
	dx = cell_dofs[fc][np.arange(fc.size)[:, None], loc]
	dy = cell_dofs[fc][np.arange(fc.size)[:, None], loc + 4]
	for lam, w in zip(np.asarray(tp), np.asarray(tw)):
		Xq = np.einsum('a,fad->fd', lam, Pf)
		bq = beta_vertex[fv] @ lam
		ux, uy = EXACT['ux'](Xq[:, 0], Xq[:, 1], Xq[:, 2]), EXACT['uy'](Xq[:, 0], Xq[:, 1], Xq[:, 2])
		np.add.at(b, dx.ravel(), ((w * area * bq * ux)[:, None] * lam[None, :]).ravel())
		np.add.at(b, dy.ravel(), ((w * area * bq * uy)[:, None] * lam[None, :]).ravel())
	return b


def l2_error(coords, cells, cell_dofs, u, degree=6):
	"""(||u_h - u*||_L2, ||u*||_L2) over the mesh."""
	P = coords[cells]
	_, vol = _p1_gradients(P)
	pts, wts = tetrahedron_rule(degree)
	e2 = n2 = 0.0
	ux_h, uy_h = u[cell_dofs[:, 0:4]], u[cell_dofs[:, 4:8]]
	for lam, w in zip(np.asarray(pts), np.asarray(wts)):
		Xq = np.einsum('a,tad->td', lam, P)
		ex, ey = EXACT['ux'](Xq[:, 0], Xq[:, 1], Xq[:, 2]), EXACT['uy'](Xq[:, 0], Xq[:, 1], Xq[:, 2])
		e2 += w * (vol * ((ux_h @ lam - ex) ** 2 + (uy_h @ lam - ey) ** 2)).sum()
		n2 += w * (vol * (ex ** 2 + ey ** 2)).sum()
	return np.sqrt(e2), np.sqrt(n2)


def mms_model(n, nz):
	from cslvr_b200 import BoxMesh, Point, D3Model
	mesh = BoxMesh(Point(0, 0, 0), Point(L, L, 1), n, n, nz)
	model = D3Model(mesh, use_periodic=True)
	model.calculate_boundaries()
	model.deform_mesh_to_geometry(S_fun, B_fun)
	model.init_beta(beta_fun)
	model.init_A(A0)
	return model
