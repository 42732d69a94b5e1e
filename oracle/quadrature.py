"""Quadrature rules the reference's form compiler would pick (oracle copy).

Test infrastructure (see ``oracle/__init__.py``).  The reference never names a
rule: FFC 2017.2 picks it from UFL's estimated polynomial degree and FIAT's
default scheme table (third-party code that is NOT under
/root/reference -- recalled, not verifiable here).  The scripts only ever set
``parameters['form_compiler']['quadrature_degree']``
(simulations/ismip_hom/ismip_hom_a.py:10).

All rules are returned in barycentric coordinates (one row per point, rows sum
to one) with weights summing to ONE, i.e. the physical weight is ``w * |T|``.
Every rule here is checked for exactness in tests/test_oracle_quadrature.py.
"""
import itertools
import numpy as np


def _perm(p):
	return sorted(set(itertools.permutations(p)), reverse=True)


def _s31(a):
	return _perm((a, a, a, 1.0 - 3.0 * a))


def _s22(b):
	return _perm((b, b, 0.5 - b, 0.5 - b))


def _s211(a, b):
	return _perm((a, a, b, 1.0 - 2.0 * a - b))


def _gauss_jacobi_tet(m):
	"""Collapsed Gauss-Jacobi rule with m points per axis (FIAT's fallback)."""
	from scipy.special import roots_jacobi
	x0, w0 = roots_jacobi(m, 0, 0)
	x1, w1 = roots_jacobi(m, 1, 0)
	x2, w2 = roots_jacobi(m, 2, 0)
	pts, wts = [], []
	for i in range(m):
		for j in range(m):
			for k in range(m):
				# Duffy collapse of [-1,1]^3 onto the reference tetrahedron
				z = 0.5 * (1.0 + x2[k])
				y = 0.5 * (1.0 + x1[j]) * (1.0 - z)
				x = 0.5 * (1.0 + x0[i]) * (1.0 - z - y)
				pts.append((1.0 - x - y - z, x, y, z))
				wts.append(w0[i] * w1[j] * w2[k])
	pts = np.array(pts)
	wts = np.array(wts)
	return pts, wts / wts.sum()


def tetrahedron_rule(degree):
	"""(points[nq,4] barycentric, weights[nq] summing to 1) for ``degree``."""
	if degree <= 1:
		pts = [(0.25, 0.25, 0.25, 0.25)]
		wts = [1.0]
	elif degree == 2:
		a = 0.138196601125010515179541316563436
		pts = _s31(a)
		wts = [0.25] * 4
	elif degree == 3:
		# 5 points, one negative weight
		pts = [(0.25, 0.25, 0.25, 0.25)] + _s31(1.0 / 6.0)
		wts = [-0.8] + [0.45] * 4
	elif degree == 4:
		# Keast, 14 points (edge mid-points + two vertex orbits)
		a1, w1 = 0.10052676522520447969, 0.08858982474298071043
		a2, w2 = 0.31437287349319219275, 0.13283874668559071814
		w3 = 2.0 / 105.0
		pts = _s22(0.0) + _s31(a1) + _s31(a2)
		wts = [w3] * 6 + [w1] * 4 + [w2] * 4
	elif degree == 5:
		# Keast, 15 points: centroid, two vertex orbits (a=1/3, 1/11), one edge orbit
		w0 = 0.03028367809708917580637697 * 6.0
		w1 = (27.0 / 4480.0) * 6.0
		w2 = 0.01164524908602896941089361 * 6.0
		w3 = 0.01094914156138645934564302 * 6.0
		b = 0.06655015357366429823988046
		pts = [(0.25, 0.25, 0.25, 0.25)] + _s31(1.0 / 3.0) + _s31(1.0 / 11.0) \
		      + _s22(b)
		wts = [w0] + [w1] * 4 + [w2] * 4 + [w3] * 6
	elif degree == 6:
		# Keast, 24 points: three vertex orbits + one 12-point orbit
		a1, w1 = 0.21460287125915202929, 0.03992275025816749210
		a2, w2 = 0.04067395853461135312, 0.01007721105532064295
		a3, w3 = 0.32233789014227551034, 0.05535718154365472210
		a, b, w4 = 0.06366100187501752530, 0.26967233145831580803, 27.0 / 560.0
		pts = _s31(a1) + _s31(a2) + _s31(a3) + _s211(a, b)
		wts = [w1] * 4 + [w2] * 4 + [w3] * 4 + [w4] * 12
	else:
		return _gauss_jacobi_tet((degree + 2) // 2)
	pts = np.array(pts, dtype=np.float64)
	wts = np.array(wts, dtype=np.float64)
	return pts, wts


def triangle_rule(degree):
	"""(points[nq,3] barycentric, weights[nq] summing to 1) for ``degree``."""
	if degree <= 1:
		pts = [(1.0 / 3.0,) * 3]
		wts = [1.0]
	elif degree == 2:
		pts = sorted(set(itertools.permutations((2.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0))))
		wts = [1.0 / 3.0] * 3
	elif degree == 3:
		# Strang-Fix, 6 points
		a, b = 0.659027622374092, 0.231933368553031
		c = 1.0 - a - b
		pts = sorted(set(itertools.permutations((a, b, c))))
		wts = [1.0 / 6.0] * 6
	else:
		# Gauss-Jacobi collapsed rule, generous enough for anything on this path
		from scipy.special import roots_jacobi
		m = (degree + 2) // 2
		x0, w0 = roots_jacobi(m, 0, 0)
		x1, w1 = roots_jacobi(m, 1, 0)
		pts, wts = [], []
		for i in range(m):
			for j in range(m):
				y = 0.5 * (1.0 + x1[j])
				x = 0.5 * (1.0 + x0[i]) * (1.0 - y)
				pts.append((1.0 - x - y, x, y))
				wts.append(w0[i] * w1[j])
		wts = list(np.array(wts) / np.sum(wts))
	return np.array(pts, dtype=np.float64), np.array(wts, dtype=np.float64)
