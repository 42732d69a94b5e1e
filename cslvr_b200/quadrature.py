"""Tetrahedron quadrature tables handed to the kernels through ``bp_params``.

For P1 velocities the strain rate is cell-constant, so the only factor of the BP
integrands that needs quadrature is ``A(x)^(-1/n)`` with ``A`` the P1 rate factor
(cslvr/momentum.py:181-185).  To reproduce the reference's numbers the rule must be
the one its form compiler picks: UFL estimates degree 5 for the cell integral of
cslvr/momentumbp.py:481-500 with a P1 ``A``, and FIAT's
default scheme for that degree is Keast's 15-point rule; scripts may override the
degree (simulations/ismip_hom/ismip_hom_a.py:10).

Rules are stored as symmetric orbits in barycentric coordinates,
``(kind, parameters, weight)``; weights sum to one.
"""
import itertools

import numpy as np

# kind -> barycentric generator
_ORBIT = {
	'c':   lambda: (0.25, 0.25, 0.25, 0.25),                       # centroid, 1 point
	'v':   lambda a: (a, a, a, 1.0 - 3.0 * a),                     # vertex orbit, 4 points
	'e':   lambda b: (b, b, 0.5 - b, 0.5 - b),                     # edge orbit, 6 points
	'f':   lambda a, b: (a, a, b, 1.0 - 2.0 * a - b),              # 12 points
}

_RULES = {
	1: [('c', (), 1.0)],
	2: [('v', (0.13819660112501051518,), 0.25)],
	3: [('c', (), -0.8), ('v', (1.0 / 6.0,), 0.45)],
	4: [('e', (0.0,), 2.0 / 105.0),
	    ('v', (0.10052676522520447969,), 0.08858982474298071043),
	    ('v', (0.31437287349319219275,), 0.13283874668559071814)],
	5: [('c', (), 0.18170206858253505484),
	    ('v', (1.0 / 3.0,), 81.0 / 2240.0),
	    ('v', (1.0 / 11.0,), 0.06987149451617381646),
	    ('e', (0.06655015357366429824,), 0.06569484936831875607)],
	6: [('v', (0.21460287125915202929,), 0.03992275025816749210),
	    ('v', (0.04067395853461135312,), 0.01007721105532064295),
	    ('v', (0.32233789014227551034,), 0.05535718154365472210),
	    ('f', (0.06366100187501752530, 0.26967233145831580803), 27.0 / 560.0)],
}


def _collapsed(m):
	"""m^3-point collapsed Gauss-Jacobi rule (the fallback beyond degree 6)."""
	from scipy.special import roots_jacobi
	(r, wr), (s, ws), (t, wt) = (roots_jacobi(m, k, 0) for k in (0, 1, 2))
	R, S, T = np.meshgrid(r, s, t, indexing='ij')
	W = wr[:, None, None] * ws[None, :, None] * wt[None, None, :]
	z = 0.5 * (1.0 + T)
	y = 0.5 * (1.0 + S) * (1.0 - z)
	x = 0.5 * (1.0 + R) * (1.0 - z - y)
	pts = np.stack([1.0 - x - y - z, x, y, z], axis=-1).reshape(-1, 4)
	w = W.ravel()
	return pts, w / w.sum()


def tetrahedron_rule(degree):
	"""barycentric points [nq,4] and weights [nq] (sum 1) exact to ``degree``."""
	degree = max(int(degree), 1)
	if degree not in _RULES:
		return _collapsed((degree + 2) // 2)
	pts, wts = [], []
	for kind, par, w in _RULES[degree]:
		orbit = sorted(set(itertools.permutations(_ORBIT[kind](*par))), reverse=True)
		pts += orbit
		wts += [w] * len(orbit)
	return np.array(pts, dtype=np.float64), np.array(wts, dtype=np.float64)
