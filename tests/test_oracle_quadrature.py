"""The quadrature rules of the oracle and of the product integrate their nominal
degree exactly and agree with each other (the reference never names a rule, the
form compiler picks it: SURVEY.md Appendix A-Q)."""
import itertools
import math

import numpy as np
import pytest

from oracle.quadrature import tetrahedron_rule as o_tet, triangle_rule as o_tri
from cslvr_b200.quadrature import tetrahedron_rule as p_tet


def _exact(e):
	d = len(e) - 1
	return np.prod([math.factorial(i) for i in e]) * math.factorial(d) / math.factorial(sum(e) + d)


def _max_err(pts, wts, deg):
	mx = 0.0
	for d in range(deg + 1):
		for e in itertools.product(range(d + 1), repeat=pts.shape[1]):
			if sum(e) == d:
				mx = max(mx, abs((wts * np.prod(pts ** np.array(e), axis=1)).sum() - _exact(e)))
	return mx


@pytest.mark.parametrize('deg', range(1, 10))
def test_tetrahedron_rules_exact(deg):
	for rule in (o_tet, p_tet):
		p, w = rule(deg)
		assert abs(w.sum() - 1) < 1e-14 and np.abs(p.sum(axis=1) - 1).max() < 1e-14
		assert _max_err(p, w, deg) < 5e-16 * 10


def test_point_counts_match_fiat_defaults():
	# FIAT default schemes (SURVEY A-Q): 1, 4, 5, 14, 15, 24 points for degree 1..6, then m^3
	assert [len(o_tet(d)[1]) for d in range(1, 10)] == [1, 4, 5, 14, 15, 24, 64, 125, 125]
	assert [len(p_tet(d)[1]) for d in range(1, 10)] == [1, 4, 5, 14, 15, 24, 64, 125, 125]
	assert [len(o_tri(d)[1]) for d in (1, 2, 3)] == [1, 3, 6]


@pytest.mark.parametrize('deg', range(1, 10))
def test_product_and_oracle_rules_are_the_same_point_sets(deg):
	po, wo = o_tet(deg)
	pp, wp = p_tet(deg)
	key = lambda p, w: np.round(np.column_stack([np.sort(p, axis=1), w]), 12)
	a = key(po, wo); b = key(pp, wp)
	a = a[np.lexsort(a.T[::-1])]; b = b[np.lexsort(b.T[::-1])]
	assert np.allclose(a, b, atol=1e-12)


@pytest.mark.parametrize('deg', range(1, 8))
def test_triangle_rules_exact(deg):
	p, w = o_tri(deg)
	assert _max_err(p, w, deg) < 5e-15
