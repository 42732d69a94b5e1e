"""rank-0 logging with cslvr's call signatures (cslvr/inputoutput.py:468-555).

``print_text`` / ``print_min_max`` keep their argument meaning; the colour codes
of the reference need the ``colored`` package, so text is printed plain.  Set
``cslvr_b200.inputoutput.QUIET = True`` to silence (tests, benchmarks).
"""
import os

import numpy as np

QUIET = os.environ.get('CSLVR_B200_QUIET', '0') == '1'


def _rank():
	try:
		return int(os.environ.get('RANK', '0'))
	except ValueError:
		return 0


def get_text(text, color=None, atrb=0, cls=None):
	return text


def print_text(text, color=None, atrb=0, cls=None):
	if not QUIET and _rank() == 0:
		print(get_text(text, color, atrb, cls))


def print_min_max(u, title, color='97', cls=None):
	"""print ``title <min, max>`` of a Function, array or number."""
	if hasattr(u, 'values'):
		u = u.values
	if isinstance(u, np.ndarray):
		print_text(title + ' <min, max> : <%.3e, %.3e>' % (u.min(), u.max()), color, cls=cls)
	elif isinstance(u, (int, float)):
		print_text(title + ' : %.3e' % u, color, cls=cls)
	else:
		print_text(title + ": print_min_max function requires a Vector, Function, array, int or float, not %s." % type(u), 'red', 1)
