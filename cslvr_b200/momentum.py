"""``cslvr.Momentum``: parameter handling and the closed-form constitutive relations.

Mirrors cslvr/momentum.py:27-66 (constructor), :78-95 (``reset``) and :302-328
(default parameters).  ``get_viscosity`` / ``effective_strain_rate`` /
``get_viscous_dissipation`` (:159-251) are available as numpy evaluations of the
same formulas for post-processing; the solve itself evaluates them in the kernels.
"""
import json
import sys
from copy import deepcopy

import numpy as np

from .inputoutput import print_text
from .physics import Physics


class Momentum(Physics):

	def __init__(self, model, solve_params=None, linear=False, use_lat_bcs=False,
	             use_pressure_bc=True, **kwargs):
		print_text("::: INITIALIZING MOMENTUM :::", cls=self)
		self.linear          = linear
		self.use_pressure_bc = use_pressure_bc
		self.use_lat_bcs     = use_lat_bcs
		if isinstance(solve_params, dict):
			self.solve_params = solve_params
		elif solve_params is None:
			self.solve_params = self.default_solve_params()
			print_text("::: using default parameters :::", cls=self)
			print_text(json.dumps(self.solve_params, sort_keys=True, indent=2), '230')
		else:
			s = ">>> Momentum REQUIRES A 'dict' INSTANCE OF SOLVER PARAMETERS, NOT %s <<<"
			print_text(s % type(solve_params), 'red', 1)
			sys.exit(1)
		self.solve_params_s    = deepcopy(self.solve_params)
		self.linear_s          = linear
		self.use_lat_bcs_s     = use_lat_bcs
		self.use_pressure_bc_s = use_pressure_bc
		self.kwargs            = kwargs
		self.initialize(model=model, solve_params=self.solve_params, linear=self.linear,
		                use_lat_bcs=self.use_lat_bcs, use_pressure_bc=self.use_pressure_bc, **kwargs)

	def initialize(self, model, solve_params=None, linear=False, use_lat_bcs=False,
	               use_pressure_bc=True, **kwargs):
		raise NotImplementedError

	def reset(self):
		print_text("::: RE-INITIALIZING MOMENTUM PHYSICS :::", cls=self)
		self.initialize(model=self.model, solve_params=self.solve_params_s, linear=self.linear_s,
		                use_lat_bcs=self.use_lat_bcs_s, use_pressure_bc=self.use_pressure_bc_s, **self.kwargs)

	def linearize_viscosity(self, reset_orig_config=True):
		"""linearize the viscosity using the velocity stored in ``model.u``
		(cslvr/momentum.py:97-138): re-initialise with linear=True, 2 Newton iterations max,
		relaxation 1.0, no u_z / p solves."""
		print_text("::: RE-INITIALIZING MOMENTUM PHYSICS WITH LINEAR VISCOSITY :::", cls=self)
		mom_params = deepcopy(self.solve_params_s)
		new_params = mom_params['solver']['newton_solver']
		mom_params['solve_vert_velocity']     = False
		mom_params['solve_pressure']          = False
		new_params['relaxation_parameter']    = 1.0
		new_params['maximum_iterations']      = 2
		new_params['error_on_nonconvergence'] = False
		print_text("::: altering solver parameters for optimal convergence :::", cls=self)
		print_text(json.dumps(mom_params, sort_keys=True, indent=2), '230')
		if reset_orig_config:
			print_text("::: reseting the original config to use linear viscosity :::", cls=self)
			self.linear_s       = True
			self.solve_params_s = mom_params
		self.linear = True
		self.solve_params = mom_params
		self.initialize(model=self.model, solve_params=mom_params, linear=True,
		                use_lat_bcs=self.use_lat_bcs_s, use_pressure_bc=self.use_pressure_bc_s, **self.kwargs)

	def color(self):
		return 'cyan'

	# ---- constitutive relations, evaluated per cell ---------------------------
	def effective_strain_rate(self, g):
		"""0.5 tr(eps.eps) of the BP strain-rate tensor from velocity gradients
		g[...,c,j] (cslvr/momentum.py:245-251 with cslvr/momentumbp.py:122-135)."""
		sh = 0.5 * (g[..., 0, 1] + g[..., 1, 0])
		return (g[..., 0, 0] ** 2 + g[..., 1, 1] ** 2 + g[..., 0, 0] * g[..., 1, 1] + sh ** 2
		        + 0.25 * g[..., 0, 2] ** 2 + 0.25 * g[..., 1, 2] ** 2)

	def get_viscosity(self, epsdot, A):
		"""eta = 1/2 A^(-1/n) (epsdot + eps_reg)^((1-n)/(2n)), cslvr/momentum.py:181-185."""
		n, e0 = self.model.n, self.model.eps_reg
		return 0.5 * A ** (-1.0 / n) * (epsdot + e0) ** ((1.0 - n) / (2.0 * n))

	def get_viscous_dissipation(self, epsdot, A):
		"""V = 2n/(n+1) A^(-1/n) (epsdot + eps_reg)^((n+1)/(2n)), cslvr/momentum.py:218-221."""
		n, e0 = self.model.n, self.model.eps_reg
		return (2 * n) / (n + 1) * A ** (-1.0 / n) * (epsdot + e0) ** ((n + 1) / (2 * n))
