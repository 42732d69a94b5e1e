"""``cslvr.Physics`` with the dolfin Newton call replaced by the C-ABI.

The reference leaves Python exactly once on this path,
``solve(resid == 0, u, J=Jac, bcs=bcs, solver_parameters=params['solver'])`` at
cslvr/physics.py:673-678; :meth:`Physics.solve` below makes the same call into
``bp_newton_solve`` and keeps everything around it: the parameter reads
(:665-667), the banner, the "time to solve" line (:681-688) and
``update_model_var`` (:691).
"""
import sys
from time import time

from .inputoutput import print_text


class Physics(object):

	def __new__(cls, model, *args, **kwargs):
		instance = object.__new__(cls)
		instance.model = model
		return instance

	def color(self):
		return 'white'

	def default_solve_params(self):
		return {'solver': 'mumps'}

	def get_solve_params(self):
		return self.solve_params

	# residual / Jacobian are not symbolic objects here: they are the two things
	# ``assemble()`` can be asked for (cslvr/physics.py:54-90)
	def get_residual(self):
		return self.resid

	def set_residual(self, resid):
		self.resid = resid
		if not self.linear:
			self.Jac = ('jacobian-of', resid)

	def get_jacobian(self):
		assert not self.linear, "this problem is linear; there is no Jacobian."
		return self.Jac

	def set_boundary_conditions(self, bc_list):
		self.bcs = bc_list

	def get_boundary_conditions(self):
		return self.bcs

	def get_unknown(self):
		return self.u

	def set_unknown(self, u):
		self.u = u

	def solve(self, annotate=False):
		"""Newton solve of ``resid == 0`` on the GPU (non-linear branch of
		cslvr/physics.py:644-691)."""
		params = self.solve_params
		if annotate:
			raise NotImplementedError("annotate=True (dolfin-adjoint taping, cslvr/physics.py:481) "
			                          "cannot be served by an external solver")
		t0 = time()
		if self.linear:
			# cslvr/physics.py:655-661: solve(lhs == rhs, u, bcs, solver_parameters=params['solver'])
			print_text("::: solving linear equations :::", cls=self)
			self.stats = self._linear_solve(params['solver'])
			t_tot = time() - t0
			print_text("time to solve: %02d:%02d:%02d:%03d" % (t_tot / 3600, (t_tot / 60) % 60, t_tot % 60, (t_tot % 1) * 1000), 'red', 1)
			self.update_model_var(self.get_unknown(), annotate=annotate)
			return
		ns     = params['solver']['newton_solver']
		rtol   = ns['relative_tolerance']
		maxit  = ns['maximum_iterations']
		alpha  = ns['relaxation_parameter']
		s      = "::: solving nonlinear system with %i max iterations and step size = %.1f :::"
		print_text(s % (maxit, alpha), cls=self)

		self.stats = self._newton_solve(ns)

		t_tot = time() - t0
		m     = t_tot / 60.0
		h     = m / 60.0
		m     = m % 60
		s     = t_tot % 60
		ms    = (s % 1) * 1000
		print_text("time to solve: %02d:%02d:%02d:%03d" % (h, m, s, ms), 'red', 1)
		self.update_model_var(self.get_unknown(), annotate=annotate)

	def update_model_var(self, u, annotate=False):
		raise NotImplementedError

	def _newton_solve(self, newton_params):
		raise NotImplementedError

	def _linear_solve(self, solver_params):
		raise NotImplementedError
