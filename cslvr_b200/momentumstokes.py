"""``MomentumDukowiczStokesReduced`` on the B200 path.

cslvr/momentumstokes.py:300-507: the Dukowicz reduced full-Stokes model shares the
Blatter-Pattyn unknown (u_x, u_y) in QM2 and ``MomentumBPBase``; what differs is

* the strain-rate tensor (:406-418) carries the frozen P1 vertical velocity ``self.u_z``:
  eps_02 = (u_x,z + u_z,x)/2, eps_12 = (u_y,z + u_z,y)/2, eps_22 = -(u_x,x + u_y,y);
* sliding acts on the grounded bed only and includes u_z_b = (-B_ring - u_h.n_h)/n_z (:369-370);
* ``solve()`` (:449-486) runs ``Model.home_rolled_newton_method`` (cslvr/model.py:2610-2730)
  with ``solve_vert_velocity`` as the per-iteration callback, atol 1e-6, relaxation 0.8,
  50 iterations, THEN ``solve_pressure`` and only then ``update_model_var``.

The kernels are the BP ones with the two shifted strain components and the extra facet
terms (``bp_params.reduced_stokes``); the loop is ``bp_rs_newton_solve``.
"""
import numpy as np

from . import capi
from .inputoutput import print_text, print_min_max
from .momentumbp import MomentumBPBase


class MomentumDukowiczStokesReduced(MomentumBPBase):

	_reduced_stokes = True

	def initialize(self, model, solve_params=None, linear=False, use_lat_bcs=False, use_pressure_bc=True):
		print_text("::: INITIALIZING DUKOWICZ REDUCED FULL-STOKES PHYSICS :::", cls=self)
		if (not model.use_periodic and use_pressure_bc):
			print_text("    - using water pressure lateral boundary condition -", cls=self)
		self.set_residual('rs-dukowicz')

	def get_velocity(self):
		"""(u_x, u_y, u_z): horizontal components from the unknown, vertical from ``self.u_z``
		(cslvr/momentumstokes.py:396-404)."""
		v = self.get_unknown().values
		return v[0::2], v[1::2], self.u_z.values

	def default_solve_params(self):
		"""cslvr/momentumstokes.py:420-447."""
		nparams = {'newton_solver': {
			'linear_solver'           : 'cg',
			'preconditioner'          : 'hypre_amg',
			'relative_tolerance'      : 1e-9,
			'relaxation_parameter'    : 0.8,
			'maximum_iterations'      : 50,
			'error_on_nonconvergence' : False,
			'krylov_solver'           : {'monitor_convergence': False}}}
		return {'solver'              : nparams,
		        'solve_vert_velocity' : True,
		        'solve_pressure'      : True,
		        'vert_solve_method'   : 'mumps'}

	def solve_pressure(self, annotate=False):
		"""``MomentumBPBase.solve_pressure`` as this class reaches it: ``self.eta`` is the viscosity
		of ``model.u`` (cslvr/momentumbp.py:92, 219) -- inside ``solve()`` that is the velocity the
		model held before the Newton loop plus the u_z of the last callback, because
		``update_model_var`` runs afterwards (cslvr/momentumstokes.py:478-482)."""
		print_text("::: solving BP pressure :::", cls=self)
		if annotate:
			raise NotImplementedError("annotate=True cannot be served by an external solver")
		m = self.model
		c = self._context()
		c.set_field(capi.FIELD_UZ, np.ascontiguousarray(m.u.values[2::3][m.vertex_to_node]))
		u_model = np.empty(self.Q.dim())
		u_model[0::2], u_model[1::2] = m.u.values[0::3], m.u.values[1::3]
		p_vertex = c.rs_solve_pressure(self.get_unknown().values, u_model)
		c.set_field(capi.FIELD_UZ, m.vertex_field(self.u_z))
		m.p.values[:] = p_vertex[m.node_vertex]
		print_min_max(m.p, 'p', cls=self)

	def solve(self, annotate=False):
		"""Newton solve of the reduced full-Stokes equations, cslvr/momentumstokes.py:449-486."""
		model  = self.model
		params = self.solve_params
		if annotate:
			raise NotImplementedError("annotate=True cannot be served by an external solver")
		if self.linear:
			raise NotImplementedError("MomentumDukowiczStokesReduced.solve() always runs the non-linear loop "
			                          "(cslvr/momentumstokes.py:471); use Physics.solve for linear=True")
		ns       = params['solver']['newton_solver']
		rtol     = ns['relative_tolerance']
		maxit    = ns['maximum_iterations']
		alpha    = ns['relaxation_parameter']
		s    = "::: solving Dukowicz full-Stokes reduced equations with %i max" + \
		         " iterations and step size = %.1f :::"
		print_text(s % (maxit, alpha), cls=self)

		c = self._context()
		kw = self._solver_kwargs(ns)
		c.set_params(**kw)
		u  = self.get_unknown()
		uz = model.vertex_field(self.u_z)
		# cb_ftn = self.solve_vert_velocity (:468); atol = 1e-6 (:475)
		self.stats = c.rs_newton_solve(u.values, uz, atol=1e-6, callback=True)
		for k, r in enumerate(self.stats['residuals']):
			print_text("Newton iteration %d: r (abs) = %.3e (tol = %.3e) r (rel) = %.3e (tol = %.3e)"
			           % (k + 1, r, 1e-6, r / self.stats['residuals'][0] if self.stats['residuals'][0] else 0.0, rtol), cls=self)
		# what the callback leaves behind: self.u_z and model.u.sub(2) (cslvr/momentumbp.py:252)
		self.u_z.values[:] = uz[model.node_vertex]
		model.u.values[2::3] = self.u_z.values
		print_min_max(self.u_z, 'w')

		if params['solve_pressure']:    self.solve_pressure(annotate)
		self.update_model_var(self.get_unknown(), annotate=annotate)
