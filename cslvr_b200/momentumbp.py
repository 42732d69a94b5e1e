"""``MomentumBPBase`` / ``MomentumBP`` / ``MomentumDukowiczBP`` on the B200 path.

Same class names, constructor arguments, ``solve_params`` dictionary, ``solve()``
sequence and error behaviour as cslvr/momentumbp.py; the variational forms are not
built symbolically -- the kernels implement them in closed form:

* ``MomentumDukowiczBP`` (cslvr/momentumbp.py:406-506): action
  ``A = (V_d - P_e) dOmega - S_l dGamma_b - P_b dGamma_bw [- P_b dGamma_lt]``,
  residual = first variation (:500), Jacobian = second variation (physics.py:75-77);
* ``MomentumBP`` (cslvr/momentumbp.py:289-399): ``inner(sigma, grad v) dOmega +
  rho g grad_h S . v dOmega + beta u.v dGamma_b - p_e n_h.v dGamma_bw[,lt]`` (:372-381),
  which is the same residual term by term (p_e = f_w, sigma:grad v = 2 eta d(epsdot));
  both classes therefore drive the same kernels.
"""
import sys

import numpy as np

from . import capi
from .d3model import D3Model, Function, DOLFIN_EPS
from .inputoutput import print_text, print_min_max
from .momentum import Momentum


class MomentumBPBase(Momentum):

	_reduced_stokes = False      # MomentumDukowiczStokesReduced (momentumstokes.py) flips this

	def __init__(self, model, solve_params=None, linear=False, use_lat_bcs=False,
	             use_pressure_bc=True, device=0):
		print_text("::: INITIALIZING BP VELOCITY BASE PHYSICS :::", cls=self)
		if type(model) != D3Model:
			s = ">>> MomentumBPBase REQUIRES A 'D3Model' INSTANCE, NOT %s <<<"
			print_text(s % type(model), 'red', 1)
			sys.exit(1)
		if model.ff is None:
			raise RuntimeError("model.calculate_boundaries() must be called before a momentum object is created")
		self.Q = model.QM2
		self.device = device
		self.set_unknown(Function(self.Q, name='u_h'))
		self.u_z = Function(model.Q, name='u_z')
		# cslvr/momentumbp.py:101-112: DirichletBC(Q.sub(0), u_x_lat, ff, GAMMA_L_DVD), DirichletBC(Q.sub(1), u_y_lat, ...):
		# kept as (dofs, Function, component) so that the values are read at solve time, as dolfin does
		mom_bcs = []
		self.lateral_u_z_bc = False
		if not model.use_periodic and model.mark_divide and use_lat_bcs:
			print_text("    - using Dirichlet lateral boundary conditions -", cls=self)
			nodes = model.divide_nodes()
			mom_bcs.append((2 * nodes, model.u_x_lat, nodes))
			mom_bcs.append((2 * nodes + 1, model.u_y_lat, nodes))
			self.lateral_u_z_bc = nodes.size > 0      # u_z_bcs gets DirichletBC(Q, u_z_lat, ff, GAMMA_L_DVD) as well
		self.set_boundary_conditions(mom_bcs)
		self._ctx = None
		self._ctx_key = None
		super(MomentumBPBase, self).__init__(model=model, solve_params=solve_params, linear=linear,
		                                     use_lat_bcs=use_lat_bcs, use_pressure_bc=use_pressure_bc)

	def get_velocity(self):
		"""(u_x, u_y, 0) from the unknown, cslvr/momentumbp.py:156-163."""
		v = self.get_unknown().values
		return v[0::2], v[1::2], np.zeros(v.size // 2)

	def default_solve_params(self):
		"""cslvr/momentumbp.py:165-203: Newton + CG + AMG, rtol 1e-9, 25 iterations,
		relaxation 1.0 iff (isothermal or uniform T) and periodic, else 0.7."""
		if (self.model.energy_dependent_rate_factor == False
		    or (self.model.energy_dependent_rate_factor == True
		        and np.unique(self.model.T.vector().get_local()).size == 1)) \
		   and self.model.use_periodic:
			relax = 1.0
		else:
			relax = 0.7
		if self.linear:
			params = {'linear_solver': 'cg', 'preconditioner': 'hypre_amg'}
		else:
			params = {'newton_solver': {
				'linear_solver'           : 'cg',
				'preconditioner'          : 'hypre_amg',
				'relative_tolerance'      : 1e-9,
				'relaxation_parameter'    : relax,
				'maximum_iterations'      : 25,
				'error_on_nonconvergence' : False,
				'krylov_solver'           : {'monitor_convergence': False}}}
		return {'solver'              : params,
		        'solve_vert_velocity' : True,
		        'solve_pressure'      : True,
		        'vert_solve_method'   : 'mumps'}

	# ---- the device context --------------------------------------------------
	def _context(self):
		m = self.model
		X = m.mesh.coordinates()
		# everything bp_create digests: coordinates, connectivity, periodic map, markers, constants.  The
		# connectivity of a mesh object is hashed once (stored on the mesh), coordinates and markers every time.
		import hashlib
		topo = getattr(m.mesh, '_b200_topology_hash', None)
		if topo is None or topo[0] != (m.mesh.cells().shape, bool(m.use_periodic)):
			ht = hashlib.blake2b(digest_size=16)
			for a in (m.mesh.cells(), m.vertex_to_node if m.use_periodic else np.zeros(0), m.facet_cell, m.facet_local):
				ht.update(np.ascontiguousarray(a).view(np.uint8))
			topo = ((m.mesh.cells().shape, bool(m.use_periodic)), ht.hexdigest())
			try:
				m.mesh._b200_topology_hash = topo
			except AttributeError:
				pass
		h = hashlib.blake2b(digest_size=16)
		h.update(topo[1].encode())
		for a in (X, m.ff):
			h.update(np.ascontiguousarray(a).view(np.uint8))
		key = (h.hexdigest(), float(m.n), float(m.eps_reg), float(m.rho_i), float(m.rhosw), float(m.g),
		       bool(m.use_periodic), bool(self.use_pressure_bc), bool(m.energy_dependent_rate_factor),
		       bool(self._reduced_stokes), self.device)
		if self._ctx is None or key != self._ctx_key:
			if self._ctx is not None:
				self._ctx.close()
			params = dict(n=m.n, eps_reg=m.eps_reg, rho_i=m.rho_i, rhosw=m.rhosw, g=m.g,
			              periodic=m.use_periodic, use_pressure_bc=self.use_pressure_bc,
			              energy_dependent_rate_factor=m.energy_dependent_rate_factor,
			              reduced_stokes=self._reduced_stokes,
			              a_T_l=m.a_T_l, a_T_u=m.a_T_u, Q_T_l=m.Q_T_l, Q_T_u=m.Q_T_u, R=m.R)
			self._ctx = capi.BPContext(X, m.mesh.cells(), self.Q.dim(),
			                           (m.facet_cell, m.facet_local, m.ff), params,
			                           vertex_to_node=m.vertex_to_node if m.use_periodic else None,
			                           device=self.device)
			self._ctx_key = key
		c = self._ctx
		c.set_field(capi.FIELD_S, m.vertex_field(m.S))
		if m.energy_dependent_rate_factor:
			c.set_field(capi.FIELD_TP, m.vertex_field(m.Tp))
			c.set_field(capi.FIELD_W, m.vertex_field(m.W))
			c.set_field(capi.FIELD_E, m.vertex_field(m.E))
		else:
			c.set_field(capi.FIELD_A, m.vertex_field(m.A))
		c.set_field(capi.FIELD_BETA, m.vertex_field(m.beta))
		if self._reduced_stokes:
			# the frozen u_z of the strain-rate tensor and B_ring of u_z_b (cslvr/momentumstokes.py:322, 361-370)
			c.set_field(capi.FIELD_UZ, m.vertex_field(self.u_z))
			c.set_field(capi.FIELD_B, m.vertex_field(m.B))
			if getattr(m, 'B_ring', None) is not None:
				c.set_field(capi.FIELD_B_RING, m.vertex_field(m.B_ring))
		return c

	# the reference asks PETSc for hypre BoomerAMG; the device PCG offers Jacobi-type,
	# Chebyshev-polynomial, vertical-line and aggregation-multigrid preconditioners -- AMG
	# requests get the multigrid V-cycle
	_PC_MAP = {'hypre_amg': 'mg', 'amg': 'mg', 'default': 'mg', 'mg': 'mg',
	           'none': 'none', 'jacobi': 'jacobi', 'block_jacobi': 'block_jacobi',
	           'chebyshev': 'chebyshev', 'column': 'column'}

	def _solver_kwargs(self, ns):
		ks = ns.get('krylov_solver', {})
		pc = ns.get('preconditioner', 'default')
		if ns.get('linear_solver', 'cg') not in ('cg', 'default'):
			print_text("    - linear_solver '%s' is served by the device PCG (the BP Hessian is SPD) -"
			           % ns['linear_solver'], cls=self)
		kw = dict(relative_tolerance=ns['relative_tolerance'],
		          absolute_tolerance=ns.get('absolute_tolerance', 1e-10),
		          maximum_iterations=ns['maximum_iterations'],
		          relaxation_parameter=ns['relaxation_parameter'],
		          error_on_nonconvergence=ns.get('error_on_nonconvergence', False),
		          krylov_rtol=ks.get('relative_tolerance', 1e-5),
		          krylov_atol=ks.get('absolute_tolerance', 1e-50),
		          krylov_maxit=ks.get('maximum_iterations', 10000),
		          preconditioner=self._PC_MAP.get(pc, 'mg'),
		          verbose=ks.get('monitor_convergence', False) or ns.get('report', False),
		          krylov_error_on_nonconvergence=ks.get('error_on_nonconvergence', True),
		          quadrature_degree=self._quadrature_degree())
		return kw

	def _quadrature_degree(self):
		"""UFL's estimate for the cell integrals (SURVEY A-Q): 5 with a P1 rate factor, 9 with the
		energy-dependent expression at the quadrature points; ``solve_params['quadrature_degree']``
		overrides, as ``parameters['form_compiler']['quadrature_degree']`` does in the reference."""
		return self.solve_params.get('quadrature_degree', 9 if self.model.energy_dependent_rate_factor else 5)

	def _apply_boundary_conditions(self, c):
		"""hand the DirichletBCs (cslvr/momentumbp.py:101-112) to the library: bp_set_dirichlet."""
		bcs = self.get_boundary_conditions() or []
		if bcs:
			dofs = np.concatenate([b[0] for b in bcs])
			vals = np.concatenate([b[1].values[b[2]] for b in bcs])
			c.set_dirichlet(dofs, vals)
		else:
			c.set_dirichlet(np.zeros(0, dtype=np.int32), np.zeros(0))

	def _newton_solve(self, ns):
		c = self._context()
		self._apply_boundary_conditions(c)
		c.set_params(**self._solver_kwargs(ns))
		u = self.get_unknown()
		stats = c.newton_solve(u.values)
		return stats

	def _linear_solve(self, sp):
		"""linear=True: viscosity from model.u (cslvr/momentumbp.py:91-93)."""
		c = self._context()
		# after linearize_viscosity() the dictionary still has the non-linear layout
		# (cslvr/momentum.py:105-118), so look for the Krylov settings at both levels
		sp = sp if isinstance(sp, dict) else {}
		inner = sp.get('newton_solver', sp)
		pc = inner.get('preconditioner', 'default')
		ks = inner.get('krylov_solver', {})
		c.set_params(krylov_rtol=ks.get('relative_tolerance', 1e-5), krylov_atol=ks.get('absolute_tolerance', 1e-50),
		             krylov_maxit=ks.get('maximum_iterations', 10000), preconditioner=self._PC_MAP.get(pc, 'mg'),
		             quadrature_degree=self._quadrature_degree())
		self._apply_boundary_conditions(c)
		ul = np.empty(self.Q.dim())
		ul[0::2], ul[1::2] = self.model.u.values[0::3], self.model.u.values[1::3]
		self.get_unknown().values[:] = c.linear_solve(ul)
		return dict(converged=True, iterations=1, krylov_iterations=c.last_linear_iterations)

	# ---- assemble(residual) / assemble(jacobian) ------------------------------
	def assemble_residual(self, u=None):
		c = self._context()
		c.set_params(quadrature_degree=self._quadrature_degree())
		F, _ = c.assemble(self.get_unknown().values if u is None else u, want_F=True, want_J=False)
		return F

	def assemble_jacobian(self, u=None):
		"""(indptr, indices, values) of the CSR Jacobian in the QM2 dof numbering."""
		c = self._context()
		c.set_params(quadrature_degree=self._quadrature_degree())
		_, J = c.assemble(self.get_unknown().values if u is None else u, want_F=False, want_J=True)
		ip, ix = c.csr_pattern()
		return ip, ix, J

	# ---- cslvr/momentumbp.py:205-253 ------------------------
	def solve_pressure(self, annotate=False):
		print_text("::: solving BP pressure :::", cls=self)
		if annotate:
			raise NotImplementedError("annotate=True cannot be served by an external solver")
		# p = project(rho g (S - z) - 2 eta (u_x,x + u_y,y), Q), eta = eta(model.u) which at
		# this point holds the velocity just solved (cslvr/momentumbp.py:216-223)
		c = self._context()
		p_vertex = c.solve_pressure(self.get_unknown().values)
		self.model.p.values[:] = p_vertex[self.model.node_vertex]
		print_min_max(self.model.p, 'p', cls=self)

	def solve_vert_velocity(self, annotate=False):
		print_text("::: solving BP vertical velocity :::", cls=self)
		if annotate:
			raise NotImplementedError("annotate=True cannot be served by an external solver")
		# project the basal kinematic condition, assemble the continuity system, apply the
		# basal Dirichlet rows and solve (cslvr/momentumbp.py:231-253) -- all on the device
		m = self.model
		c = self._context()
		if getattr(self, 'lateral_u_z_bc', False):         # u_z_bcs also holds DirichletBC(Q, u_z_lat, ff, GAMMA_L_DVD), momentumbp.py:110-111
			nodes = m.divide_nodes()
			c.set_dirichlet_uz(np.asarray(m.node_vertex)[nodes], m.u_z_lat.values[nodes])
		else:
			c.set_dirichlet_uz(np.zeros(0, dtype=np.int32), np.zeros(0))
		c.set_field(capi.FIELD_B, m.vertex_field(m.B))
		if getattr(m, 'B_ring', None) is not None:
			c.set_field(capi.FIELD_B_RING, m.vertex_field(m.B_ring))
		uz_vertex = c.solve_vert_velocity(self.get_unknown().values)
		self.u_z.values[:] = uz_vertex[m.node_vertex]
		m.u.values[2::3] = self.u_z.values
		print_min_max(self.u_z, 'w')

	def solve(self, annotate=False):
		"""Newton solve of the first-order equations, cslvr/momentumbp.py:255-273:
		cold start from DOLFIN_EPS, Physics.solve, then u_z and p if requested."""
		model  = self.model
		params = self.solve_params
		model.assign_variable(self.get_unknown(), DOLFIN_EPS, annotate=annotate, cls=self)
		super(MomentumBPBase, self).solve(annotate)
		if params['solve_vert_velocity']: self.solve_vert_velocity(annotate)
		if params['solve_pressure']:      self.solve_pressure(annotate)

	def update_model_var(self, u, annotate=False):
		"""copy (u_x, u_y) into model.u, cslvr/momentumbp.py:275-282."""
		self.model.u.values[0::3] = u.values[0::2]
		self.model.u.values[1::3] = u.values[1::2]
		print_min_max(self.model.u, 'model.u', cls=self)


class MomentumBP(MomentumBPBase):

	def initialize(self, model, solve_params=None, linear=False, use_lat_bcs=False, use_pressure_bc=True):
		print_text("::: INITIALIZING BP VELOCITY PHYSICS :::", cls=self)
		if (not model.use_periodic and use_pressure_bc):
			print_text("    - using water pressure lateral boundary condition -", cls=self)
		self.set_residual('bp-classic')


class MomentumDukowiczBP(MomentumBPBase):

	def initialize(self, model, solve_params=None, linear=False, use_lat_bcs=False, use_pressure_bc=True):
		print_text("::: INITIALIZING DUKOWICZ BP VELOCITY PHYSICS :::", cls=self)
		if (not model.use_periodic and use_pressure_bc):
			print_text("    - using water pressure lateral boundary condition -", cls=self)
		if use_lat_bcs:
			# cslvr/momentumbp.py:495 calls quasi_stress_tensor with three arguments (it takes two, :150): the
			# reference raises here as well; MomentumBP (whose stress term is commented out, :384-393) does not
			raise TypeError("quasi_stress_tensor() takes 3 positional arguments but 4 were given "
			                "(use_lat_bcs=True is broken in cslvr/momentumbp.py:495 too; use MomentumBP)")
		self.set_residual('bp-dukowicz')
