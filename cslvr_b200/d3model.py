"""Host-side mirror of ``cslvr.D3Model`` for the Blatter-Pattyn hot path.

Keeps the names and call order of a cslvr script (docs/get_started.rst:13-45,
simulations/ismip_hom/ismip_hom_c.py) so the same lines run against the B200 path::

    mesh  = BoxMesh(Point(0, 0, 0), Point(L, L, 1), 15, 15, 5)
    model = D3Model(mesh, out_dir='./results/', use_periodic=True)
    model.calculate_boundaries()
    model.deform_mesh_to_geometry(surface, bed)
    model.init_beta(1000); model.init_A(1e-16)
    mom = MomentumDukowiczBP(model); mom.solve()

Only what the momentum forms consume is here: the mesh, the periodic vertex->node
map (cslvr/d3model.py:62-108), facet/cell markers (:337-475), the measures' marker
sets (:556-602), mesh deformation (:629-667), the P1 coefficient Functions
(cslvr/model.py:2500-2535) and the constants (cslvr/model.py:190-285).  FEniCS
``Expression`` objects are replaced by callables ``f(x, y, z)`` on numpy arrays.
Everything is numpy: this is one-off set-up, the hot path is in the CUDA library.
"""
import sys

import numpy as np

from .mesh import Mesh, facet_geometry
from .inputoutput import print_text, print_min_max

DOLFIN_EPS = 3.0e-16


class FunctionSpace(object):
	"""a P1-type space: 'Q' / 'QM2' / 'V' live on nodes (periodic identification
	applied), 'Q_non_periodic' on mesh vertices (cslvr/model.py:440-481)."""

	def __init__(self, model, kind, n_comp, periodic):
		self.model, self.kind, self.n_comp, self.periodic = model, kind, n_comp, periodic

	def dim(self):
		n = self.model.n_nodes if self.periodic else self.model.num_vertices
		return n * self.n_comp


class _Vector(object):
	"""the slice of dolfin's GenericVector API the reference scripts use."""

	def __init__(self, f): self._f = f
	def get_local(self): return self._f.values.copy()
	def array(self): return self._f.values.copy()
	def set_local(self, v): self._f.values[:] = np.asarray(v, dtype=np.float64)
	def apply(self, mode): pass
	def min(self): return float(self._f.values.min())
	def max(self): return float(self._f.values.max())
	def size(self): return self._f.values.size
	def norm(self, kind='l2'): return float(np.linalg.norm(self._f.values))
	def __setitem__(self, k, v): self._f.values[k] = v
	def __getitem__(self, k): return self._f.values[k]


class Function(object):
	"""a P1 Function: nodal values in ``values`` (interleaved for vector spaces)."""

	def __init__(self, space, name='f'):
		self.space, self._name = space, name
		self.values = np.zeros(space.dim())

	def name(self): return self._name
	def function_space(self): return self.space
	def vector(self): return _Vector(self)

	def sub(self, c):
		return self.values[c::self.space.n_comp]

	def compute_vertex_values(self):
		"""values at the mesh vertices (periodic slaves take the master's value)."""
		m = self.space.model
		idx = m.vertex_to_node if self.space.periodic else np.arange(m.num_vertices)
		nc = self.space.n_comp
		if nc == 1:
			return self.values[idx]
		return np.concatenate([self.values[c::nc][idx] for c in range(nc)])


class Boundary(object):
	"""the marker list a measure integrates over (cslvr/helper.py:41-52)."""

	def __init__(self, kind, markers, description):
		self.kind, self.markers, self.description = kind, list(markers), description


class D3Model(object):
	OMEGA_GND   = 0   # internal cells over bedrock
	OMEGA_FLT   = 1   # internal cells over water
	GAMMA_S_GND = 2   # grounded upper surface
	GAMMA_B_GND = 3   # grounded lower surface (bedrock)
	GAMMA_S_FLT = 6   # shelf upper surface
	GAMMA_B_FLT = 5   # shelf lower surface
	GAMMA_L_DVD = 7   # basin divides
	GAMMA_L_OVR = 4   # terminus over water
	GAMMA_L_UDR = 10  # terminus under water
	GAMMA_U_GND = 8   # grounded surface with U observations
	GAMMA_U_FLT = 9   # shelf surface with U observations
	GAMMA_ACC   = 1   # areas with positive surface accumulation

	def __init__(self, mesh, out_dir='./results/', order=1, use_periodic=False):
		print_text("::: INITIALIZING 3D MODEL :::", cls=self)
		if order != 1:
			raise NotImplementedError("the B200 Blatter-Pattyn path covers P1 (order=1) only "
			                          "(cslvr/model.py:438-444 allows order 2)")
		if not isinstance(mesh, Mesh):
			raise TypeError("D3Model requires a cslvr_b200.Mesh (e.g. BoxMesh(...))")
		self.out_dir = out_dir
		self.order = order
		self.use_periodic = use_periodic
		self.energy_dependent_rate_factor = False
		self.mark_divide = False
		self.generate_constants()
		self.set_mesh(mesh)
		if use_periodic:
			self.generate_pbc()
		else:
			self.vertex_to_node = np.arange(self.num_vertices, dtype=np.int32)
			self.n_nodes = self.num_vertices
			self.node_vertex = np.arange(self.num_vertices, dtype=np.int64)
		self.generate_function_spaces()
		self.initialize_variables()

	def color(self):
		return '130'

	# ---- cslvr/model.py:119-285 ------------------------------------------
	def generate_constants(self):
		spy = 31556926.0
		self.spy     = spy
		self.eps_reg = 1e-15
		self.n       = 3.0
		self.rho_i   = 910.0
		self.rho_w   = 1000.0
		self.rhosw   = 1028.0
		self.g       = 9.80665
		self.R       = 8.3144621
		self.T_w     = 273.15
		self.a_T_l   = 3.985e-13 * spy
		self.a_T_u   = 1.916e3 * spy
		self.Q_T_l   = 6e4
		self.Q_T_u   = 13.9e4

	# ---- cslvr/d3model.py:110-135 ----------------------------------------
	def set_mesh(self, mesh):
		print_text("::: setting 3D mesh :::", cls=self)
		self.mesh = mesh
		self.dim = 3
		self.num_cells = mesh.num_cells()
		self.num_vertices = mesh.num_vertices()
		self.flat_coordinates = mesh.coordinates().copy()
		print_text("    - 3D mesh set, %i cells, %i vertices - " % (self.num_cells, self.num_vertices), cls=self)

	# ---- cslvr/d3model.py:62-108 -----------------------------------------
	def generate_pbc(self):
		"""periodic identification x = xmax -> xmin, y = ymax -> ymin (corner -> origin),
		z unchanged; evaluated on the flat mesh, before any deformation
		(cslvr/model.py:473-475)."""
		print_text("    - using 3D periodic boundaries -", cls=self)
		X = self.mesh.coordinates()
		lo = X.min(axis=0)
		hi = X.max(axis=0)
		img = X.copy()
		for d in (0, 1):
			on_max = np.abs(X[:, d] - hi[d]) <= DOLFIN_EPS * max(1.0, abs(hi[d]))
			img[on_max, d] -= hi[d] - lo[d]
		ext = np.where(hi > lo, hi - lo, 1.0)

		def key(P):
			q = np.rint((P - lo) / ext * (1 << 20)).astype(np.int64)
			return (q[:, 0] << 42) | (q[:, 1] << 21) | q[:, 2]
		kv = key(X)
		order = np.argsort(kv, kind='stable')
		pos = np.searchsorted(kv[order], key(img))
		master = order[np.minimum(pos, order.size - 1)]
		if not np.array_equal(kv[master], key(img)):
			raise ValueError("periodic image of a boundary vertex is not a mesh vertex")
		is_master = master == np.arange(master.size)
		node_of = np.cumsum(is_master) - 1
		self.vertex_to_node = node_of[master].astype(np.int32)
		self.n_nodes = int(is_master.sum())
		self.node_vertex = np.flatnonzero(is_master)

	# ---- cslvr/model.py:416-484 ------------------------------------------
	def generate_function_spaces(self):
		print_text("::: generating O(1) fundamental function spaces :::", cls=self)
		self.Q   = FunctionSpace(self, 'Q', 1, True)
		self.V   = FunctionSpace(self, 'V', 3, True)
		self.QM2 = FunctionSpace(self, 'QM2', 2, True)
		self.Q_non_periodic = FunctionSpace(self, 'Q_non_periodic', 1, False)
		# UFC mixed P1xP1 cell dofs, component-major; canonical dof = 2*node + c
		nd = self.vertex_to_node[self.mesh.cells()]
		self.cell_dofs = np.concatenate([2 * nd, 2 * nd + 1], axis=1).astype(np.int32)

	# ---- cslvr/model.py:2500-2535, cslvr/d3model.py:847-861 ---------------
	def initialize_variables(self):
		print_text("::: initializing 3D variables :::", cls=self)
		self.mask   = Function(self.Q, 'mask')
		self.S      = Function(self.Q_non_periodic, 'S')
		self.B      = Function(self.Q_non_periodic, 'B')
		self.u      = Function(self.V, 'u')
		self.p      = Function(self.Q, 'p')
		self.beta   = Function(self.Q, 'beta')
		self.E      = Function(self.Q, 'E')
		self.A      = Function(self.Q, 'A')
		self.sigma  = Function(self.Q_non_periodic, 'sigma')
		self.T      = Function(self.Q, 'T')             # cslvr/model.py:2544-2551
		self.Tp     = Function(self.Q, 'Tp')
		self.W      = Function(self.Q, 'W')
		self.u_x_lat = Function(self.Q, 'u_x_lat')      # cslvr/model.py:2531-2533: velocity on the basin divides
		self.u_y_lat = Function(self.Q, 'u_y_lat')
		self.u_z_lat = Function(self.Q, 'u_z_lat')
		self.E.values[:] = 1.0
		self.ff = None

	# ---- coefficient setters: cslvr/model.py:486-497, 691-733, 2166-2221 ---
	def assign_variable(self, u, var, annotate=False, cls=None):
		"""``var`` may be a number, an array of nodal values, a Function on the same
		space, or a callable f(x, y, z) standing in for a FEniCS Expression (nodal
		interpolation at the current mesh coordinates)."""
		if cls is None: cls = self
		if isinstance(var, (int, float, np.floating, np.integer)):
			u.values[:] = float(var)
		elif isinstance(var, np.ndarray):
			if var.size != u.values.size:
				raise ValueError("assign_variable: %s has %d dofs, got %d values" % (u.name(), u.values.size, var.size))
			u.values[:] = var.astype(np.float64).ravel()
		elif isinstance(var, Function):
			u.values[:] = var.values
		elif callable(var):
			X = self.mesh.coordinates()
			if u.space.periodic:
				X = X[self.node_vertex]
			u.values[:] = np.asarray(var(X[:, 0], X[:, 1], X[:, 2]), dtype=np.float64) * np.ones(X.shape[0])
		else:
			raise TypeError("assign_variable() requires a number, array, Function or callable, not %s" % type(var))
		print_min_max(u, u.name(), cls=cls)

	def init_S(self, S):
		print_text("::: initializng surface topography :::", cls=self)
		self.assign_variable(self.S, S)

	def init_B(self, B):
		print_text("::: initializng basal topography :::", cls=self)
		self.assign_variable(self.B, B)

	def init_u_lat(self, u_lat):
		"""x-component of the lateral (basin divide) velocity, cslvr/model.py:976-985 (which assigns to a
		``self.u_lat`` that does not exist; ``u_x_lat`` is the Function momentumbp.py:106 reads)."""
		print_text("::: initializing u lateral b.c. :::", cls=self)
		self.assign_variable(self.u_x_lat, u_lat)

	def init_u_y_lat(self, u_y_lat):
		print_text("::: initializing v lateral b.c. :::", cls=self)                   # cslvr/model.py:987-996
		self.assign_variable(self.u_y_lat, u_y_lat)

	def init_u_z_lat(self, u_z_lat):
		print_text("::: initializing w lateral b.c. :::", cls=self)                   # cslvr/model.py:998-1007
		self.assign_variable(self.u_z_lat, u_z_lat)

	def divide_nodes(self):
		"""nodes of the facets marked GAMMA_L_DVD: where DirichletBC(Q, g, ff, GAMMA_L_DVD) constrains."""
		sel = np.flatnonzero(self.ff == self.GAMMA_L_DVD)
		if sel.size == 0:
			return np.zeros(0, dtype=np.int64)
		cv = self.mesh.cells()[self.facet_cell[sel]]
		keep = np.ones(cv.shape, dtype=bool)
		keep[np.arange(sel.size), self.facet_local[sel]] = False                     # facet i is opposite local vertex i
		return np.unique(np.asarray(self.vertex_to_node)[cv[keep]])

	def init_beta(self, beta):
		print_text("::: initializing basal traction coefficient :::", cls=self)
		self.assign_variable(self.beta, beta)

	def init_A(self, A):
		print_text("::: initializing flow-rate factor over grounded and shelves :::", cls=self)
		self.assign_variable(self.A, A)

	def init_E(self, E):
		print_text("::: initializing enhancement factor over grounded and shelves :::", cls=self)
		self.assign_variable(self.E, E)

	def init_T(self, T):
		print_text("::: initializing absolute temperature :::", cls=self)
		self.assign_variable(self.T, T)

	def init_Tp(self, Tp):
		print_text("::: initializing pressure-adjusted temperature :::", cls=self)
		self.assign_variable(self.Tp, Tp)

	def init_W(self, W):
		print_text("::: initializing water content :::", cls=self)
		self.assign_variable(self.W, W)

	def form_energy_dependent_rate_factor(self):
		"""A = E a_T (1 + 181.25 min(W, 0.01)) exp(-Q_T / (R T')) as an expression of the P1
		fields T', W, E, evaluated at the quadrature points of the momentum forms instead of a
		nodal A (cslvr/model.py:1815-1859)."""
		print_text("::: formulating energy-dependent rate-factor :::", cls=self)
		self.energy_dependent_rate_factor = True

	def init_mask(self, mask):
		print_text("::: initializing shelf mask :::", cls=self)
		self.assign_variable(self.mask, mask)

	def init_sigma(self, sigma):
		self.assign_variable(self.sigma, sigma)

	def rate_factor(self, Tp, W=0.0, E=1.0):
		"""A(T', W) of ``Model.form_energy_dependent_rate_factor``
		(cslvr/model.py:1815-1859) evaluated nodally, for ``init_A``."""
		Tp = np.asarray(Tp, dtype=np.float64)
		a_T = np.where(Tp < 263.15, self.a_T_l, self.a_T_u)
		Q_T = np.where(Tp < 263.15, self.Q_T_l, self.Q_T_u)
		W_T = np.minimum(W, 0.01)
		return E * a_T * (1 + 181.25 * W_T) * np.exp(-Q_T / (self.R * Tp))

	# ---- cslvr/d3model.py:337-475 -----------------------------------------
	def calculate_boundaries(self, mask=None, S_ring=None, U_mask=None,
	                         mark_divide=False, contour=None):
		"""facet markers from the outward normal's z component (tol 1e-6), the shelf
		mask at the facet midpoint and the midpoint elevation; cell markers from the
		mask at the cell midpoint."""
		print_text("::: calculating boundaries :::", cls=self)
		self.mark_divide = mark_divide
		if (contour is None and mark_divide) or (contour is not None and not mark_divide):
			print_text(">>> IF PARAMETER <mark_divide> OF calculate_boundaries() IS "
			           "TRUE, PARAMETER <contour> MUST BE A NUMPY ARRAY OF COORDINATES "
			           "OF THE ICE-SHEET EXTERIOR BOUNDARY <<<", 'red', 1)
			sys.exit(1)
		tree = None
		if contour is not None and mark_divide:
			# cslvr/d3model.py:360-364: lateral facets further than 1 km from the ice-sheet contour are basin divides
			print_text("    - marking the interior facets for incomplete meshes -", cls=self)
			from scipy.spatial import cKDTree
			tree = cKDTree(np.asarray(contour, dtype=np.float64)[:, 0:2])
			self.contour_tree = tree
		self.init_mask(1.0 if mask is None else mask)
		X = self.mesh.coordinates()
		cells = self.mesh.cells()
		fc, fl = self.mesh.exterior_facets()
		fv, _, nrm, mid = facet_geometry(X, cells, fc, fl)
		if mask is None:
			m_f = np.ones(len(fc))
			m_c = np.ones(self.num_cells)
		elif callable(mask):
			m_f = np.asarray(mask(mid[:, 0], mid[:, 1], mid[:, 2])) * np.ones(len(fc))
			cm = X[cells].mean(axis=1)
			m_c = np.asarray(mask(cm[:, 0], cm[:, 1], cm[:, 2])) * np.ones(self.num_cells)
		else:
			mv = self.mask.compute_vertex_values()
			m_f = mv[fv].mean(axis=1)
			m_c = mv[cells].mean(axis=1)
		tol = 1e-6
		print_text("    - iterating through %i facets - " % self.mesh.num_facets(), cls=self)
		up, dn = nrm[:, 2] >= tol, nrm[:, 2] <= -tol
		side = ~(up | dn)
		flt = m_f > 1
		ff = np.zeros(len(fc), dtype=np.int32)
		ff[up & flt]    = self.GAMMA_U_FLT          # default U_mask = 1 everywhere
		ff[up & ~flt]   = self.GAMMA_U_GND
		ff[dn & flt]    = self.GAMMA_B_FLT
		ff[dn & ~flt]   = self.GAMMA_B_GND
		ff[side & (mid[:, 2] > 0)]  = self.GAMMA_L_OVR
		ff[side & ~(mid[:, 2] > 0)] = self.GAMMA_L_UDR
		if tree is not None:                         # cslvr/d3model.py:441-448
			far = ~(tree.query(mid[:, 0:2])[0] < 1000)
			ff[side & far] = self.GAMMA_L_DVD
		self.facet_cell, self.facet_local, self.ff = fc, fl, ff
		print_text("    - iterating through %i cells - " % self.num_cells, cls=self)
		self.cf = np.where(m_c > 1, self.OMEGA_FLT, self.OMEGA_GND).astype(np.int32)
		self.set_measures()

	# ---- cslvr/d3model.py:477-626 -----------------------------------------
	def set_measures(self):
		for name, mk in (('GAMMA_S_GND', 2), ('GAMMA_B_GND', 3), ('GAMMA_S_FLT', 6), ('GAMMA_B_FLT', 5),
		                 ('GAMMA_L_DVD', 7), ('GAMMA_L_OVR', 4), ('GAMMA_L_UDR', 10),
		                 ('GAMMA_U_GND', 8), ('GAMMA_U_FLT', 9)):
			setattr(self, 'N_' + name, int((self.ff == mk).sum()))
		self.N_OMEGA_GND = int((self.cf == self.OMEGA_GND).sum())
		self.N_OMEGA_FLT = int((self.cf == self.OMEGA_FLT).sum())
		B = Boundary
		self.dOmega    = B('dx', [self.OMEGA_GND, self.OMEGA_FLT], 'entire interior')
		self.dGamma_bg = B('ds', [self.GAMMA_B_GND], 'grounded basal surface')
		self.dGamma_bw = B('ds', [self.GAMMA_B_FLT], 'floating basal surface')
		self.dGamma_b  = B('ds', [self.GAMMA_B_GND, self.GAMMA_B_FLT], 'entire basal surface')
		self.dGamma_s  = B('ds', [self.GAMMA_S_GND, self.GAMMA_S_FLT, self.GAMMA_U_GND, self.GAMMA_U_FLT],
		                   'entire upper surface')
		self.dGamma_lt = B('ds', [self.GAMMA_L_OVR, self.GAMMA_L_UDR], 'entire exterior lateral surface')
		self.measures = {m.description: m for m in
		                 (self.dOmega, self.dGamma_bg, self.dGamma_bw, self.dGamma_b, self.dGamma_s, self.dGamma_lt)}

	# ---- cslvr/d3model.py:629-667 -----------------------------------------
	def deform_mesh_to_geometry(self, S, B):
		"""z <- (z / h) (S - B) + B per vertex, then re-mark the boundaries."""
		print_text("::: deforming mesh to geometry :::", cls=self)
		self.init_S(S)
		self.init_B(B)
		X = self.mesh.coordinates()
		zmin, zmax = X[:, 2].min(), X[:, 2].max()
		h = zmax - zmin
		self.init_sigma((X[:, 2] - zmin) / h)
		print_text("    - iterating through %i vertices - " % self.num_vertices, cls=self)
		X[:, 2] = (X[:, 2] / h) * (self.S.values - self.B.values)
		X[:, 2] = X[:, 2] + self.B.values
		self.calculate_boundaries()

	# ---- what the C-ABI consumes --------------------------------------------
	def vertex_field(self, f):
		"""vertex-indexed values of a coefficient Function (BP_FIELD_* layout)."""
		return np.ascontiguousarray(f.compute_vertex_values(), dtype=np.float64)
