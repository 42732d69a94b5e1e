"""numpy restatement of cslvr's Blatter-Pattyn momentum path (oracle).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- PARITY UNPINNED.

Everything here follows the reference by file:line (relative to /root/reference):

* mesh / periodic map / markers / deformation ........ cslvr/d3model.py
* strain rate, viscosity, dissipation ................ cslvr/momentum.py,
                                                       cslvr/momentumbp.py
* action, residual (both formulations), Jacobian ..... cslvr/momentumbp.py,
                                                       cslvr/physics.py
* Newton loop ......................................... cslvr/physics.py:664-678
                                                       (dolfin NewtonSolver, SURVEY A-N)

The element tensors are evaluated the way the reference's form compiler does it:
a loop over quadrature points with tabulated P1 basis values, *not* the closed
form the CUDA kernels use -- so agreement between the two is a real check.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .quadrature import tetrahedron_rule, triangle_rule

# marker ids, cslvr/d3model.py:16-27
OMEGA_GND, OMEGA_FLT = 0, 1
GAMMA_S_GND, GAMMA_B_GND, GAMMA_S_FLT, GAMMA_B_FLT = 2, 3, 6, 5
GAMMA_L_DVD, GAMMA_L_OVR, GAMMA_L_UDR = 7, 4, 10
GAMMA_U_GND, GAMMA_U_FLT = 8, 9

DOLFIN_EPS = 3.0e-16        # SURVEY A-C

# physical constants, cslvr/model.py:212-236
CONSTANTS = dict(eps_reg=1e-15, n=3.0, rho_i=910.0, rhosw=1028.0, g=9.80665)


# --------------------------------------------------------------------------
# mesh (third-party dolfin BoxMesh, SURVEY A-M; counts pinned by
# simulations/verification/ismip_hom_a.py:5: BoxMesh(10,10,2) = 1200 cells,
# 2680 facets, 363 vertices)
# --------------------------------------------------------------------------
def box_mesh(p1, p2, nx, ny, nz):
	"""dolfin ``BoxMesh(Point(*p1), Point(*p2), nx, ny, nz)`` (docs/get_started.rst:15-18)."""
	x = p1[0] + np.arange(nx + 1) * ((p2[0] - p1[0]) / nx)
	y = p1[1] + np.arange(ny + 1) * ((p2[1] - p1[1]) / ny)
	z = p1[2] + np.arange(nz + 1) * ((p2[2] - p1[2]) / nz)
	Z, Y, X = np.meshgrid(z, y, x, indexing='ij')
	coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
	iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing='ij')
	v0 = (iz * (ny + 1) * (nx + 1) + iy * (nx + 1) + ix).ravel()
	v1 = v0 + 1
	v2 = v0 + (nx + 1)
	v3 = v1 + (nx + 1)
	v4 = v0 + (nx + 1) * (ny + 1)
	v5 = v1 + (nx + 1) * (ny + 1)
	v6 = v2 + (nx + 1) * (ny + 1)
	v7 = v3 + (nx + 1) * (ny + 1)
	tets = np.stack([np.stack(t, axis=1) for t in (
		(v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4),
		(v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7))], axis=1)
	cells = np.sort(tets.reshape(-1, 4), axis=1)       # mesh.order()
	return coords, cells.astype(np.int64)


def periodic_vertex_map(coords):
	"""vertex -> master vertex under the 3-D periodic SubDomain, d3model.py:62-108.

	x = xmax maps to xmin, y = ymax to ymin, the corner to the origin, z is
	unchanged; must be evaluated on the FLAT mesh (model.py:473-475).
	"""
	xmin, xmax = coords[:, 0].min(), coords[:, 0].max()
	ymin, ymax = coords[:, 1].min(), coords[:, 1].max()
	tgt = coords.copy()
	on_x = np.abs(coords[:, 0] - xmax) <= DOLFIN_EPS * max(1.0, abs(xmax))
	on_y = np.abs(coords[:, 1] - ymax) <= DOLFIN_EPS * max(1.0, abs(ymax))
	tgt[on_x, 0] -= (xmax - xmin)
	tgt[on_y, 1] -= (ymax - ymin)
	# match target coordinates to vertices (exact on a structured mesh up to rounding)
	scale = np.array([xmax - xmin, ymax - ymin, max(np.ptp(coords[:, 2]), 1e-300)])
	def key(c):
		q = np.round((c - np.array([xmin, ymin, coords[:, 2].min()])) / scale * 2 ** 20).astype(np.int64)
		return (q[:, 0] * (2 ** 21) + q[:, 1]) * (2 ** 21) + q[:, 2]
	kv = key(coords)
	order = np.argsort(kv)
	pos = np.searchsorted(kv[order], key(tgt))
	master = order[pos]
	assert np.all(kv[master] == key(tgt)), "periodic image vertex not found"
	return master


def nodes_from_master(master):
	"""compact node numbering of the master vertices, in vertex order."""
	is_master = master == np.arange(master.size)
	node_of_master = np.cumsum(is_master) - 1
	return node_of_master[master].astype(np.int64), int(is_master.sum())


def exterior_facets(cells):
	"""(cell, local facet) of every facet that belongs to exactly one cell.

	Facet i of a tet is the one opposite local vertex i (UFC, SURVEY A-D).
	"""
	nt = cells.shape[0]
	loc = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
	f = np.sort(cells[:, loc].reshape(-1, 3), axis=1)
	order = np.lexsort((f[:, 2], f[:, 1], f[:, 0]))
	fs = f[order]
	same_next = np.zeros(fs.shape[0], dtype=bool)
	same_next[:-1] = np.all(fs[1:] == fs[:-1], axis=1)
	same_prev = np.zeros_like(same_next)
	same_prev[1:] = same_next[:-1]
	ext = order[~(same_next | same_prev)]
	ext.sort()
	return (ext // 4).astype(np.int64), (ext % 4).astype(np.int64)


def facet_geometry(coords, cells, fc, fl):
	"""area, outward unit normal and midpoint of facets (cell fc, local facet fl)."""
	loc = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
	fv = cells[fc][np.arange(fc.size)[:, None], loc[fl]]
	p = coords[fv]                                     # [nf,3,3]
	cr = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
	area = 0.5 * np.linalg.norm(cr, axis=1)
	nrm = cr / np.linalg.norm(cr, axis=1)[:, None]
	opp = coords[cells[fc, fl]]
	flip = np.einsum('ij,ij->i', nrm, opp - p[:, 0]) > 0
	nrm[flip] *= -1.0
	return fv, area, nrm, p.mean(axis=1)


def mark_facets(coords, cells, fc, fl, mask=None, contour=None):
	"""facet markers of ``D3Model.calculate_boundaries``, d3model.py:399-455.

	``mask`` is an optional callable (x, y, z) -> value evaluated at the facet
	midpoint (>1 means floating); default: all grounded (d3model.py:372-373).
	Upper-surface facets get GAMMA_U_GND/FLT because the default U_mask is 1
	(d3model.py:384-385, 423-432).
	"""
	tol = 1e-6
	_, _, nrm, mid = facet_geometry(coords, cells, fc, fl)
	m = np.ones(fc.size) if mask is None else np.asarray(mask(mid[:, 0], mid[:, 1], mid[:, 2])) * np.ones(fc.size)
	ff = np.zeros(fc.size, dtype=np.int64)
	up = nrm[:, 2] >= tol
	dn = nrm[:, 2] <= -tol
	sd = ~(up | dn)
	ff[up & (m > 1)] = GAMMA_U_FLT
	ff[up & ~(m > 1)] = GAMMA_U_GND
	ff[dn & (m > 1)] = GAMMA_B_FLT
	ff[dn & ~(m > 1)] = GAMMA_B_GND
	ff[sd & (mid[:, 2] > 0)] = GAMMA_L_OVR
	ff[sd & ~(mid[:, 2] > 0)] = GAMMA_L_UDR
	if contour is not None:
		# mark_divide=True (d3model.py:441-448): lateral facets whose midpoint is not within 1000 m of a point
		# of the ice-sheet contour are basin divides (brute-force nearest point; the reference uses a cKDTree)
		c2 = np.asarray(contour, dtype=np.float64)[:, 0:2]
		for i in np.flatnonzero(sd):
			if not (np.sqrt(((c2 - mid[i, 0:2]) ** 2).sum(axis=1)).min() < 1000):
				ff[i] = GAMMA_L_DVD
	return ff


def deform(coords, S_v, B_v):
	"""``D3Model.deform_mesh_to_geometry``, d3model.py:651-662 (per vertex)."""
	z = coords[:, 2]
	h = z.max() - z.min()
	out = coords.copy()
	out[:, 2] = (z / h) * (S_v - B_v)
	out[:, 2] = out[:, 2] + B_v
	return out


def canonical_cell_dofs(cells, v2n):
	"""UFC mixed P1xP1 cell dofs, component-major (SURVEY A-D), dof = 2*node + c."""
	nd = v2n[cells]
	return np.concatenate([2 * nd, 2 * nd + 1], axis=1).astype(np.int64)


def csr_pattern(cell_dofs, n_dofs):
	"""dolfin SparsityPatternBuilder semantics (SURVEY A-S): union over cells of
	cell_dofs x cell_dofs, rows sorted by column, explicit zeros kept."""
	nt, k = cell_dofs.shape
	half = k // 2
	if k == 8 and np.array_equal(cell_dofs[:, half:], cell_dofs[:, :half] + 1) and not np.any(cell_dofs[:, :half] & 1):
		# blocked numbering dof = 2*node + c: build the node pattern, expand 2x2 (same result, 1/4 the keys)
		ip, ix = csr_pattern(cell_dofs[:, :half] // 2, n_dofs // 2)
		cnt = np.diff(ip)
		cols = np.repeat(2 * ix.astype(np.int64), 2) + np.tile(np.array([0, 1]), ix.size)
		# dof rows 2i and 2i+1 both carry node row i's expanded column list
		r = np.repeat(np.arange(cnt.size), 4 * cnt)
		w = np.arange(4 * ix.size) - np.repeat(4 * ip[:-1], 4 * cnt)
		indices = cols[2 * ip[:-1][r] + w % np.maximum(2 * cnt[r], 1)].astype(np.int32)
		indptr = np.zeros(n_dofs + 1, dtype=np.int64)
		indptr[1:] = np.cumsum(np.repeat(2 * cnt, 2))
		return indptr, indices
	rows = np.repeat(cell_dofs, k, axis=1).ravel()
	cols = np.tile(cell_dofs, (1, k)).ravel()
	key = np.unique(rows * np.int64(n_dofs) + cols)
	r = key // n_dofs
	c = key % n_dofs
	indptr = np.zeros(n_dofs + 1, dtype=np.int64)
	np.add.at(indptr, r + 1, 1)
	return np.cumsum(indptr), c.astype(np.int32)


# --------------------------------------------------------------------------
# the problem
# --------------------------------------------------------------------------
class BPProblem(object):
	"""One ``MomentumDukowiczBP`` / ``MomentumBP`` problem on a tetrahedral mesh.

	coords[N_v,3], cells[N_t,4] (sorted ascending per cell), cell_dofs[N_t,8]
	(component-major), exterior facets (fc, fl, ff) and vertex-indexed
	coefficient arrays S, A, beta (``D = -min(0,z)`` is derived,
	d3model.py:858-861).
	"""

	def __init__(self, coords, cells, cell_dofs, n_dofs, fc, fl, ff, S, A, beta,
	             periodic, use_pressure_bc=True, quad_degree=None, consts=None, energy=None):
		self.coords = np.asarray(coords, dtype=np.float64)
		self.cells = np.asarray(cells, dtype=np.int64)
		self.cell_dofs = np.asarray(cell_dofs, dtype=np.int64)
		self.n_dofs = int(n_dofs)
		self.fc, self.fl, self.ff = (np.asarray(a, dtype=np.int64) for a in (fc, fl, ff))
		nv = self.coords.shape[0]
		self.S = np.asarray(S, dtype=np.float64) * np.ones(nv)
		self.A = np.asarray(A, dtype=np.float64) * np.ones(nv)
		self.beta = np.asarray(beta, dtype=np.float64) * np.ones(nv)
		self.D = -np.minimum(0.0, self.coords[:, 2])
		self.periodic = bool(periodic)
		self.use_pressure_bc = bool(use_pressure_bc)
		c = dict(CONSTANTS)
		c.update(consts or {})
		self.c = c
		# UFL degree estimation, SURVEY A-Q: P1 A -> 5; constant A is still a
		# P1 Function in cslvr (model.py:2530) so the estimate does not change
		# energy = dict(Tp=, W=, E=) switches to Model.form_energy_dependent_rate_factor
		# (model.py:1815-1859): A is then a UFL expression evaluated at the quadrature points and
		# the estimated degree rises to 9 (SURVEY A-Q)
		self.energy = None
		if energy is not None:
			self.energy = {k: np.asarray(energy[k], dtype=np.float64) * np.ones(nv) for k in ('Tp', 'W', 'E')}
		self.quad_degree = (9 if energy is not None else 5) if quad_degree is None else int(quad_degree)
		self._geometry()
		self._pattern = None

	# ---- geometry --------------------------------------------------------
	def _geometry(self):
		X = self.coords[self.cells]                       # [nt,4,3]
		Jm = X[:, 1:, :] - X[:, :1, :]                     # rows = edges
		Ji = np.linalg.inv(Jm)                             # columns = grad lambda_1..3
		G = np.empty((self.cells.shape[0], 4, 3))
		G[:, 1:, :] = np.transpose(Ji, (0, 2, 1))
		G[:, 0, :] = -G[:, 1:, :].sum(axis=1)
		self.G = G                                         # grad lambda_a [nt,4,3]
		self.vol = np.abs(np.linalg.det(Jm)) / 6.0

	# ---- pointwise physics ----------------------------------------------
	def _grad_u(self, u):
		"""g[t,c,j] = d u_c / d x_j, cell-constant for P1."""
		U = u[self.cell_dofs].reshape(-1, 2, 4)
		return np.einsum('tca,taj->tcj', U, self.G)

	@staticmethod
	def _epsdot(g):
		"""effective strain rate squared of the BP tensor,
		momentumbp.py:122-135 + momentum.py:245-251 (0.5*tr(eps.eps))."""
		e00 = g[:, 0, 0]
		e11 = g[:, 1, 1]
		e01 = 0.5 * (g[:, 0, 1] + g[:, 1, 0])
		e02 = 0.5 * g[:, 0, 2]
		e12 = 0.5 * g[:, 1, 2]
		e22 = -e00 - e11
		return 0.5 * (e00 ** 2 + e11 ** 2 + e22 ** 2) + e01 ** 2 + e02 ** 2 + e12 ** 2

	def _depsdot(self, g):
		"""d(epsdot)/d(u_{c,a}) for the 8 local dofs, [nt,8]."""
		G = self.G
		e00 = g[:, 0, 0]
		e11 = g[:, 1, 1]
		e01 = 0.5 * (g[:, 0, 1] + g[:, 1, 0])
		e02 = 0.5 * g[:, 0, 2]
		e12 = 0.5 * g[:, 1, 2]
		e22 = -e00 - e11
		# d/du_x,a : e00 -> Gx, e22 -> -Gx, e01 -> .5 Gy, e02 -> .5 Gz
		dx = (e00 - e22)[:, None] * G[:, :, 0] + (2 * e01 * 0.5)[:, None] * G[:, :, 1] \
		     + (2 * e02 * 0.5)[:, None] * G[:, :, 2]
		dy = (e11 - e22)[:, None] * G[:, :, 1] + (2 * e01 * 0.5)[:, None] * G[:, :, 0] \
		     + (2 * e12 * 0.5)[:, None] * G[:, :, 2]
		return np.concatenate([dx, dy], axis=1)

	def _d2epsdot(self):
		"""second derivative of epsdot (constant in u), [nt,8,8]."""
		G = self.G
		xx = np.einsum('ta,tb->tab', G[:, :, 0], G[:, :, 0])
		yy = np.einsum('ta,tb->tab', G[:, :, 1], G[:, :, 1])
		zz = np.einsum('ta,tb->tab', G[:, :, 2], G[:, :, 2])
		xy = np.einsum('ta,tb->tab', G[:, :, 0], G[:, :, 1])
		H = np.empty((G.shape[0], 8, 8))
		H[:, :4, :4] = 2 * xx + 0.5 * yy + 0.5 * zz
		H[:, 4:, 4:] = 2 * yy + 0.5 * xx + 0.5 * zz
		H[:, :4, 4:] = xy + 0.5 * np.transpose(xy, (0, 2, 1))
		H[:, 4:, :4] = np.transpose(H[:, :4, 4:], (0, 2, 1))
		return H

	def _A_at(self, lam, Av):
		if self.energy is None:
			return Av @ lam                                # P1 interpolant of model.A
		spy = 31556926.0                                   # model.py:190, 275-285
		a_T_l, a_T_u, Q_T_l, Q_T_u, R = 3.985e-13 * spy, 1.916e3 * spy, 6e4, 13.9e4, 8.3144621
		Tp = self.energy['Tp'][self.cells] @ lam
		W = self.energy['W'][self.cells] @ lam
		E = self.energy['E'][self.cells] @ lam
		a_T = np.where(Tp < 263.15, a_T_l, a_T_u)
		Q_T = np.where(Tp < 263.15, Q_T_l, Q_T_u)
		W_T = np.where(W < 0.01, W, 0.01)
		return E * a_T * (1 + 181.25 * W_T) * np.exp(-Q_T / (R * Tp))

	# ---- cell integrals ---------------------------------------------------
	def cell_tensors(self, u, want_J=True, form='dukowicz', picard=False):
		"""element action, residual [nt,8] and Jacobian [nt,8,8] by quadrature.

		Dukowicz: momentumbp.py:460-500 with Vd from momentum.py:187-222;
		classic: momentumbp.py:350-375 with eta from momentum.py:159-185 and
		sigma from momentumbp.py:137-154.  Jacobian: physics.py:75-77.
		"""
		n, eps0 = self.c['n'], self.c['eps_reg']
		rho_g = self.c['rho_i'] * self.c['g']
		pts, wts = tetrahedron_rule(self.quad_degree)
		g = self._grad_u(u)
		ed = self._epsdot(g) + eps0
		Av = self.A[self.cells]                            # [nt,4]
		Uloc = u[self.cell_dofs].reshape(-1, 2, 4)
		gS = np.einsum('ta,taj->tj', self.S[self.cells], self.G)   # grad S
		nt = self.cells.shape[0]
		act = np.zeros(nt)
		F = np.zeros((nt, 8))
		J = np.zeros((nt, 8, 8)) if want_J else None
		de = self._depsdot(g)
		if want_J:
			H = self._d2epsdot()
			dede = np.einsum('ti,tj->tij', de, de)
		if form == 'classic':
			# sigma = 2 eta * quasi-strain (2x3), momentumbp.py:137-154
			q = np.empty((nt, 2, 3))
			q[:, 0, 0] = 2 * g[:, 0, 0] + g[:, 1, 1]
			q[:, 0, 1] = 0.5 * (g[:, 0, 1] + g[:, 1, 0])
			q[:, 0, 2] = 0.5 * g[:, 0, 2]
			q[:, 1, 0] = q[:, 0, 1]
			q[:, 1, 1] = 2 * g[:, 1, 1] + g[:, 0, 0]
			q[:, 1, 2] = 0.5 * g[:, 1, 2]
		for lam, w in zip(pts, wts):
			dx = w * self.vol
			Aq = self._A_at(lam, Av)                       # A at the point
			uq = Uloc @ lam                                # [nt,2]
			Ainv = Aq ** (-1.0 / n)
			act += dx * ((2 * n) / (n + 1) * Ainv * ed ** ((n + 1) / (2 * n))
			             + rho_g * (uq[:, 0] * gS[:, 0] + uq[:, 1] * gS[:, 1]))
			two_eta = Ainv * ed ** ((1 - n) / (2 * n))        # = 2*eta, momentum.py:185
			if form == 'classic':
				# inner(sigma, grad v): v = e_c lambda_a -> sigma[c,:].grad lambda_a
				Fv = np.einsum('tcj,taj->tca', two_eta[:, None, None] * q, self.G).reshape(nt, 8)
			else:
				Fv = two_eta[:, None] * de
			grav = np.concatenate([np.outer(gS[:, 0], lam), np.outer(gS[:, 1], lam)], axis=1)
			F += dx[:, None] * (Fv + rho_g * grav)
			if want_J:
				two_eta_p = 0.0 * Ainv if picard else Ainv * ((1 - n) / (2 * n)) * ed ** ((1 - n) / (2 * n) - 1)
				J += dx[:, None, None] * (two_eta[:, None, None] * H + two_eta_p[:, None, None] * dede)
		return act, F, J

	# ---- facet integrals --------------------------------------------------
	def _facet_sets(self):
		basal = np.isin(self.ff, (GAMMA_B_GND, GAMMA_B_FLT))      # dGamma_b, d3model.py:578-579
		press = self.ff == GAMMA_B_FLT                             # dGamma_bw, d3model.py:576-577
		if (not self.periodic) and self.use_pressure_bc:          # momentumbp.py:488-491
			press = press | np.isin(self.ff, (GAMMA_L_OVR, GAMMA_L_UDR))   # dGamma_lt
		return basal, press

	def facet_tensors(self, u):
		"""per-facet action, residual [nf,8] and Jacobian [nf,8,8] (cell-local
		dof indexing, zero rows for the vertex opposite the facet)."""
		rho_g = self.c['rho_i'] * self.c['g']
		rsw_g = self.c['rhosw'] * self.c['g']
		nf = self.fc.size
		act = np.zeros(nf)
		F = np.zeros((nf, 8))
		J = np.zeros((nf, 8, 8))
		if nf == 0:
			return act, F, J
		basal, press = self._facet_sets()
		fv, area, nrm, _ = facet_geometry(self.coords, self.cells, self.fc, self.fl)
		loc = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])[self.fl]   # [nf,3] local ids
		Uc = u[self.cell_dofs[self.fc]].reshape(-1, 2, 4)
		Uf = np.take_along_axis(Uc, loc[:, None, :], axis=2)    # [nf,2,3]
		bf = self.beta[fv]
		fw = rho_g * (self.S[fv] - self.coords[fv, 2]) - rsw_g * self.D[fv]   # momentumbp.py:477
		rows = np.arange(nf)[:, None]
		# sliding: -Sl = +0.5 beta u.u, quadrature degree 3 (SURVEY A-Q)
		pts, wts = triangle_rule(3)
		for lam, w in zip(pts, wts):
			dx = w * area * basal
			bq = bf @ lam
			uq = Uf @ lam
			act += dx * 0.5 * bq * (uq ** 2).sum(axis=1)
			for c in range(2):
				for a in range(3):
					F[rows[:, 0], 4 * c + loc[:, a]] += dx * bq * uq[:, c] * lam[a]
					for b in range(3):
						J[rows[:, 0], 4 * c + loc[:, a], 4 * c + loc[:, b]] += dx * bq * lam[a] * lam[b]
		# water pressure: -Pb = -f_w u.n_h, degree 2
		pts, wts = triangle_rule(2)
		for lam, w in zip(pts, wts):
			dx = w * area * press
			fq = fw @ lam
			uq = Uf @ lam
			act -= dx * fq * (uq[:, 0] * nrm[:, 0] + uq[:, 1] * nrm[:, 1])
			for c in range(2):
				for a in range(3):
					F[rows[:, 0], 4 * c + loc[:, a]] -= dx * fq * nrm[:, c] * lam[a]
		return act, F, J

	# ---- global assembly ---------------------------------------------------
	def pattern(self):
		if self._pattern is None:
			self._pattern = csr_pattern(self.cell_dofs, self.n_dofs)
		return self._pattern

	def action(self, u):
		a, _, _ = self.cell_tensors(u, want_J=False)
		af, _, _ = self.facet_tensors(u)
		return a.sum() + af.sum()

	def residual(self, u, form='dukowicz'):
		_, F, _ = self.cell_tensors(u, want_J=False, form=form)
		out = np.zeros(self.n_dofs)
		np.add.at(out, self.cell_dofs.ravel(), F.ravel())
		if self.fc.size:
			_, Ff, _ = self.facet_tensors(u)
			np.add.at(out, self.cell_dofs[self.fc].ravel(), Ff.ravel())
		return out

	def jacobian(self, u):
		"""CSR values aligned with :meth:`pattern` (explicit zeros kept)."""
		_, _, J = self.cell_tensors(u, want_J=True)
		indptr, indices = self.pattern()
		nd = self.n_dofs
		rows = np.repeat(self.cell_dofs, 8, axis=1).ravel()
		cols = np.tile(self.cell_dofs, (1, 8)).ravel()
		vals = J.ravel()
		if self.fc.size:
			_, _, Jf = self.facet_tensors(u)
			cd = self.cell_dofs[self.fc]
			rows = np.concatenate([rows, np.repeat(cd, 8, axis=1).ravel()])
			cols = np.concatenate([cols, np.tile(cd, (1, 8)).ravel()])
			vals = np.concatenate([vals, Jf.ravel()])
		M = sp.coo_matrix((vals, (rows, cols)), shape=(nd, nd)).tocsr()
		M.sort_indices()
		# re-express on the full pattern (facet zeros never add entries; cell
		# zeros are kept by coo->csr summation only if present, so map explicitly)
		out = np.zeros(indices.size)
		full_key = np.repeat(np.arange(nd, dtype=np.int64), np.diff(indptr)) * nd + indices
		m_key = np.repeat(np.arange(nd, dtype=np.int64), np.diff(M.indptr)) * nd + M.indices
		out[np.searchsorted(full_key, m_key)] = M.data
		return out

	def jacobian_matrix(self, u):
		indptr, indices = self.pattern()
		return sp.csr_matrix((self.jacobian(u), indices, indptr), shape=(self.n_dofs,) * 2)

	# ---- solve_pressure, cslvr/momentumbp.py:205-229 ---------------------------
	def scalar_pattern(self):
		"""node-level (model.Q) CSR pattern and per-cell node ids; Q dof = node = cell_dofs[:, :4] // 2
		for the canonical blocked numbering."""
		nd = self.cell_dofs[:, :4] // 2
		return csr_pattern(nd, self.n_dofs // 2), nd

	def mass_matrix(self):
		"""assemble(u*v*dx) on model.Q (degree-2 rule is exact: |T| (1 + delta_ab) / 20)."""
		(ip, ix), nd = self.scalar_pattern()
		nn = self.n_dofs // 2
		Me = (self.vol[:, None, None] / 20.0) * (np.ones((4, 4)) + np.eye(4))[None]
		M = sp.coo_matrix((Me.ravel(), (np.repeat(nd, 4, axis=1).ravel(), np.tile(nd, (1, 4)).ravel())), shape=(nn, nn)).tocsr()
		return M

	def pressure(self, u, quad_degree=6):
		"""p = project(rho g (S - z) - 2 eta (u_x,x + u_y,y), Q) with eta = eta(model.u)
		(momentumbp.py:216-223, Appendix B quirk: always the 'linear' viscosity of the
		velocity just solved).  UFL degree estimate: 1 (v) + 3 (A^(-1/n)) + 2 ((.)^p) = 6
		-> Keast 24-point rule; dolfin's project() solves the mass system directly."""
		n, eps0 = self.c['n'], self.c['eps_reg']
		rho_g = self.c['rho_i'] * self.c['g']
		pts, wts = tetrahedron_rule(quad_degree)
		g = self._grad_u(u)
		ed = self._epsdot(g) + eps0
		div = g[:, 0, 0] + g[:, 1, 1]
		Av = self.A[self.cells]
		Sz = self.S[self.cells] - self.coords[self.cells, 2]
		b = np.zeros((self.cells.shape[0], 4))
		for lam, w in zip(pts, wts):
			dx = w * self.vol
			eta = 0.5 * (Av @ lam) ** (-1.0 / n) * ed ** ((1 - n) / (2 * n))
			pf = rho_g * (Sz @ lam) - 2.0 * eta * div
			b += (dx * pf)[:, None] * lam[None, :]
		(_, _), nd = self.scalar_pattern()
		rhs = np.zeros(self.n_dofs // 2)
		np.add.at(rhs, nd.ravel(), b.ravel())
		return spla.spsolve(self.mass_matrix().tocsc(), rhs)

	# ---- solve_vert_velocity, cslvr/momentumbp.py:60-85, 231-253 ----------------
	def vert_velocity(self, u, B, B_ring=None, lateral_bc=None):
		"""u_z from the continuity equation the way the reference does it:

		1. u_z_b = project((-B_ring - u_x n_b0 - u_y n_b1) / n_b2, Q) with
		   n_b = (grad B - k) / |grad B - k|  (momentumbp.py:68-72, 240);
		2. A_ij = int d(phi_j)/dz phi_i dx,  b_i = -int (u_x,x + u_y,y) phi_i dx
		   (lhs/rhs of u_z_F, momentumbp.py:75, 244-245);
		3. DirichletBC(Q, u_z_b, ff, GAMMA_B_GND / GAMMA_B_FLT) applied to (A, b): rows of the
		   basal dofs become identity rows (momentumbp.py:79-85, 246-248);
		4. direct solve (LUSolver('mumps'), momentumbp.py:249-250).
		Returns (u_z, u_z_b), both per node of model.Q."""
		(ip, ix), nd = self.scalar_pattern()
		nn = self.n_dofs // 2
		nt = self.cells.shape[0]
		B = np.asarray(B, dtype=np.float64)
		Br = np.zeros_like(B) if B_ring is None else np.asarray(B_ring, dtype=np.float64) * np.ones_like(B)
		gB = np.einsum('ta,taj->tj', B[self.cells], self.G)
		mag = np.sqrt(1.0 + (gB ** 2).sum(axis=1))
		nb = np.stack([gB[:, 0], gB[:, 1], gB[:, 2] - 1.0], axis=1) / mag[:, None]
		Uloc = u[self.cell_dofs].reshape(-1, 2, 4)
		M = self.mass_matrix()
		Me = (self.vol[:, None, None] / 20.0) * (np.ones((4, 4)) + np.eye(4))[None]
		# nodal values of the (cell-wise) integrand: P1 in u and B_ring, cell-constant normal
		f = (-Br[self.cells] - Uloc[:, 0, :] * nb[:, 0:1] - Uloc[:, 1, :] * nb[:, 1:2]) / nb[:, 2:3]
		rhs = np.zeros(nn)
		np.add.at(rhs, nd.ravel(), np.einsum('tab,tb->ta', Me, f).ravel())
		uzb = spla.spsolve(M.tocsc(), rhs)
		# the continuity system
		g = self._grad_u(u)
		div = g[:, 0, 0] + g[:, 1, 1]
		Ae = (self.vol / 4.0)[:, None, None] * np.broadcast_to(self.G[:, None, :, 2], (nt, 4, 4))   # [t, i, j] = dz(phi_j) |T|/4
		A = sp.coo_matrix((Ae.ravel(), (np.repeat(nd, 4, axis=1).ravel(), np.tile(nd, (1, 4)).ravel())), shape=(nn, nn)).tolil()
		b = np.zeros(nn)
		np.add.at(b, nd.ravel(), np.repeat(-div * self.vol / 4.0, 4))
		loc = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
		basal = np.isin(self.ff, (GAMMA_B_GND, GAMMA_B_FLT))
		bn = np.unique(nd[self.fc[basal]][np.arange(basal.sum())[:, None], loc[self.fl[basal]]])
		for i in bn:
			A.rows[i] = [i]
			A.data[i] = [1.0]
		b[bn] = uzb[bn]
		if lateral_bc is not None:                             # DirichletBC(Q, u_z_lat, ff, GAMMA_L_DVD), momentumbp.py:110-111: applied last
			for i, g_ in zip(np.asarray(lateral_bc[0], dtype=np.int64), np.asarray(lateral_bc[1], dtype=np.float64)):
				A.rows[i] = [i]
				A.data[i] = [1.0]
				b[i] = g_
		return spla.spsolve(A.tocsc(), b), uzb

	# ---- linear=True, cslvr/momentumbp.py:396 / 503, physics.py:655-661 ----------
	def linear_solve(self, u_l):
		"""one linear solve with the viscosity frozen at eta(u_l) (``linear=True``:
		Vd = 2 eta(model.u) epsdot(u), momentumbp.py:91-93 / momentum.py:213-216; the
		residual with u_h replaced by the trial function splits into lhs == rhs).
		K = second variation without the eta' term, f = the u-independent part of F."""
		_, Fc, Jc = self.cell_tensors(u_l, want_J=True, picard=True)
		indptr, indices = self.pattern()
		nd = self.n_dofs
		rows = np.repeat(self.cell_dofs, 8, axis=1).ravel()
		cols = np.tile(self.cell_dofs, (1, 8)).ravel()
		vals = Jc.ravel()
		zero = np.zeros(nd)
		f = self.residual(zero) - self._viscous_residual(zero)        # gravity + water pressure only
		if self.fc.size:
			_, _, Jf = self.facet_tensors(u_l)
			cd = self.cell_dofs[self.fc]
			rows = np.concatenate([rows, np.repeat(cd, 8, axis=1).ravel()])
			cols = np.concatenate([cols, np.tile(cd, (1, 8)).ravel()])
			vals = np.concatenate([vals, Jf.ravel()])
		K = sp.coo_matrix((vals, (rows, cols)), shape=(nd, nd)).tocsc()
		return spla.spsolve(K, -f)

	def _viscous_residual(self, u):
		"""int 2 eta(u) d(epsdot)(u; v): the part of F that is not gravity / facets."""
		n, eps0 = self.c['n'], self.c['eps_reg']
		pts, wts = tetrahedron_rule(self.quad_degree)
		g = self._grad_u(u)
		ed = self._epsdot(g) + eps0
		de = self._depsdot(g)
		Av = self.A[self.cells]
		F = np.zeros((self.cells.shape[0], 8))
		for lam, w in zip(pts, wts):
			F += (w * self.vol * (Av @ lam) ** (-1.0 / n) * ed ** ((1 - n) / (2 * n)))[:, None] * de
		out = np.zeros(self.n_dofs)
		np.add.at(out, self.cell_dofs.ravel(), F.ravel())
		return out

	# ---- Newton (dolfin NewtonSolver semantics, SURVEY A-N) ----------------
	def newton(self, u0=None, rtol=1e-9, atol=1e-10, maxit=25, relax=1.0,
	           linear_solver='direct', krylov_rtol=1e-5, verbose=False, bcs=None):
		"""physics.py:664-678 -> dolfin.solve(F == 0, u, J=J, bcs=bcs).

		Cold start from DOLFIN_EPS (momentumbp.py:265-266).  ``linear_solver``:
		'direct' (sparse LU, the parity setting) or 'cg' (Jacobi-PCG with a
		PETSc-like relative tolerance on the preconditioned residual).
		``bcs = (dofs, values)``: the DirichletBCs of momentumbp.py:101-112 the way dolfin's
		NewtonSolver applies them (SURVEY A-N): b_i = u_i - g_i on the constrained rows, those
		rows of J replaced by identity rows -- so the update carries u_i to g_i.
		"""
		u = np.full(self.n_dofs, DOLFIN_EPS) if u0 is None else np.array(u0, dtype=np.float64)
		if bcs is not None:
			bd, bg = np.asarray(bcs[0], dtype=np.int64), np.asarray(bcs[1], dtype=np.float64)
			free = np.ones(self.n_dofs); free[bd] = 0.0
			_res, _jac = self.residual, self.jacobian_matrix

			def residual(v):
				r = _res(v); r[bd] = v[bd] - bg
				return r

			def jacobian_matrix(v):
				Jm = sp.diags(free) @ _jac(v).tocsr()                 # bc.apply(A): the rows only
				return (Jm + sp.diags(1.0 - free)).tocsr()
			return self._newton_core(u, residual, jacobian_matrix, rtol, atol, maxit, relax, linear_solver, krylov_rtol, verbose)
		return self._newton_core(u, self.residual, self.jacobian_matrix, rtol, atol, maxit, relax, linear_solver, krylov_rtol, verbose)

	def _newton_core(self, u, residual, jacobian_matrix, rtol, atol, maxit, relax, linear_solver, krylov_rtol, verbose):
		b = residual(u)
		r0 = np.linalg.norm(b)
		hist = [r0]
		it = 0
		kits = 0
		conv = (r0 < atol)
		while not conv and it < maxit:
			Jm = jacobian_matrix(u)
			if linear_solver == 'direct':
				dx = spla.splu(Jm.tocsc()).solve(b)
			else:
				d = Jm.diagonal()
				Mop = spla.LinearOperator(Jm.shape, matvec=lambda x: x / d)
				cnt = [0]
				dx, info = spla.cg(Jm, b, rtol=krylov_rtol, atol=0.0, M=Mop, maxiter=10000,
				                   callback=lambda xk: cnt.__setitem__(0, cnt[0] + 1))
				kits += cnt[0]
			u = u - relax * dx
			b = residual(u)
			r = np.linalg.norm(b)
			hist.append(r)
			it += 1
			if verbose:
				print("Newton iteration %d: r (abs) = %.3e (tol = %.3e) r (rel) = %.3e (tol = %.3e)"
				      % (it, r, atol, r / r0, rtol))
			conv = (r / r0 < rtol) or (r < atol)
		return u, dict(converged=bool(conv), iterations=it, residuals=hist, krylov_iterations=kits)


# --------------------------------------------------------------------------
# convenience: the BASELINE.json configurations on BoxMesh
# --------------------------------------------------------------------------
def ismip_hom(exp='A', L=80000.0, nx=15, ny=15, nz=5, periodic=True, quad_degree=None):
	"""ISMIP-HOM A (docs/get_started.rst:13-45) or C
	(simulations/ismip_hom/ismip_hom_c.py:4-25) on a BoxMesh."""
	coords, cells = box_mesh((0.0, 0.0, 0.0), (L, L, 1.0), nx, ny, nz)
	if periodic:
		v2n, n_nodes = nodes_from_master(periodic_vertex_map(coords))
	else:
		v2n, n_nodes = np.arange(coords.shape[0], dtype=np.int64), coords.shape[0]
	x, y = coords[:, 0], coords[:, 1]
	if exp == 'A':
		a = 0.5 * np.pi / 180
		S = -x * np.tan(a)
		B = S - 1000.0 + 500.0 * np.sin(2 * np.pi * x / L) * np.sin(2 * np.pi * y / L)
		beta = 1000.0 * np.ones_like(x)
	elif exp == 'C':
		a = 0.1 * np.pi / 180
		S = -x * np.tan(a)
		B = S - 1000.0
		beta = 1000.0 + 1000.0 * np.sin(2 * np.pi * x / L) * np.sin(2 * np.pi * y / L)
	else:
		raise ValueError(exp)
	if periodic:
		# beta lives in the periodic space (model.py:2523-2530): slaves take the master value
		master = periodic_vertex_map(coords)
		beta = beta[master]
	coords = deform(coords, S, B)
	fc, fl = exterior_facets(cells)
	ff = mark_facets(coords, cells, fc, fl)
	cd = canonical_cell_dofs(cells, v2n)
	return BPProblem(coords, cells, cd, 2 * n_nodes, fc, fl, ff, S, 1e-16, beta,
	                 periodic=periodic, quad_degree=quad_degree)
