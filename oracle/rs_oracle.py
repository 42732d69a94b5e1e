"""numpy restatement of ``MomentumDukowiczStokesReduced`` (oracle).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- PARITY UNPINNED.

cslvr/momentumstokes.py:300-507: the Dukowicz "reduced" full-Stokes model keeps the
Blatter-Pattyn unknown (u_x, u_y) in QM2 and the frozen P1 vertical velocity ``self.u_z``
enters the strain-rate tensor (:406-418)

    eps = sym grad (u_x, u_y, u_z)   with   eps_22 := -(u_x,x + u_y,y)

so the effective strain rate (momentum.py:245-251) is the BP one with u_x,z -> u_x,z + u_z,x
and u_y,z -> u_y,z + u_z,y.  The action (:381) is

    A = (V_d - P_e) dOmega - S_l dGamma_bg - P_b dGamma_bw [- P_b dGamma_lt]
    P_e = -rho g (u_h . grad_h S + u_z)                                   (:366)
    S_l = -1/2 beta (u_x^2 + u_y^2 + u_z_b^2),  u_z_b = (-B_ring - u_h . n_h) / n_z   (:369-370)
    P_b = f_w u . n,  f_w = rho g (S - z) - rho_sw g D                    (:373-374)

with n the facet normal; sliding acts on the GROUNDED bed only (dGamma_bg, marker 3) where the
BP classes use the whole bed (dGamma_b).  The terms in u_z alone drop out of the first
variation with respect to u_h (:398).

The solve (:449-486) is ``Model.home_rolled_newton_method`` (cslvr/model.py:2610-2730) with
``solve_vert_velocity`` as the per-iteration callback, atol 1e-6, then ``solve_pressure`` and
only then ``update_model_var`` -- so the pressure's viscosity (``self.eta`` = eta(model.u),
SURVEY Appendix B) sees the horizontal velocity the model held BEFORE the solve and the u_z
of the last callback.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .bp_oracle import BPProblem, GAMMA_B_GND, GAMMA_B_FLT, GAMMA_L_OVR, GAMMA_L_UDR, facet_geometry
from .quadrature import tetrahedron_rule, triangle_rule


class RSProblem(BPProblem):
	"""``BPProblem`` + bed height ``B`` / basal melt ``B_ring`` (vertex arrays) and the frozen
	vertical velocity ``uz`` (vertex array, :meth:`set_uz`; zero like ``Function(model.Q)``)."""

	def __init__(self, *args, **kwargs):
		B = kwargs.pop('B')
		B_ring = kwargs.pop('B_ring', None)
		super(RSProblem, self).__init__(*args, **kwargs)
		nv = self.coords.shape[0]
		self.B = np.asarray(B, dtype=np.float64) * np.ones(nv)
		self.B_ring = np.zeros(nv) if B_ring is None else np.asarray(B_ring, dtype=np.float64) * np.ones(nv)
		self.uz = np.zeros(nv)

	def set_uz(self, uz_vertex):
		self.uz = np.asarray(uz_vertex, dtype=np.float64).copy()

	def _grad_w(self, uz=None):
		uz = self.uz if uz is None else uz
		return np.einsum('ta,taj->tj', uz[self.cells], self.G)

	def _grad_u(self, u, uz=None):
		"""velocity gradient entering the strain-rate tensor of momentumstokes.py:406-418:
		(0,2) and (1,2) carry u_z,x and u_z,y (eps_02 = (u_x,z + u_z,x)/2, ...)."""
		g = super(RSProblem, self)._grad_u(u)
		gw = self._grad_w(uz)
		g[:, 0, 2] += gw[:, 0]
		g[:, 1, 2] += gw[:, 1]
		return g

	# ---- facet integrals, momentumstokes.py:369-386 -----------------------------
	def _facet_sets(self):
		grounded = self.ff == GAMMA_B_GND                              # dGamma_bg, d3model.py:574-575
		press = self.ff == GAMMA_B_FLT                                 # dGamma_bw
		if (not self.periodic) and self.use_pressure_bc:               # momentumstokes.py:383-386
			press = press | np.isin(self.ff, (GAMMA_L_OVR, GAMMA_L_UDR))
		return grounded, press

	def facet_tensors(self, u):
		rho_g = self.c['rho_i'] * self.c['g']
		rsw_g = self.c['rhosw'] * self.c['g']
		nf = self.fc.size
		act = np.zeros(nf)
		F = np.zeros((nf, 8))
		J = np.zeros((nf, 8, 8))
		if nf == 0:
			return act, F, J
		grounded, press = self._facet_sets()
		fv, area, nrm, _ = facet_geometry(self.coords, self.cells, self.fc, self.fl)
		loc = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])[self.fl]
		Uc = u[self.cell_dofs[self.fc]].reshape(-1, 2, 4)
		Uf = np.take_along_axis(Uc, loc[:, None, :], axis=2)           # [nf,2,3]
		bf = self.beta[fv]
		Brf = self.B_ring[fv]
		fw = rho_g * (self.S[fv] - self.coords[fv, 2]) - rsw_g * self.D[fv]
		r = np.arange(nf)
		nz = np.where(grounded, nrm[:, 2], 1.0)                        # only used where grounded
		# -S_l = 1/2 beta (u_x^2 + u_y^2 + u_z_b^2): beta (P1) x (P1)^2 -> degree 3
		pts, wts = triangle_rule(3)
		for lam, w in zip(pts, wts):
			dx = w * area * grounded
			bq = bf @ lam
			uq = Uf @ lam                                              # [nf,2]
			uzb = (-(Brf @ lam) - uq[:, 0] * nrm[:, 0] - uq[:, 1] * nrm[:, 1]) / nz
			act += dx * 0.5 * bq * ((uq ** 2).sum(axis=1) + uzb ** 2)
			for c in range(2):
				duzb_c = -nrm[:, c] / nz                               # d u_z_b / d u_c
				for a in range(3):
					F[r, 4 * c + loc[:, a]] += dx * bq * (uq[:, c] + uzb * duzb_c) * lam[a]
					for d in range(2):
						duzb_d = -nrm[:, d] / nz
						for b in range(3):
							J[r, 4 * c + loc[:, a], 4 * d + loc[:, b]] += dx * bq * ((c == d) + duzb_c * duzb_d) * lam[a] * lam[b]
		# -P_b = -f_w u.n: the u_z n_z part does not depend on u_h
		pts, wts = triangle_rule(2)
		uzf = self.uz[fv]
		for lam, w in zip(pts, wts):
			dx = w * area * press
			fq = fw @ lam
			uq = Uf @ lam
			act -= dx * fq * (uq[:, 0] * nrm[:, 0] + uq[:, 1] * nrm[:, 1] + (uzf @ lam) * nrm[:, 2])
			for c in range(2):
				for a in range(3):
					F[r, 4 * c + loc[:, a]] -= dx * fq * nrm[:, c] * lam[a]
		return act, F, J

	def action(self, u):
		"""the action including the P_e term in u_z alone (momentumstokes.py:366), which has no
		first variation in u_h."""
		a = super(RSProblem, self).action(u)
		rho_g = self.c['rho_i'] * self.c['g']
		return a + rho_g * (self.vol * self.uz[self.cells].mean(axis=1)).sum()

	# ---- linear=True ---------------------------------------------------------------
	def linear_solve(self, u_l, uz_l=None):
		"""Vd = 2 eta(model.u) epsdot(u_h, u_z) (momentum.py:213-216): K = second variation with
		the viscosity frozen at (u_l, uz_l), f = first variation at u_h = 0 (which here keeps the
		u_z,x / u_z,y part of d epsdot)."""
		n, eps0 = self.c['n'], self.c['eps_reg']
		uz_l = self.uz if uz_l is None else uz_l
		pts, wts = tetrahedron_rule(self.quad_degree)
		ed_l = self._epsdot(self._grad_u(u_l, uz_l)) + eps0
		zero = np.zeros(self.n_dofs)
		de0 = self._depsdot(self._grad_u(zero))
		H = self._d2epsdot()
		Av = self.A[self.cells]
		nt = self.cells.shape[0]
		Kc = np.zeros((nt, 8, 8))
		fc_ = np.zeros((nt, 8))
		for lam, w in zip(pts, wts):
			two_eta = (w * self.vol) * self._A_at(lam, Av) ** (-1.0 / n) * ed_l ** ((1 - n) / (2 * n))
			Kc += two_eta[:, None, None] * H
			fc_ += two_eta[:, None] * de0
		# gravity and facets at u_h = 0: residual(0) minus its (non-linear) viscous part
		f = self.residual(zero) - self._viscous_residual(zero)
		np.add.at(f, self.cell_dofs.ravel(), fc_.ravel())
		nd = self.n_dofs
		rows = np.repeat(self.cell_dofs, 8, axis=1).ravel()
		cols = np.tile(self.cell_dofs, (1, 8)).ravel()
		vals = Kc.ravel()
		if self.fc.size:
			_, _, Jf = self.facet_tensors(zero)
			cd = self.cell_dofs[self.fc]
			rows = np.concatenate([rows, np.repeat(cd, 8, axis=1).ravel()])
			cols = np.concatenate([cols, np.tile(cd, (1, 8)).ravel()])
			vals = np.concatenate([vals, Jf.ravel()])
		K = sp.coo_matrix((vals, (rows, cols)), shape=(nd, nd)).tocsc()
		return spla.spsolve(K, -f)

	# ---- solve_pressure as MomentumDukowiczStokesReduced.solve reaches it ------------
	def pressure_rs(self, u, u_model, uz_model, quad_degree=6):
		"""p = project(rho g (S - z) - 2 eta (u_x,x + u_y,y), Q), momentumbp.py:205-229, with
		eta = eta(model.u) under the reduced-Stokes strain-rate tensor (``self.eta`` is built
		with a truthy ``linear`` argument) and the divergence of the unknown."""
		n, eps0 = self.c['n'], self.c['eps_reg']
		rho_g = self.c['rho_i'] * self.c['g']
		pts, wts = tetrahedron_rule(quad_degree)
		ed = self._epsdot(self._grad_u(u_model, uz_model)) + eps0
		g = BPProblem._grad_u(self, u)
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

	# ---- Model.home_rolled_newton_method, cslvr/model.py:2610-2730 -------------------
	def newton_rs(self, u0=None, rtol=1e-9, atol=1e-6, maxit=50, relax=0.8, callback=True, verbose=False):
		"""while not converged and nIter < max_iter:  A, b = assemble_system(J, -R);
		solve(A, d, b);  residual = |b|;  converged = residual < atol or residual / residual_0
		< rtol;  U += lmbda d;  nIter += 1;  cb_ftn()  -- the update is applied in the iteration
		that detects convergence as well, and the callback (``solve_vert_velocity``,
		momentumstokes.py:468) refreshes u_z after every update.

		No cold start: the reference's is commented out (momentumstokes.py:463-466)."""
		u = np.zeros(self.n_dofs) if u0 is None else np.array(u0, dtype=np.float64)
		v2n = np.empty(self.coords.shape[0], dtype=np.int64)
		nd = self.cell_dofs[:, :4] // 2
		v2n[self.cells.ravel()] = nd.ravel()
		conv, it, hist = False, 0, []
		r0 = None
		while not conv and it < maxit:
			b = -self.residual(u)
			d = spla.splu(self.jacobian_matrix(u).tocsc()).solve(b)
			r = np.linalg.norm(b)
			if it == 0:
				r0 = r
			hist.append(r)
			conv = (r < atol) or (r / r0 < rtol)
			u = u + relax * d
			it += 1
			if verbose:
				print("Newton iteration %d: r (abs) = %.3e (tol = %.3e) r (rel) = %.3e (tol = %.3e)" % (it, r, atol, r / r0, rtol))
			if callback:
				uz_node, _ = self.vert_velocity(u, self.B, self.B_ring)
				self.set_uz(uz_node[v2n])
		return u, dict(converged=bool(conv), iterations=it, residuals=hist)
