"""ctypes wrapper of oracle/bp_oracle_c.c (TEST INFRASTRUCTURE; see oracle/__init__.py).

``COracle(problem)`` takes a :class:`oracle.bp_oracle.BPProblem` (for its arrays and
CSR pattern) and assembles / solves with the C twin -- used for parity at sizes the
numpy oracle is too slow for and as the CPU baseline of bench.py."""
import ctypes as C
import os
import time

import numpy as np

from . import build_c
from .quadrature import tetrahedron_rule, triangle_rule


class _P(C.Structure):
	_fields_ = [('n', C.c_double), ('eps_reg', C.c_double), ('rho_i', C.c_double),
	            ('rho_sw', C.c_double), ('g', C.c_double), ('periodic', C.c_int),
	            ('use_pressure_bc', C.c_int)]


def _lib():
	lib = C.CDLL(build_c.build())
	lib.bpo_num_threads.restype = C.c_int
	lib.bpo_pcg_jacobi.restype = C.c_int
	return lib


def _p(a):
	return a.ctypes.data_as(C.c_void_p)


class COracle(object):
	def __init__(self, prob, threads=None):
		"""``threads``: OpenMP threads of the C twin (default: what the OpenMP runtime picks; under
		torch.distributed.run that is 1 unless asked for explicitly)."""
		self.lib = _lib()
		if threads:
			self.lib.bpo_set_num_threads(C.c_int(int(threads)))
		self.prob = prob
		self.coords = np.ascontiguousarray(prob.coords, dtype=np.float64)
		self.cells = np.ascontiguousarray(prob.cells, dtype=np.int32)
		self.cell_dofs = np.ascontiguousarray(prob.cell_dofs, dtype=np.int32)
		self.S = np.ascontiguousarray(prob.S)
		self.A = np.ascontiguousarray(prob.A)
		self.beta = np.ascontiguousarray(prob.beta)
		self.fc = np.ascontiguousarray(prob.fc, dtype=np.int32)
		self.fl = np.ascontiguousarray(prob.fl, dtype=np.int32)
		self.ff = np.ascontiguousarray(prob.ff, dtype=np.int32)
		ip, ix = prob.pattern()
		self.indptr = np.ascontiguousarray(ip, dtype=np.int64)
		self.indices = np.ascontiguousarray(ix, dtype=np.int32)
		self.q = [np.ascontiguousarray(a) for a in tetrahedron_rule(prob.quad_degree)]
		self.q3 = [np.ascontiguousarray(a) for a in triangle_rule(3)]
		self.q2 = [np.ascontiguousarray(a) for a in triangle_rule(2)]
		c = prob.c
		self.P = _P(c['n'], c['eps_reg'], c['rho_i'], c['rhosw'], c['g'], int(prob.periodic), int(prob.use_pressure_bc))
		self.threads = self.lib.bpo_num_threads()

	def assemble(self, u, want_F=True, want_J=True, n_cells=None):
		"""returns (F, Jvals, seconds); ``n_cells`` bounds the cell loop (timing samples)."""
		u = np.ascontiguousarray(u, dtype=np.float64)
		F = np.zeros(self.prob.n_dofs) if want_F else None
		Jv = np.zeros(self.indices.size) if want_J else None
		nt = self.cells.shape[0] if n_cells is None else int(n_cells)
		t0 = time.perf_counter()
		self.lib.bpo_assemble_cells(C.c_int64(nt), _p(self.coords), _p(self.cells), _p(self.cell_dofs),
		                            _p(self.S), _p(self.A), _p(u), C.c_int(len(self.q[1])), _p(self.q[0]), _p(self.q[1]),
		                            C.byref(self.P), _p(self.indptr), _p(self.indices),
		                            _p(F) if want_F else None, _p(Jv) if want_J else None)
		if n_cells is None and self.fc.size:
			self.lib.bpo_assemble_facets(C.c_int64(self.fc.size), _p(self.fc), _p(self.fl), _p(self.ff),
			                             _p(self.coords), _p(self.cells), _p(self.cell_dofs), _p(self.S), _p(self.beta), _p(u),
			                             C.c_int(len(self.q3[1])), _p(self.q3[0]), _p(self.q3[1]),
			                             C.c_int(len(self.q2[1])), _p(self.q2[0]), _p(self.q2[1]),
			                             C.byref(self.P), _p(self.indptr), _p(self.indices),
			                             _p(F) if want_F else None, _p(Jv) if want_J else None)
		return F, Jv, time.perf_counter() - t0

	def pcg(self, Jv, b, rtol=1e-5, atol=1e-50, maxit=10000):
		n = self.prob.n_dofs
		x = np.zeros(n)
		work = np.empty(5 * n)
		it = self.lib.bpo_pcg_jacobi(C.c_int64(n), _p(self.indptr), _p(self.indices), _p(Jv), _p(np.ascontiguousarray(b)),
		                             _p(x), C.c_double(rtol), C.c_double(atol), C.c_int(maxit), _p(work))
		return x, it

	def newton(self, rtol=1e-9, atol=1e-10, maxit=25, relax=1.0, krylov_rtol=1e-5):
		"""dolfin NewtonSolver semantics with Jacobi-PCG (SURVEY A-N, A-K)."""
		u = np.full(self.prob.n_dofs, 3.0e-16)
		b, Jv, _ = self.assemble(u)
		r0 = np.linalg.norm(b)
		hist, it, kits = [r0], 0, 0
		conv = r0 < atol
		while not conv and it < maxit:
			dx, k = self.pcg(Jv, b, rtol=krylov_rtol)
			kits += k
			u -= relax * dx
			b, Jv, _ = self.assemble(u)
			r = np.linalg.norm(b)
			hist.append(r)
			it += 1
			conv = (r / r0 < rtol) or (r < atol)
		return u, dict(converged=bool(conv), iterations=it, residuals=hist, krylov_iterations=kits)
