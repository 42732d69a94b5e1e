"""ctypes binding of ``include/cslvr_b200.h`` -- the only way Python reaches the kernels.

PyTorch is used for nothing but owning device buffers (``torch.empty(...,
device='cuda')`` -> ``data_ptr()``) and the current stream.  There is no CPU
fallback: if the shared library is missing or no GPU is visible the calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import _build

BP_OK, BP_ERR_ARG, BP_ERR_CUDA, BP_ERR_STATE, BP_ERR_COMM, BP_ERR_NONCONV = range(6)
FIELD_S, FIELD_A, FIELD_BETA, FIELD_B, FIELD_B_RING, FIELD_TP, FIELD_W, FIELD_E, FIELD_UZ = range(9)
ASSEMBLE_F, ASSEMBLE_J = 1, 2
PC = {'none': 0, 'jacobi': 1, 'block_jacobi': 2, 'chebyshev': 3, 'column': 4, 'mg': 5}

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)


class bp_mesh(C.Structure):
	_fields_ = [('n_vertices', C.c_int64), ('n_cells', C.c_int64), ('n_dofs', C.c_int64),
	            ('coords', c_f64p), ('cells', c_i32p), ('cell_dofs', c_i32p),
	            ('vertex_to_node', c_i32p),
	            ('n_facets', C.c_int64), ('facet_cell', c_i32p), ('facet_local', c_i32p),
	            ('facet_marker', c_i32p),
	            ('n_owned_nodes', C.c_int64), ('n_neighbors', C.c_int32),
	            ('neighbor_rank', c_i32p), ('send_ptr', c_i64p), ('send_nodes', c_i32p),
	            ('recv_ptr', c_i64p)]


class bp_params(C.Structure):
	_fields_ = [('n', C.c_double), ('eps_reg', C.c_double), ('rho_i', C.c_double),
	            ('rho_sw', C.c_double), ('g', C.c_double),
	            ('periodic', C.c_int32), ('use_pressure_bc', C.c_int32),
	            ('n_quad', C.c_int32), ('quad_points', c_f64p), ('quad_weights', c_f64p),
	            ('newton_rtol', C.c_double), ('newton_atol', C.c_double),
	            ('newton_maxit', C.c_int32), ('newton_relax', C.c_double),
	            ('error_on_nonconvergence', C.c_int32),
	            ('krylov_rtol', C.c_double), ('krylov_atol', C.c_double),
	            ('krylov_maxit', C.c_int32), ('pc', C.c_int32), ('pc_degree', C.c_int32),
	            ('verbose', C.c_int32), ('assembly', C.c_int32), ('pc_eig_ratio', C.c_double),
	            ('n_quad_aux', C.c_int32), ('quad_points_aux', c_f64p), ('quad_weights_aux', c_f64p),
	            ('energy_dependent', C.c_int32), ('a_T_l', C.c_double), ('a_T_u', C.c_double),
	            ('Q_T_l', C.c_double), ('Q_T_u', C.c_double), ('R_gas', C.c_double),
	            ('krylov_norm_preconditioned', C.c_int32), ('mg_smoother', C.c_int32),
	            ('reduced_stokes', C.c_int32), ('mg_coarse_scale', C.c_double),
	            ('krylov_error_on_nonconvergence', C.c_int32)]


class bp_stats(C.Structure):
	_fields_ = [('converged', C.c_int32), ('newton_iterations', C.c_int32),
	            ('krylov_iterations', C.c_int32),
	            ('residual_norm0', C.c_double), ('residual_norm', C.c_double),
	            ('t_assemble_ms', C.c_double), ('t_krylov_ms', C.c_double),
	            ('t_total_ms', C.c_double), ('n_assemblies', C.c_int32),
	            ('gpu_launches', C.c_int32), ('residual_history', C.c_double * 64),
	            ('krylov_unconverged', C.c_int32), ('krylov_last_relres', C.c_double),
	            ('host_launches', C.c_int32)]

	def as_dict(self):
		n = min(self.newton_iterations + 1, 64)
		return dict(converged=bool(self.converged), iterations=int(self.newton_iterations),
		            krylov_iterations=int(self.krylov_iterations),
		            residual_norm0=self.residual_norm0, residual_norm=self.residual_norm,
		            residuals=[self.residual_history[i] for i in range(n)],
		            t_assemble_ms=self.t_assemble_ms, t_krylov_ms=self.t_krylov_ms,
		            t_total_ms=self.t_total_ms, n_assemblies=int(self.n_assemblies),
		            gpu_launches=int(self.gpu_launches),
		            krylov_unconverged=int(self.krylov_unconverged), krylov_last_relres=self.krylov_last_relres,
		            host_launches=int(self.host_launches))


# every symbol include/cslvr_b200.h declares (tests check the library exports them all)
SYMBOLS = ['bp_create', 'bp_destroy', 'bp_last_error', 'bp_set_params', 'bp_set_stream',
           'bp_set_field', 'bp_set_dirichlet', 'bp_set_dirichlet_uz', 'bp_get_sizes', 'bp_csr_pattern', 'bp_assemble', 'bp_spmv',
           'bp_krylov_solve', 'bp_newton_solve', 'bp_time_assemble', 'bp_time_spmv',
           'bp_launch_count', 'bp_comm_unique_id', 'bp_comm_init',
           'bp_newton_solve_group', 'bp_assemble_group', 'bp_solve_pressure', 'bp_solve_vert_velocity', 'bp_linear_solve',
           'bp_rs_newton_solve', 'bp_rs_solve_pressure',
           'bp_debug_create_host', 'bp_debug_bsr_pattern', 'bp_debug_tiles_emulate', 'bp_debug_tiles_emulate_facets', 'bp_debug_tiles_info', 'bp_debug_tiles_why']

_lib = None


def load_library():
	"""dlopen cslvr_b200/lib/libcslvr_b200.so; raises if it has not been built."""
	global _lib
	if _lib is not None:
		return _lib
	path = _build.lib_path()
	if not os.path.exists(path):
		raise RuntimeError("libcslvr_b200.so is missing (%s): run `python -c 'import "
		                   "__graft_entry__ as g; g.build()'`; there is no CPU fallback" % path)
	lib = C.CDLL(path)
	vp = C.c_void_p
	lib.bp_create.argtypes = [C.POINTER(vp), C.POINTER(bp_mesh), C.POINTER(bp_params), C.c_int]
	lib.bp_destroy.argtypes = [vp]
	lib.bp_destroy.restype = None
	lib.bp_last_error.argtypes = [vp]
	lib.bp_last_error.restype = C.c_char_p
	lib.bp_set_params.argtypes = [vp, C.POINTER(bp_params)]
	lib.bp_set_stream.argtypes = [vp, vp]
	lib.bp_set_field.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int]
	lib.bp_get_sizes.argtypes = [vp] + [c_i64p] * 6
	lib.bp_csr_pattern.argtypes = [vp, c_i64p, C.POINTER(c_i64p), C.POINTER(c_i32p)]
	lib.bp_assemble.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int]
	lib.bp_spmv.argtypes = [vp, vp, vp, C.c_int]
	lib.bp_krylov_solve.argtypes = [vp, vp, vp, C.c_int, c_i32p, c_f64p]
	lib.bp_newton_solve.argtypes = [vp, vp, C.c_int, C.POINTER(bp_stats)]
	lib.bp_linear_solve.argtypes = [vp, vp, vp, C.c_int, c_i32p]
	lib.bp_solve_pressure.argtypes = [vp, vp, vp, C.c_int, c_i32p]
	lib.bp_solve_vert_velocity.argtypes = [vp, vp, vp, C.c_int, c_i32p]
	lib.bp_rs_newton_solve.argtypes = [vp, vp, vp, C.c_int, C.c_double, C.c_int, C.POINTER(bp_stats)]
	lib.bp_rs_solve_pressure.argtypes = [vp, vp, vp, vp, C.c_int, c_i32p]
	lib.bp_time_assemble.argtypes = [vp, C.c_int, C.c_int, C.c_int, c_f64p]
	lib.bp_time_spmv.argtypes = [vp, C.c_int, C.c_int, c_f64p]
	lib.bp_launch_count.argtypes = [vp]
	lib.bp_launch_count.restype = C.c_int64
	lib.bp_comm_unique_id.argtypes = [C.c_char_p, C.c_char_p]
	lib.bp_comm_init.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int, C.c_int]
	lib.bp_newton_solve_group.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_int, C.POINTER(bp_stats)]
	lib.bp_assemble_group.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp), C.c_int]
	_lib = lib
	return lib


def nccl_library_path():
	"""the NCCL torch itself is linked against (so both agree on one libnccl)."""
	try:
		import nvidia.nccl
		for d in list(getattr(nvidia.nccl, '__path__', [])):
			p = os.path.join(d, 'lib', 'libnccl.so.2')
			if os.path.exists(p):
				return p
	except ImportError:
		pass
	return 'libnccl.so.2'


def _is_torch(x):
	return hasattr(x, 'data_ptr') and hasattr(x, 'is_cuda')


def _ptr(x, dtype=np.float64, out=False):
	"""(void*, on_device, keepalive) of a numpy array or a torch tensor.

	Stream contract (include/cslvr_b200.h): the library reads and writes device buffers on the
	context's stream.  A torch CUDA tensor produced on another stream is ordered by
	synchronising that torch stream first (BPContext._order); every C-ABI call returns with
	its outputs complete.  ``out=True``: the buffer is written by the library, so an array that
	numpy would have to copy (wrong dtype, non-contiguous) is rejected instead of silently lost."""
	if x is None:
		return None, 0, None
	if _is_torch(x):
		import torch
		if x.dtype != torch.float64:
			raise ValueError("tensors passed to the C-ABI must be float64, got %s" % x.dtype)
		if not x.is_contiguous():
			raise ValueError("device tensors passed to the C-ABI must be contiguous")
		return C.c_void_p(x.data_ptr()), int(x.is_cuda), x
	if out and not (isinstance(x, np.ndarray) and x.dtype == dtype and x.flags.c_contiguous and x.flags.writeable):
		raise ValueError("output buffers must be writeable C-contiguous %s numpy arrays (a copy would lose the result)" % np.dtype(dtype).name)
	a = np.ascontiguousarray(x, dtype=dtype)
	return a.ctypes.data_as(C.c_void_p), 0, a


class BPError(RuntimeError):
	def __init__(self, code, msg):
		RuntimeError.__init__(self, "cslvr_b200 error %d: %s" % (code, msg))
		self.code = code


class BPContext(object):
	"""One ``bp_ctx``: a (partition of a) mesh resident on one GPU."""

	def __init__(self, coords, cells, n_dofs, facets, params, cell_dofs=None,
	             vertex_to_node=None, partition=None, device=0):
		"""``facets`` = (facet_cell, facet_local, facet_marker); ``params`` = dict for
		:func:`make_params`; ``partition`` = dict(n_owned_nodes, neighbor_rank, send_ptr,
		send_nodes, recv_ptr) or None."""
		self.lib = load_library()
		self._keep = []
		m = bp_mesh()
		coords = np.ascontiguousarray(coords, dtype=np.float64)
		cells = np.ascontiguousarray(cells, dtype=np.int32)
		m.n_vertices, m.n_cells, m.n_dofs = coords.shape[0], cells.shape[0], int(n_dofs)
		m.coords = coords.ctypes.data_as(c_f64p)
		m.cells = cells.ctypes.data_as(c_i32p)
		self._keep += [coords, cells]

		def i32(a):
			if a is None:
				return None
			a = np.ascontiguousarray(a, dtype=np.int32)
			self._keep.append(a)
			return a.ctypes.data_as(c_i32p)

		def i64(a):
			a = np.ascontiguousarray(a, dtype=np.int64)
			self._keep.append(a)
			return a.ctypes.data_as(c_i64p)
		m.cell_dofs = i32(cell_dofs)
		m.vertex_to_node = i32(vertex_to_node)
		fc, fl, fm = facets
		m.n_facets = len(fc)
		m.facet_cell, m.facet_local, m.facet_marker = i32(fc), i32(fl), i32(fm)
		if partition is None:
			m.n_owned_nodes, m.n_neighbors = -1, 0
		else:
			m.n_owned_nodes = int(partition['n_owned_nodes'])
			m.n_neighbors = len(partition['neighbor_rank'])
			m.neighbor_rank = i32(partition['neighbor_rank'])
			m.send_ptr = i64(partition['send_ptr'])
			m.send_nodes = i32(partition['send_nodes'])
			m.recv_ptr = i64(partition['recv_ptr'])
		self.params = dict(params)
		p = self._make_params(self.params)
		h = C.c_void_p()
		rc = self.lib.bp_create(C.byref(h), C.byref(m), C.byref(p), int(device))
		if rc != BP_OK:
			raise BPError(rc, self.lib.bp_last_error(None).decode())
		self.h = h
		self.device = int(device)
		self.n_vertices = int(m.n_vertices)
		self.n_dofs = int(n_dofs)
		sz = [C.c_int64() for _ in range(6)]
		self._check(self.lib.bp_get_sizes(self.h, *[C.byref(s) for s in sz]))
		self.n_nodes, _, self.nnz, self.nnz_blocks, self.n_cells, self.n_basal = [s.value for s in sz]

	# -- plumbing ---------------------------------------------------------
	def _make_params(self, d):
		from .quadrature import tetrahedron_rule
		p = bp_params()
		p.n = d.get('n', 3.0)
		p.eps_reg = d.get('eps_reg', 1e-15)
		p.rho_i = d.get('rho_i', 910.0)
		p.rho_sw = d.get('rhosw', 1028.0)
		p.g = d.get('g', 9.80665)
		p.periodic = int(bool(d.get('periodic', False)))
		p.use_pressure_bc = int(bool(d.get('use_pressure_bc', True)))
		p.energy_dependent = int(bool(d.get('energy_dependent_rate_factor', False)))
		spy = 31556926.0
		p.a_T_l, p.a_T_u = d.get('a_T_l', 3.985e-13 * spy), d.get('a_T_u', 1.916e3 * spy)
		p.Q_T_l, p.Q_T_u, p.R_gas = d.get('Q_T_l', 6e4), d.get('Q_T_u', 13.9e4), d.get('R', 8.3144621)
		# UFL's degree estimate: 5 with a P1 rate factor, 9 with the energy-dependent expression (SURVEY A-Q)
		pts, wts = tetrahedron_rule(int(d.get('quadrature_degree', 9 if p.energy_dependent else 5)))
		pts = np.ascontiguousarray(pts, dtype=np.float64)
		wts = np.ascontiguousarray(wts, dtype=np.float64)
		self._quad = (pts, wts)
		p.n_quad = len(wts)
		p.quad_points = pts.ctypes.data_as(c_f64p)
		p.quad_weights = wts.ctypes.data_as(c_f64p)
		pa, wa = tetrahedron_rule(int(d.get('projection_quadrature_degree', 6)))
		pa = np.ascontiguousarray(pa, dtype=np.float64)
		wa = np.ascontiguousarray(wa, dtype=np.float64)
		self._quad_aux = (pa, wa)
		p.n_quad_aux = len(wa)
		p.quad_points_aux = pa.ctypes.data_as(c_f64p)
		p.quad_weights_aux = wa.ctypes.data_as(c_f64p)
		p.newton_rtol = d.get('relative_tolerance', 1e-9)
		p.newton_atol = d.get('absolute_tolerance', 1e-10)
		p.newton_maxit = int(d.get('maximum_iterations', 25))
		p.newton_relax = d.get('relaxation_parameter', 1.0)
		p.error_on_nonconvergence = int(bool(d.get('error_on_nonconvergence', False)))
		p.krylov_rtol = d.get('krylov_rtol', 1e-5)
		p.krylov_atol = d.get('krylov_atol', 1e-50)
		p.krylov_maxit = int(d.get('krylov_maxit', 10000))
		pc = d.get('preconditioner', 'mg')
		p.pc = PC[pc] if isinstance(pc, str) else int(pc)
		is_mg = (pc == 'mg' or pc == 5)          # Chebyshev as a smoother wants a short interval
		p.pc_degree = int(d.get('pc_degree', 3 if is_mg else 8))
		p.pc_eig_ratio = float(d.get('pc_eig_ratio', 6.0 if is_mg else 100.0))
		p.verbose = int(bool(d.get('verbose', False)))
		p.krylov_norm_preconditioned = int(d.get('krylov_norm', 'unpreconditioned') == 'preconditioned')
		p.mg_smoother = {'auto': -1, 'block': 0, 'column': 1}[d.get('mg_smoother', 'auto')]
		p.assembly = {'auto': 0, 'scatter': 1, 'gather': 2, 'tiled': 3}[d.get('assembly', 'auto')]
		p.reduced_stokes = int(bool(d.get('reduced_stokes', False)))
		p.mg_coarse_scale = float(d.get('mg_coarse_scale', 0.0))
		p.krylov_error_on_nonconvergence = int(bool(d.get('krylov_error_on_nonconvergence', True)))
		self._last_params = p
		return p

	def pc_label(self):
		"""human-readable preconditioner of the last parameter set (for logs / bench lines)."""
		names = {v: k for k, v in PC.items()}
		p = self._last_params
		nm = names.get(p.pc, str(p.pc))
		return nm if p.pc <= 2 or p.pc == 4 else "%s(degree %d, ratio %g)" % (nm, p.pc_degree, p.pc_eig_ratio)

	def _check(self, rc, allow=()):
		if rc != BP_OK and rc not in allow:
			raise BPError(rc, self.lib.bp_last_error(self.h).decode())
		return rc

	def set_params(self, **kw):
		self.params.update(kw)
		p = self._make_params(self.params)
		self._check(self.lib.bp_set_params(self.h, C.byref(p)))

	def use_torch_stream(self):
		"""run the library on torch's current stream of this device (no cross-stream ordering needed any more,
		torch.cuda.Event sees the kernels, and the stream can be captured into CUDA graphs)."""
		import torch
		h = torch.cuda.current_stream(self.device).cuda_stream
		self._check(self.lib.bp_set_stream(self.h, C.c_void_p(h)))
		self._stream_handle = h

	def _order(self, x):
		"""the library works on its own stream unless use_torch_stream() adopted torch's: a CUDA tensor
		handed in may still be in flight on torch's current stream, so that stream is drained first."""
		if _is_torch(x) and x.is_cuda:
			import torch
			s = torch.cuda.current_stream(x.device)
			if s.cuda_stream != getattr(self, '_stream_handle', None):
				s.synchronize()

	def close(self):
		if getattr(self, 'h', None):
			self.lib.bp_destroy(self.h)
			self.h = None

	def __del__(self):
		try:
			self.close()
		except Exception:
			pass

	# -- the C-ABI calls ----------------------------------------------------
	def set_field(self, field, values):
		if np.isscalar(values):
			values = np.full(self.n_vertices, float(values))
		p, dev, keep = _ptr(values)
		self._order(keep)
		n = keep.numel() if _is_torch(keep) else keep.size
		self._check(self.lib.bp_set_field(self.h, int(field), p, n, dev))

	def csr_pattern(self):
		nnz = C.c_int64()
		ip, ix = c_i64p(), c_i32p()
		self._check(self.lib.bp_csr_pattern(self.h, C.byref(nnz), C.byref(ip), C.byref(ix)))
		indptr = np.ctypeslib.as_array(ip, shape=(self.n_dofs + 1,)).copy()
		indices = np.ctypeslib.as_array(ix, shape=(max(nnz.value, 1),))[:nnz.value].copy()
		return indptr, indices

	def assemble(self, u, want_F=True, want_J=False, F=None, J=None):
		"""assemble(F) and/or assemble(J) at u; numpy in -> numpy out, torch cuda in -> torch out."""
		what = (ASSEMBLE_F if want_F else 0) | (ASSEMBLE_J if want_J else 0)
		pu, dev, ku = _ptr(u)
		self._order(ku)
		if dev:
			import torch
			if want_F and F is None:
				F = torch.empty(self.n_dofs, dtype=torch.float64, device=ku.device)
			if want_J and J is None and J is not False:
				J = torch.empty(self.nnz_csr(), dtype=torch.float64, device=ku.device)
		else:
			if want_F and F is None:
				F = np.empty(self.n_dofs)
			if want_J and J is None and J is not False:
				J = np.empty(self.nnz_csr())
		if J is False:
			J = None
		pF, _, kF = _ptr(F, out=True)
		pJ, _, kJ = _ptr(J, out=True)
		self._check(self.lib.bp_assemble(self.h, what, pu, pF, pJ, dev))
		return F, J

	def nnz_csr(self):
		if not hasattr(self, '_nnz_csr'):
			nnz = C.c_int64()
			self._check(self.lib.bp_csr_pattern(self.h, C.byref(nnz), None, None))
			self._nnz_csr = nnz.value
		return self._nnz_csr

	def spmv(self, x, y=None):
		px, dev, kx = _ptr(x)
		self._order(kx)
		if y is None:
			if dev:
				import torch
				y = torch.empty_like(kx)
			else:
				y = np.empty(self.n_dofs)
		py, _, ky = _ptr(y, out=True)
		self._check(self.lib.bp_spmv(self.h, px, py, dev))
		return y

	def krylov_solve(self, b, x=None):
		pb, dev, kb = _ptr(b)
		self._order(kb)
		if x is None:
			if dev:
				import torch
				x = torch.empty_like(kb)
			else:
				x = np.empty(self.n_dofs)
		px, _, kx = _ptr(x, out=True)
		it, rr = C.c_int32(), C.c_double()
		self._check(self.lib.bp_krylov_solve(self.h, pb, px, dev, C.byref(it), C.byref(rr)))
		return x, it.value, rr.value

	def newton_solve(self, u):
		"""in-place Newton on ``u`` (numpy float64 array or torch cuda tensor); returns stats."""
		if not _is_torch(u) and not (isinstance(u, np.ndarray) and u.dtype == np.float64 and u.flags.c_contiguous):
			raise ValueError("u must be a contiguous float64 numpy array or a torch tensor (updated in place)")
		pu, dev, ku = _ptr(u)
		self._order(ku)
		st = bp_stats()
		rc = self._check(self.lib.bp_newton_solve(self.h, pu, dev, C.byref(st)), allow=(BP_ERR_NONCONV,))
		d = st.as_dict()
		if rc == BP_ERR_NONCONV:
			raise BPError(rc, self.lib.bp_last_error(self.h).decode())
		return d

	def linear_solve(self, u_lin, u_out=None):
		"""one linear solve with the viscosity frozen at eta(u_lin) (``linear=True``)."""
		pl, dev, kl = _ptr(u_lin)
		self._order(kl)
		if u_out is None:
			if dev:
				import torch
				u_out = torch.empty_like(kl)
			else:
				u_out = np.empty(self.n_dofs)
		po, _, ko = _ptr(u_out, out=True)
		it = C.c_int32()
		self._check(self.lib.bp_linear_solve(self.h, pl, po, dev, C.byref(it)))
		self.last_linear_iterations = it.value
		return u_out

	def solve_pressure(self, u, p=None):
		"""BP pressure (one value per mesh vertex) from the velocity dofs ``u``."""
		pu, dev, ku = _ptr(u)
		self._order(ku)
		if p is None:
			if dev:
				import torch
				p = torch.empty(self.n_vertices, dtype=torch.float64, device=ku.device)
			else:
				p = np.empty(self.n_vertices)
		pp, _, kp = _ptr(p, out=True)
		it = C.c_int32()
		self._check(self.lib.bp_solve_pressure(self.h, pu, pp, dev, C.byref(it)))
		self.last_projection_iterations = it.value
		return p

	def solve_vert_velocity(self, u, uz=None):
		"""u_z (one value per mesh vertex) from continuity; needs FIELD_B."""
		pu, dev, ku = _ptr(u)
		self._order(ku)
		if uz is None:
			if dev:
				import torch
				uz = torch.empty(self.n_vertices, dtype=torch.float64, device=ku.device)
			else:
				uz = np.empty(self.n_vertices)
		pz, _, kz = _ptr(uz, out=True)
		it = C.c_int32()
		self._check(self.lib.bp_solve_vert_velocity(self.h, pu, pz, dev, C.byref(it)))
		self.last_vert_iterations = it.value
		return uz

	def rs_newton_solve(self, u, uz=None, atol=1e-6, callback=True):
		"""``Model.home_rolled_newton_method`` with ``solve_vert_velocity`` as the callback
		(MomentumDukowiczStokesReduced.solve): in place on ``u`` and on ``uz`` (per-vertex
		u_z: start value in, last callback's value out; None keeps the context's field)."""
		if not _is_torch(u) and not (isinstance(u, np.ndarray) and u.dtype == np.float64 and u.flags.c_contiguous):
			raise ValueError("u must be a contiguous float64 numpy array or a torch tensor (updated in place)")
		if uz is not None and not _is_torch(uz) and not (isinstance(uz, np.ndarray) and uz.dtype == np.float64 and uz.flags.c_contiguous):
			raise ValueError("uz must be a contiguous float64 numpy array or a torch tensor (updated in place)")
		pu, dev, ku = _ptr(u)
		self._order(ku)
		pz, _, kz = _ptr(uz, out=True)
		st = bp_stats()
		rc = self._check(self.lib.bp_rs_newton_solve(self.h, pu, pz, dev, float(atol), int(bool(callback)), C.byref(st)),
		                 allow=(BP_ERR_NONCONV,))
		d = st.as_dict()
		d['residuals'] = d['residuals'][:d['iterations']]
		if rc == BP_ERR_NONCONV:
			raise BPError(rc, self.lib.bp_last_error(self.h).decode())
		return d

	def rs_solve_pressure(self, u, u_model, p=None):
		"""pressure with the viscosity of ``u_model`` (+ the context's u_z) and the divergence of ``u``."""
		pu, dev, ku = _ptr(u)
		self._order(ku)
		pm, _, km = _ptr(u_model)
		if p is None:
			if dev:
				import torch
				p = torch.empty(self.n_vertices, dtype=torch.float64, device=ku.device)
			else:
				p = np.empty(self.n_vertices)
		pp, _, kp = _ptr(p, out=True)
		it = C.c_int32()
		self._check(self.lib.bp_rs_solve_pressure(self.h, pu, pm, pp, dev, C.byref(it)))
		self.last_projection_iterations = it.value
		return p

	def set_dirichlet(self, dofs, values):
		"""Dirichlet rows of the Newton / linear solve (bp_set_dirichlet): dofs in this context's numbering."""
		d = np.ascontiguousarray(dofs, dtype=np.int32)
		g = np.ascontiguousarray(np.broadcast_to(np.asarray(values, dtype=np.float64), d.shape))
		self.lib.bp_set_dirichlet.argtypes = [C.c_void_p, C.c_int64, c_i32p, c_f64p]
		self._check(self.lib.bp_set_dirichlet(self.h, d.size, d.ctypes.data_as(c_i32p), g.ctypes.data_as(c_f64p)))

	def set_dirichlet_uz(self, vertices, values):
		"""lateral Dirichlet values of the u_z solve (bp_set_dirichlet_uz), by mesh vertex."""
		v = np.ascontiguousarray(vertices, dtype=np.int32)
		g = np.ascontiguousarray(np.broadcast_to(np.asarray(values, dtype=np.float64), v.shape))
		self.lib.bp_set_dirichlet_uz.argtypes = [C.c_void_p, C.c_int64, c_i32p, c_f64p]
		self._check(self.lib.bp_set_dirichlet_uz(self.h, v.size, v.ctypes.data_as(c_i32p), g.ctypes.data_as(c_f64p)))

	def time_assemble(self, what=ASSEMBLE_F | ASSEMBLE_J, repeats=20, flush_l2=True):
		ms = C.c_double()
		self._check(self.lib.bp_time_assemble(self.h, what, repeats, int(flush_l2), C.byref(ms)))
		return ms.value

	def time_spmv(self, repeats=20, flush_l2=True):
		ms = C.c_double()
		self._check(self.lib.bp_time_spmv(self.h, repeats, int(flush_l2), C.byref(ms)))
		return ms.value

	def launch_count(self):
		return int(self.lib.bp_launch_count(self.h))

	def assembly_kernel_name(self):
		"""the cell kernel a fused F + J pass runs with the current parameters (bench lines / profiles)."""
		mode = int(self._last_params.assembly) if self._last_params is not None else 0
		if mode == 1:
			return "k_cell<F,J>"
		self.lib.bp_debug_tiles_why.argtypes = [C.c_void_p]
		self.lib.bp_debug_tiles_why.restype = C.c_char_p
		tiled = mode in (0, 3) and self.lib.bp_debug_tiles_why(self.h) == b"" and os.environ.get('BP_ASM_TILED', '') != '0' \
			and not int(self._last_params.reduced_stokes if self._last_params is not None else 0)
		return "k_tile" if tiled else "k_rows<F,J>"

	def comm_init(self, unique_id, rank, n_ranks):
		self._check(self.lib.bp_comm_init(self.h, nccl_library_path().encode(), unique_id, rank, n_ranks))


def context_from_model(model, use_pressure_bc=True, cell_dofs=None, device=0, **params):
	"""a :class:`BPContext` holding the mesh, markers and coefficient Functions (S, A, beta) of a
	:class:`cslvr_b200.D3Model` -- what ``MomentumBPBase`` builds behind ``solve()``, for callers that
	drive the C-ABI directly (bench.py, the parity tests).  ``params`` go to ``bp_params``."""
	p = dict(n=model.n, eps_reg=model.eps_reg, rho_i=model.rho_i, rhosw=model.rhosw, g=model.g,
	         periodic=model.use_periodic, use_pressure_bc=use_pressure_bc)
	p.update(params)
	c = BPContext(model.mesh.coordinates(), model.mesh.cells(), 2 * model.n_nodes,
	              (model.facet_cell, model.facet_local, model.ff), p,
	              cell_dofs=cell_dofs, device=device,
	              vertex_to_node=(model.vertex_to_node if (model.use_periodic and cell_dofs is None) else None))
	c.set_field(FIELD_S, model.vertex_field(model.S))
	c.set_field(FIELD_A, model.vertex_field(model.A))
	c.set_field(FIELD_BETA, model.vertex_field(model.beta))
	return c


class HostTileContext(object):
	"""DEBUG / CPU tests: a ``bp_ctx`` without a device (``bp_debug_create_host``).  It only holds the index
	tables of the tiled cell kernel, which :meth:`emulate` walks on the host exactly as the kernel does; no
	product path goes through here (include/cslvr_b200.h, "debug hooks")."""

	def __init__(self, coords, cells, n_dofs, facets, params, vertex_to_node=None, partition=None):
		self.lib = load_library()
		self.lib.bp_debug_create_host.argtypes = [C.POINTER(C.c_void_p), C.POINTER(bp_mesh), C.POINTER(bp_params)]
		self.lib.bp_debug_bsr_pattern.argtypes = [C.c_void_p, c_i64p, C.POINTER(c_i32p), C.POINTER(c_i32p)]
		self.lib.bp_debug_tiles_emulate.argtypes = [C.c_void_p, c_f64p, c_f64p, c_f64p, C.c_int, c_f64p, c_f64p]
		self.lib.bp_debug_tiles_info.argtypes = [C.c_void_p] + [c_i64p] * 5
		self.lib.bp_debug_tiles_why.argtypes = [C.c_void_p]
		self.lib.bp_debug_tiles_why.restype = C.c_char_p
		self._keep = []
		m = bp_mesh()
		coords = np.ascontiguousarray(coords, dtype=np.float64)
		cells = np.ascontiguousarray(cells, dtype=np.int32)
		m.n_vertices, m.n_cells, m.n_dofs = coords.shape[0], cells.shape[0], int(n_dofs)
		m.coords, m.cells = coords.ctypes.data_as(c_f64p), cells.ctypes.data_as(c_i32p)
		self._keep += [coords, cells]

		def i32(a):
			a = np.ascontiguousarray(a, dtype=np.int32)
			self._keep.append(a)
			return a.ctypes.data_as(c_i32p)

		def i64(a):
			a = np.ascontiguousarray(a, dtype=np.int64)
			self._keep.append(a)
			return a.ctypes.data_as(c_i64p)
		if vertex_to_node is not None:
			m.vertex_to_node = i32(vertex_to_node)
		fc, fl, fm = facets
		m.n_facets = len(fc)
		m.facet_cell, m.facet_local, m.facet_marker = i32(fc), i32(fl), i32(fm)
		if partition is None:
			m.n_owned_nodes, m.n_neighbors = -1, 0
		else:
			m.n_owned_nodes, m.n_neighbors = int(partition['n_owned_nodes']), len(partition['neighbor_rank'])
			m.neighbor_rank, m.send_ptr = i32(partition['neighbor_rank']), i64(partition['send_ptr'])
			m.send_nodes, m.recv_ptr = i32(partition['send_nodes']), i64(partition['recv_ptr'])
		self._bpc = BPContext.__new__(BPContext)
		p = BPContext._make_params(self._bpc, dict(params))
		h = C.c_void_p()
		rc = self.lib.bp_debug_create_host(C.byref(h), C.byref(m), C.byref(p))
		if rc != BP_OK:
			raise BPError(rc, self.lib.bp_last_error(None).decode())
		self.h = h
		self.coords, self.n_dofs = coords, int(n_dofs)

	def why_not(self):
		return self.lib.bp_debug_tiles_why(self.h).decode()

	def info(self):
		v = [C.c_int64() for _ in range(5)]
		if self.lib.bp_debug_tiles_info(self.h, *[C.byref(x) for x in v]) != BP_OK:
			return None
		return dict(zip(('n_tiles', 'n_prism_columns', 'n_threads_used', 'n_entries', 'n_contributions'), [x.value for x in v]))

	def bsr_pattern(self):
		n = C.c_int64()
		rp, ci = c_i32p(), c_i32p()
		self.lib.bp_debug_bsr_pattern(self.h, C.byref(n), C.byref(rp), C.byref(ci))
		rowptr = np.ctypeslib.as_array(rp, shape=(n.value + 1,)).copy()
		col = np.ctypeslib.as_array(ci, shape=(max(int(rowptr[-1]), 1),))[:rowptr[-1]].copy()
		return rowptr, col

	def emulate(self, S_vertex, u_nodes, ainv_cells, picard=False, beta_vertex=None):
		"""(F[2*n_owned], J blocks [nnzb, 2, 2] in BSR row order) of the CELL integrals at ``u_nodes`` ([2*node + c]);
		with ``beta_vertex`` also the basal facet integrals the kernel folds into its first layer."""
		xyzS = np.ascontiguousarray(np.column_stack([self.coords, np.asarray(S_vertex, dtype=np.float64)]))
		u = np.ascontiguousarray(u_nodes, dtype=np.float64)
		a = np.ascontiguousarray(ainv_cells, dtype=np.float64)
		rowptr, col = self.bsr_pattern()
		F = np.zeros(2 * (rowptr.size - 1))
		J = np.zeros(4 * int(rowptr[-1]))
		if beta_vertex is not None:
			b = np.ascontiguousarray(beta_vertex, dtype=np.float64)
			self.lib.bp_debug_tiles_emulate_facets.argtypes = [C.c_void_p, c_f64p, c_f64p, c_f64p, c_f64p, C.c_int, c_f64p, c_f64p]
			rc = self.lib.bp_debug_tiles_emulate_facets(self.h, xyzS.ctypes.data_as(c_f64p), u.ctypes.data_as(c_f64p), a.ctypes.data_as(c_f64p),
			                                            b.ctypes.data_as(c_f64p), int(bool(picard)), F.ctypes.data_as(c_f64p), J.ctypes.data_as(c_f64p))
		else:
			rc = self.lib.bp_debug_tiles_emulate(self.h, xyzS.ctypes.data_as(c_f64p), u.ctypes.data_as(c_f64p), a.ctypes.data_as(c_f64p),
			                                     int(bool(picard)), F.ctypes.data_as(c_f64p), J.ctypes.data_as(c_f64p))
		if rc != BP_OK:
			raise BPError(rc, self.lib.bp_last_error(self.h).decode())
		return F, J.reshape(-1, 2, 2), rowptr, col

	def close(self):
		if getattr(self, 'h', None):
			self.lib.bp_destroy(self.h)
			self.h = None

	def __del__(self):
		try:
			self.close()
		except Exception:
			pass


def comm_unique_id():
	lib = load_library()
	buf = C.create_string_buffer(128)
	rc = lib.bp_comm_unique_id(nccl_library_path().encode(), buf)
	if rc != BP_OK:
		raise BPError(rc, lib.bp_last_error(None).decode())
	return buf.raw


def newton_solve_group(ctxs, us):
	"""lock-step Newton over several contexts in this process (partitions on one GPU)."""
	lib = load_library()
	n = len(ctxs)
	hs = (C.c_void_p * n)(*[c.h for c in ctxs])
	ptrs, keep = [], []
	dev = None
	for u in us:
		p, d, k = _ptr(u)
		ptrs.append(p)
		keep.append(k)
		dev = d if dev is None else dev
	up = (C.c_void_p * n)(*ptrs)
	st = bp_stats()
	rc = lib.bp_newton_solve_group(hs, n, up, dev, C.byref(st))
	if rc != BP_OK:
		raise BPError(rc, lib.bp_last_error(ctxs[0].h).decode())
	return st.as_dict()


def assemble_group(ctxs, us, want_J=False):
	lib = load_library()
	n = len(ctxs)
	hs = (C.c_void_p * n)(*[c.h for c in ctxs])
	keep = [np.ascontiguousarray(u, dtype=np.float64) for u in us]
	Fs = [np.empty(c.n_dofs) for c in ctxs]
	up = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
	fp = (C.c_void_p * n)(*[f.ctypes.data for f in Fs])
	rc = lib.bp_assemble_group(hs, n, ASSEMBLE_F | (ASSEMBLE_J if want_J else 0), up, fp, 0)
	if rc != BP_OK:
		raise BPError(rc, lib.bp_last_error(ctxs[0].h).decode())
	return Fs
