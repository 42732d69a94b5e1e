"""Problem cases shared by the CPU and GPU tests: each builds the SAME mesh through
the product's host mirror (cslvr_b200.D3Model) and hands its arrays to the oracle."""
import numpy as np

from cslvr_b200 import BoxMesh, Point, D3Model
from oracle import bp_oracle as bo


def ismip_model(exp='A', L=80000.0, n=(15, 15, 5), periodic=True):
	"""docs/get_started.rst:13-45 (A) / simulations/ismip_hom/ismip_hom_c.py:4-25 (C)."""
	mesh = BoxMesh(Point(0, 0, 0), Point(L, L, 1), *n)
	model = D3Model(mesh, use_periodic=periodic)
	model.calculate_boundaries()
	if exp == 'A':
		a = 0.5 * np.pi / 180
		S = lambda x, y, z: -x * np.tan(a)
		B = lambda x, y, z: -x * np.tan(a) - 1000.0 + 500.0 * np.sin(2 * np.pi * x / L) * np.sin(2 * np.pi * y / L)
		beta = 1000.0
	else:
		a = 0.1 * np.pi / 180
		S = lambda x, y, z: -x * np.tan(a)
		B = lambda x, y, z: -x * np.tan(a) - 1000.0
		beta = lambda x, y, z: 1000.0 + 1000.0 * np.sin(2 * np.pi * x / L) * np.sin(2 * np.pi * y / L)
	model.deform_mesh_to_geometry(S, B)
	model.init_beta(beta)
	model.init_A(1e-16)
	return model


def marine_model(n=(8, 7, 4), L=40000.0, variable_A=True):
	"""non-periodic slab with a floating tongue below sea level: exercises
	dGamma_bw (marker 5), dGamma_lt (4, 10), D > 0 and a P1 rate factor."""
	mesh = BoxMesh(Point(0, 0, 0), Point(L, 0.8 * L, 1), *n)
	model = D3Model(mesh, use_periodic=False)
	S = lambda x, y, z: 600.0 - 500.0 * x / L + 30.0 * np.sin(3 * y / L)
	B = lambda x, y, z: -200.0 - 300.0 * x / L + 80.0 * np.cos(5 * x / L) * np.sin(4 * y / L)
	model.calculate_boundaries()
	model.deform_mesh_to_geometry(S, B)
	mask = lambda x, y, z: np.where(x > 0.6 * L, 2.0, 1.0)
	model.calculate_boundaries(mask=mask)
	model.mask_callable = mask                  # for the marker tests: oracle.mark_facets(mask) vs model.ff
	model.init_beta(lambda x, y, z: 200.0 + 150.0 * np.sin(7 * x / L) * np.cos(3 * y / L))
	if variable_A:
		Tp = lambda x, y, z: 250.0 + 18.0 * (1.0 - (z + 500.0) / 1100.0) + 3.0 * np.sin(5 * x / L)
		X = mesh.coordinates()
		model.init_A(model.rate_factor(Tp(X[:, 0], X[:, 1], X[:, 2])))
	else:
		model.init_A(3e-17)
	return model


def rs_problem(model, use_pressure_bc=True, quad_degree=None, B_ring=None, energy=None):
	"""the oracle's MomentumDukowiczStokesReduced problem on the model's arrays."""
	from oracle.rs_oracle import RSProblem
	return RSProblem(model.mesh.coordinates(), model.mesh.cells(), model.cell_dofs, 2 * model.n_nodes,
	                 model.facet_cell, model.facet_local, model.ff,
	                 model.vertex_field(model.S), model.vertex_field(model.A), model.vertex_field(model.beta),
	                 periodic=model.use_periodic, use_pressure_bc=use_pressure_bc, quad_degree=quad_degree, energy=energy,
	                 consts=dict(n=model.n, eps_reg=model.eps_reg, rho_i=model.rho_i, rhosw=model.rhosw, g=model.g),
	                 B=model.vertex_field(model.B), B_ring=B_ring)


def oracle_problem(model, use_pressure_bc=True, quad_degree=None, cell_dofs=None, n_dofs=None, energy=None):
	return bo.BPProblem(model.mesh.coordinates(), model.mesh.cells(),
	                    model.cell_dofs if cell_dofs is None else cell_dofs,
	                    2 * model.n_nodes if n_dofs is None else n_dofs,
	                    model.facet_cell, model.facet_local, model.ff,
	                    model.vertex_field(model.S), model.vertex_field(model.A),
	                    model.vertex_field(model.beta), periodic=model.use_periodic,
	                    use_pressure_bc=use_pressure_bc, quad_degree=quad_degree, energy=energy,
	                    consts=dict(n=model.n, eps_reg=model.eps_reg, rho_i=model.rho_i,
	                                rhosw=model.rhosw, g=model.g))


def product_context(model, **kw):
	"""the product's device context for a model (lives in the package: cslvr_b200.capi.context_from_model)."""
	from cslvr_b200 import capi
	return capi.context_from_model(model, **kw)


def state(n_dofs, seed=1234, scale=30.0):
	return np.random.default_rng(seed).normal(size=n_dofs) * scale


def relerr(a, b):
	return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def relerr_entrywise(a, b, floor=1e-10):
	"""largest |a - b| / |b| over the entries with |b| > floor * max|b|: relerr() is a max-norm measure and
	never looks at a small entry by itself; this one does."""
	a, b = np.asarray(a), np.asarray(b)
	m = np.abs(b) > floor * np.abs(b).max()
	return float((np.abs(a - b)[m] / np.abs(b)[m]).max()) if m.any() else 0.0


def delaunay_extruded_model(n_pts=60, n_layers=4, seed=3, L=30000.0):
	"""an UNSTRUCTURED extruded mesh, the other tet layout D3Model meshes come in
	(cslvr/meshing.py:771-860: gmsh footprint triangles extruded to prisms, each prism split
	into 3 tets by comparing vertex ids): random Delaunay footprint x layers, non-periodic,
	floating part, variable A.  No BoxMesh structure anywhere: vertex ids are shuffled."""
	from scipy.spatial import Delaunay
	from cslvr_b200 import Mesh
	rng = np.random.default_rng(seed)
	pts = rng.uniform(0.05, 0.95, size=(n_pts, 2)) * L
	pts = np.vstack([pts, [[0, 0], [L, 0], [0, L], [L, L]], np.column_stack([np.linspace(0, L, 7)[1:-1], np.zeros(5)]),
	                 np.column_stack([np.linspace(0, L, 7)[1:-1], L * np.ones(5)])])
	tri = Delaunay(pts).simplices
	nf = pts.shape[0]
	perm = rng.permutation(nf * (n_layers + 1))                 # shuffled vertex numbering
	vid = lambda p, k: perm[k * nf + p]
	coords = np.zeros((nf * (n_layers + 1), 3))
	for k in range(n_layers + 1):
		coords[vid(np.arange(nf), k), 0:2] = pts
		coords[vid(np.arange(nf), k), 2] = k / n_layers
	cells = []
	for t in tri:
		for k in range(n_layers):
			b = [vid(p, k) for p in t]
			tp = [vid(p, k + 1) for p in t]
			# split the prism (b0 b1 b2 | t0 t1 t2) into 3 tets, diagonals chosen by the smaller
			# footprint vertex id so that neighbouring prisms agree on their shared quad faces
			o = np.argsort(t)
			b = [b[i] for i in o]; tp = [tp[i] for i in o]
			cells += [[b[0], b[1], b[2], tp[2]], [b[0], b[1], tp[2], tp[1]], [b[0], tp[1], tp[2], tp[0]]]
	cells = np.sort(np.array(cells), axis=1)
	model = D3Model(Mesh(coords, cells), use_periodic=False)
	S = lambda x, y, z: 500.0 - 450.0 * x / L + 40.0 * np.sin(4 * y / L)
	B = lambda x, y, z: -150.0 - 250.0 * x / L + 60.0 * np.cos(6 * x / L) * np.sin(5 * y / L)
	model.calculate_boundaries()
	model.deform_mesh_to_geometry(S, B)
	model.mask_callable = lambda x, y, z: np.where(x > 0.65 * L, 2.0, 1.0)
	model.calculate_boundaries(mask=model.mask_callable)
	model.init_beta(lambda x, y, z: 300.0 + 200.0 * np.sin(6 * x / L) * np.cos(4 * y / L))
	X = model.mesh.coordinates()
	model.init_A(model.rate_factor(252.0 + 15.0 * (1.0 - (X[:, 2] + 400.0) / 900.0)))
	return model


def eismint_dome_model(n=(14, 14, 4), L=750000.0):
	"""BASELINE config 3 at test size: EISMINT-II-style dome (simulations/eismint_ii/
	eismint_ii_a.py:50-65, gen_vars.py:7-12): flat bed, beta = 1e9 (frozen), Vialov-type surface
	clipped to 1 m thickness, non-periodic, use_pressure_bc=False, nodal A(T) from a depth-linear
	temperature (our synthetic choices)."""
	mesh = BoxMesh(Point(-L, -L, 0), Point(L, L, 1), *n)
	model = D3Model(mesh, use_periodic=False)
	H0, R = 3500.0, 700000.0
	def S(x, y, z):
		r = np.minimum(np.sqrt(x ** 2 + y ** 2) / R, 1.0)
		return np.maximum(H0 * (1.0 - r ** (4.0 / 3.0)) ** (3.0 / 8.0), 1.0)
	B = lambda x, y, z: 0.0 * x
	model.calculate_boundaries()
	model.deform_mesh_to_geometry(S, B)
	model.init_beta(1e9)
	X = mesh.coordinates()
	Sv = model.vertex_field(model.S)
	Ts = 238.15 + 1.67e-5 * np.sqrt(X[:, 0] ** 2 + X[:, 1] ** 2)
	depth = np.clip((Sv - X[:, 2]) / np.maximum(Sv, 1.0), 0.0, 1.0)
	Tm = 273.15 - 9.8e-8 * 910.0 * 9.80665 * Sv
	model.init_A(model.rate_factor(Ts + (Tm - Ts) * depth))
	return model


def greenland_like_model(n=(12, 16, 4)):
	"""BASELINE config 4 at test size: an elliptical dome 1200 x 2400 km on a rough bed with a
	marine margin below sea level (D > 0, dGamma_lt water pressure, relax 0.7), smooth
	beta in [10, 1e4], nodal A(T) -- entirely synthetic."""
	Lx, Ly = 600000.0, 1200000.0
	mesh = BoxMesh(Point(-Lx, -Ly, 0), Point(Lx, Ly, 1), *n)
	model = D3Model(mesh, use_periodic=False)
	def S(x, y, z):
		r2 = np.minimum((x / (0.95 * Lx)) ** 2 + (y / (0.95 * Ly)) ** 2, 1.0)
		return np.maximum(3200.0 * (1.0 - r2) ** 0.5, 60.0)
	def B(x, y, z):
		r2 = (x / Lx) ** 2 + (y / Ly) ** 2
		return 200.0 * np.sin(7 * x / Lx) * np.cos(5 * y / Ly) - 900.0 * r2 ** 2 - 50.0
	model.calculate_boundaries()
	model.deform_mesh_to_geometry(S, B)
	model.init_beta(lambda x, y, z: 10.0 ** (1.0 + 3.0 * (0.5 + 0.5 * np.sin(4 * x / Lx) * np.cos(3 * y / Ly))))
	X = mesh.coordinates()
	Sv = model.vertex_field(model.S)
	Bv = model.vertex_field(model.B)
	depth = np.clip((Sv - X[:, 2]) / np.maximum(Sv - Bv, 1.0), 0.0, 1.0)
	model.init_A(model.rate_factor(243.0 + 28.0 * depth))
	return model
