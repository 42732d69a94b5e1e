"""The tiled cell kernel (cslvr_b200/csrc/bp_tiles.cu) without a GPU: its index tables -- mesh structure
detection, tiles, prism-column threads, the (thread, staged slot) lists of every Jacobian block -- and its
element math are walked by the library's host emulator (bp_debug_tiles_emulate: same code as the kernel,
compiled for the host) and compared with the oracle's CELL integrals: F and every J block to 1e-12 on
periodic / non-periodic BoxMeshes, an unstructured extruded mesh with shuffled vertex ids, and the slabs of
a partitioned mesh (owned rows only).  The GPU parity tests then only have to confirm the CUDA side."""
import numpy as np
import pytest
import scipy.sparse as sp

from cases import ismip_model, marine_model, delaunay_extruded_model, oracle_problem, relerr
from cslvr_b200 import capi
from cslvr_b200.quadrature import tetrahedron_rule
from oracle import bp_oracle as bo


def _params(m):
	return dict(n=m.n, eps_reg=m.eps_reg, rho_i=m.rho_i, rhosw=m.rhosw, g=m.g, periodic=m.use_periodic, use_pressure_bc=True)


def _ainv(A_vertex, cells, n=3.0, degree=5):
	pts, wts = tetrahedron_rule(degree)
	return ((A_vertex[cells] @ np.asarray(pts).T) ** (-1.0 / n)) @ np.asarray(wts)


def _cell_oracle(P):
	"""the oracle's problem with the facet integrals switched off (the tiled kernel does the cell integrals)."""
	z = np.zeros(0, dtype=np.int64)
	return bo.BPProblem(P.coords, P.cells, P.cell_dofs, P.n_dofs, z, z, z, P.S, P.A, P.beta, periodic=P.periodic,
	                    use_pressure_bc=True, consts=P.c)


def _blocks_to_csr(Jb, rowptr, col, n_dofs):
	rows = np.repeat(np.arange(rowptr.size - 1), np.diff(rowptr))
	R = 2 * rows[:, None, None] + np.array([0, 1])[None, :, None] + np.zeros((1, 1, 2), dtype=np.int64)
	Cc = 2 * col[:, None, None] + np.array([0, 1])[None, None, :] + np.zeros((1, 2, 1), dtype=np.int64)
	return sp.csr_matrix((Jb.ravel(), (R.ravel(), Cc.ravel())), shape=(n_dofs, n_dofs))


CASES = {
	'ismip_a_15x15x5_periodic': lambda: ismip_model('A'),
	'ismip_a_40x33x4_periodic_many_tiles': lambda: ismip_model('A', n=(40, 33, 4)),
	'periodic_3x3x2_columns_meet_themselves': lambda: ismip_model('A', n=(3, 3, 2)),
	'marine_nonperiodic_variable_A': lambda: marine_model(),
	'marine_37x29x6': lambda: marine_model(n=(37, 29, 6)),
	'delaunay_extruded_shuffled_ids': lambda: delaunay_extruded_model(),
	'delaunay_extruded_700_points': lambda: delaunay_extruded_model(n_pts=700, n_layers=5),
}


@pytest.mark.parametrize('name', sorted(CASES))
def test_tile_tables_and_element_math_match_the_oracle(name):
	m = CASES[name]()
	P = oracle_problem(m)
	h = capi.HostTileContext(m.mesh.coordinates(), m.mesh.cells(), 2 * m.n_nodes, (m.facet_cell, m.facet_local, m.ff), _params(m),
	                         vertex_to_node=m.vertex_to_node if m.use_periodic else None)
	info = h.info()
	assert info is not None, h.why_not()
	assert m.num_cells % (3 * info['n_prism_columns']) == 0                        # 3 tets per prism, whole layers
	assert 1.0 <= info['n_threads_used'] / info['n_prism_columns'] < 1.8           # halo factor
	u = np.random.default_rng(3).normal(size=2 * m.n_nodes) * 30.0
	F, Jb, rowptr, col = h.emulate(m.vertex_field(m.S), u, _ainv(m.vertex_field(m.A), m.mesh.cells()))
	P0 = _cell_oracle(P)
	ip, ix = P0.pattern()
	Jo = sp.csr_matrix((P0.jacobian(u), ix, ip), shape=(P.n_dofs, P.n_dofs))
	Je = _blocks_to_csr(Jb, rowptr, col, P.n_dofs)
	assert relerr(F, P0.residual(u)) < 1e-12
	assert abs(Je - Jo).max() < 1e-12 * abs(Jo).max()
	# linear=True: viscosity frozen, no eta' term, F = the u-independent part
	Fp, Jp, _, _ = h.emulate(m.vertex_field(m.S), u, _ainv(m.vertex_field(m.A), m.mesh.cells()), picard=True)
	assert relerr(Fp, P0.residual(np.zeros_like(u))) < 1e-12                       # F(0) is exactly the gravity term
	Jps = _blocks_to_csr(Jp, rowptr, col, P.n_dofs)
	assert abs(Jps - Jps.T).max() < 1e-12 * abs(Jps).max()
	# the kernel folds the basal facets (dGamma_b: sliding, dGamma_bw: water pressure) into its first layer -- the bottom
	# triangle of a prism column is its basal facet: cells + those facets against the oracle with exactly them
	basal = np.isin(m.ff, (3, 5))
	assert basal.any()
	Pb = bo.BPProblem(P.coords, P.cells, P.cell_dofs, P.n_dofs, m.facet_cell[basal], m.facet_local[basal], m.ff[basal], P.S, P.A, P.beta,
	                  periodic=P.periodic, use_pressure_bc=True, consts=P.c)
	Ff, Jf, _, _ = h.emulate(m.vertex_field(m.S), u, _ainv(m.vertex_field(m.A), m.mesh.cells()), beta_vertex=m.vertex_field(m.beta))
	Jbo = sp.csr_matrix((Pb.jacobian(u), ix, ip), shape=(P.n_dofs, P.n_dofs))
	assert relerr(Ff, Pb.residual(u)) < 1e-12
	assert abs(_blocks_to_csr(Jf, rowptr, col, P.n_dofs) - Jbo).max() < 1e-12 * abs(Jbo).max()
	assert relerr(Ff, F) > 1e-6                                          # the facets do contribute
	h.close()


def test_tiles_on_the_slabs_of_a_partitioned_mesh():
	"""owned rows of every slab = the same rows of the unpartitioned problem."""
	from cslvr_b200.partition import partition_model, local_vector
	m = ismip_model('A', n=(24, 9, 3))
	P0 = _cell_oracle(oracle_problem(m))
	u = np.random.default_rng(5).normal(size=2 * m.n_nodes) * 20.0
	Fo = P0.residual(u)
	ip, ix = P0.pattern()
	Jo = sp.csr_matrix((P0.jacobian(u), ix, ip), shape=(P0.n_dofs, P0.n_dofs))
	for rm in partition_model(m, 3):
		h = capi.HostTileContext(rm.coords, rm.cells, 2 * rm.n_nodes, rm.facets, _params(m), vertex_to_node=rm.vertex_to_node, partition=rm.partition)
		assert h.info() is not None, h.why_not()
		F, Jb, rowptr, col = h.emulate(rm.fields['S'], local_vector(rm, u), _ainv(rm.fields['A'], rm.cells))
		assert rowptr.size - 1 == rm.n_owned
		g = rm.node_global
		gd = np.stack([2 * g, 2 * g + 1], axis=1).ravel()                   # local dof -> global dof
		assert relerr(F, Fo[gd[:2 * rm.n_owned]]) < 1e-12
		Je = _blocks_to_csr(Jb, rowptr, col, 2 * rm.n_nodes).tocoo()
		Jg = sp.csr_matrix((Je.data, (gd[Je.row], gd[Je.col])), shape=Jo.shape)
		own = np.zeros(Jo.shape[0], dtype=bool)
		own[gd[:2 * rm.n_owned]] = True
		D = sp.diags(own.astype(float))
		assert abs(Jg - D @ Jo).max() < 1e-12 * abs(Jo).max()
		h.close()


def test_meshes_that_are_not_extruded_prisms_fall_back_to_the_row_gather():
	from cslvr_b200 import Mesh
	from scipy.spatial import Delaunay
	rng = np.random.default_rng(11)
	pts = rng.uniform(size=(60, 3))
	cells = np.sort(Delaunay(pts).simplices, axis=1)
	z = np.zeros(0, dtype=np.int32)
	h = capi.HostTileContext(pts, cells, 2 * pts.shape[0], (z, z, z), dict(periodic=False))
	assert h.info() is None and 'column' in h.why_not()
	h.close()
