"""Column partition of an extruded mesh across ranks (one rank per GPU).

The reference gets its parallelism from dolfin: `mpirun -np N python script.py`
partitions the mesh, PETSc scatters ghost values inside MatMult and all-reduces inside
VecDot (launch lines simulations/verification/bash_jobs.bash:30).
Here every rank owns a slab of ice columns (all nodes above a footprint point stay
together, so the vertical coupling is rank-local) and keeps, besides its owned nodes,
one layer of ghost nodes and every tetrahedron that touches an owned node.  The rows
of the Jacobian and of the residual that belong to owned nodes are then complete
without any matrix-entry communication; only vector ghost values travel:

* before assembly ......... ghost values of u
* before every SpMV ....... ghost values of the CG search direction
* CG / Newton scalars ..... all-reduce of 1-3 doubles

Local numbering handed to ``bp_create``: owned nodes first (ascending global id), then
the ghosts grouped by owning rank (ascending rank, ascending global id); a neighbour's
send list is ordered by global id too, so both sides agree without a handshake.
"""
import numpy as np


def slab_ranks(x_of_node, n_ranks):
	"""owner of every node: equal-count slabs along x (columns never split: all nodes of
	a column share x on the flat mesh)."""
	xs = np.unique(np.round(np.asarray(x_of_node, dtype=np.float64), 9))
	col = np.searchsorted(xs, np.round(x_of_node, 9))
	return ((col * n_ranks) // xs.size).astype(np.int32)


class RankMesh(object):
	"""what one rank passes to the C-ABI (arrays in local numbering)."""
	pass


def partition_mesh(coords, cells, vertex_to_node, n_nodes, facets, fields, node_rank, rank):
	"""local mesh of ``rank``.

	coords[N_v,3], cells[N_t,4], vertex_to_node[N_v], facets = (cell, local, marker),
	fields = dict of vertex-indexed arrays, node_rank[N_n] = owner of every global node.
	"""
	cells = np.asarray(cells)
	v2n = np.asarray(vertex_to_node, dtype=np.int64)
	cn = v2n[cells]                                         # global node ids per cell
	owned_cell = (node_rank[cn] == rank).any(axis=1)
	lc = np.flatnonzero(owned_cell)                         # local cells (global ids, ascending)
	lcells = cells[lc]
	lv = np.unique(lcells)                                  # local vertices (global ids)
	ln_all = np.unique(cn[lc])                              # local nodes (global ids)
	own = ln_all[node_rank[ln_all] == rank]
	gh = ln_all[node_rank[ln_all] != rank]
	gh = gh[np.lexsort((gh, node_rank[gh]))]                # by owner rank, then global id
	node_global = np.concatenate([own, gh])
	g2l_node = np.full(n_nodes, -1, dtype=np.int64)
	g2l_node[node_global] = np.arange(node_global.size)
	g2l_vert = np.full(coords.shape[0], -1, dtype=np.int64)
	g2l_vert[lv] = np.arange(lv.size)

	m = RankMesh()
	m.rank = rank
	m.vertex_global = lv
	m.node_global = node_global
	m.cell_global = lc
	m.coords = np.ascontiguousarray(coords[lv])
	m.cells = g2l_vert[lcells].astype(np.int32)
	m.vertex_to_node = g2l_node[v2n[lv]].astype(np.int32)
	m.n_nodes = int(node_global.size)
	m.n_owned = int(own.size)
	fc, fl, fm = (np.asarray(a) for a in facets)
	g2l_cell = np.full(cells.shape[0], -1, dtype=np.int64)
	g2l_cell[lc] = np.arange(lc.size)
	keep = g2l_cell[fc] >= 0
	m.facets = (g2l_cell[fc[keep]].astype(np.int32), fl[keep].astype(np.int32), fm[keep].astype(np.int32))
	m.fields = {k: np.ascontiguousarray(np.asarray(v, dtype=np.float64)[lv]) for k, v in fields.items()}

	# ghosts I receive, per owning rank
	nbr = np.unique(node_rank[gh]).astype(np.int32) if gh.size else np.zeros(0, dtype=np.int32)
	# nodes I own that other ranks hold as ghosts: owned nodes sharing a cell with a node
	# of that rank -- the mirror image of their ghost definition
	send = {}
	cr = node_rank[cn]                                     # owner of each cell node
	for k in np.unique(cr[(cr == rank).any(axis=1)]):
		if k == rank:
			continue
		touch_k = (cr == k).any(axis=1)                    # cells rank k keeps (they touch a k-node)
		mine = cn[touch_k][cr[touch_k] == rank]
		send[int(k)] = np.unique(mine)
	all_nbr = np.array(sorted(set(nbr.tolist()) | set(send.keys())), dtype=np.int32)
	recv_ptr, send_ptr, send_nodes = [0], [0], []
	for k in all_nbr:
		recv_ptr.append(recv_ptr[-1] + int((node_rank[gh] == k).sum()))
		s = send.get(int(k), np.zeros(0, dtype=np.int64))
		send_nodes.append(g2l_node[s])
		send_ptr.append(send_ptr[-1] + s.size)
	m.partition = dict(n_owned_nodes=m.n_owned, neighbor_rank=all_nbr,
	                   send_ptr=np.array(send_ptr, dtype=np.int64),
	                   send_nodes=(np.concatenate(send_nodes) if send_nodes else np.zeros(0)).astype(np.int32),
	                   recv_ptr=np.array(recv_ptr, dtype=np.int64))
	return m


def partition_model(model, n_ranks, rank=None):
	"""partition a :class:`cslvr_b200.D3Model` into x-slabs of columns; returns the
	:class:`RankMesh` of ``rank`` or the list for all ranks."""
	xflat = model.flat_coordinates[model.node_vertex, 0]
	node_rank = slab_ranks(xflat, n_ranks)
	fields = dict(S=model.vertex_field(model.S), A=model.vertex_field(model.A), beta=model.vertex_field(model.beta))
	args = (model.mesh.coordinates(), model.mesh.cells(), model.vertex_to_node, model.n_nodes,
	        (model.facet_cell, model.facet_local, model.ff), fields, node_rank)
	if rank is not None:
		return partition_mesh(*args, rank)
	return [partition_mesh(*args, r) for r in range(n_ranks)]


def rank_context(model, rm, device=0, use_pressure_bc=True, **params):
	"""a ``BPContext`` for one :class:`RankMesh` with its coefficient fields set."""
	from . import capi
	p = dict(n=model.n, eps_reg=model.eps_reg, rho_i=model.rho_i, rhosw=model.rhosw, g=model.g,
	         periodic=model.use_periodic, use_pressure_bc=use_pressure_bc)
	p.update(params)
	c = capi.BPContext(rm.coords, rm.cells, 2 * rm.n_nodes, rm.facets, p,
	                   vertex_to_node=rm.vertex_to_node, partition=rm.partition, device=device)
	c.set_field(capi.FIELD_S, rm.fields['S'])
	c.set_field(capi.FIELD_A, rm.fields['A'])
	c.set_field(capi.FIELD_BETA, rm.fields['beta'])
	return c


def local_vector(rm, u_global):
	"""global dof vector (2*node + c) -> this rank's local vector (owned + ghost)."""
	u = np.asarray(u_global).reshape(-1, 2)
	return np.ascontiguousarray(u[rm.node_global].ravel())


def gather_owned(rms, locals_, n_nodes):
	"""assemble a global dof vector from every rank's owned entries."""
	out = np.zeros((n_nodes, 2))
	for rm, v in zip(rms, locals_):
		out[rm.node_global[:rm.n_owned]] = np.asarray(v).reshape(-1, 2)[:rm.n_owned]
	return out.ravel()


def boxmesh_slab(L, n, rank, n_ranks, S, B, beta, A):
	"""the :class:`RankMesh` of ``rank`` for a *periodic* deformed ``BoxMesh((0,0,0),
	(Lx,Ly,1), nx,ny,nz)`` built directly from the slab's hexahedra -- same result as
	``partition_model`` on the global mesh (tests/test_partition.py) without ever
	holding the global mesh, which is what lets 8 ranks set up a 16-40 M-tet job.

	S, B, beta, A: callables f(x, y) (A may be a number); geometry as in
	cslvr/d3model.py:629-667, markers as :337-475 (all grounded => basal marker 3).
	"""
	(Lx, Ly), (nx, ny, nz) = L, n
	col_rank = (np.arange(nx) * n_ranks) // nx
	mine = np.flatnonzero(col_rank == rank)
	a, b = int(mine[0]), int(mine[-1]) + 1
	hx = (np.arange(a - 1, b)) % nx if n_ranks > 1 else np.arange(nx)       # hex columns touching owned nodes
	if n_ranks > 1 and b - a + 1 > nx:
		hx = np.arange(nx)
	sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
	base = (np.arange(nz)[:, None, None] * sz + np.arange(ny)[None, :, None] * sy + hx[None, None, :]).reshape(-1, 1)
	off = np.array([0, sx, sy, sx + sy, sz, sz + sx, sz + sy, sz + sx + sy])
	split = np.array([[0, 1, 3, 7], [0, 1, 7, 5], [0, 5, 7, 4], [0, 3, 2, 7], [0, 6, 4, 7], [0, 2, 6, 7]])
	gcells = (base[:, :, None] + off[split][None, :, :]).reshape(-1, 4)
	gcells.sort(axis=1)
	lv, inv = np.unique(gcells, return_inverse=True)
	cells = inv.reshape(-1, 4).astype(np.int32)
	ix, iy, iz = lv % (nx + 1), (lv // (nx + 1)) % (ny + 1), lv // sz
	x, y = ix * (Lx / nx), iy * (Ly / ny)
	Sv, Bv = S(x, y), B(x, y)
	z = (iz / nz) * (Sv - Bv) + Bv
	node_g = iz * (ny * nx) + (iy % ny) * nx + (ix % nx)                       # canonical periodic node id
	owner = col_rank[ix % nx]
	ln = np.unique(node_g)
	own_of = np.zeros(ln.size, dtype=np.int64)
	own_of[np.searchsorted(ln, node_g)] = owner
	own = ln[own_of == rank]
	gh = ln[own_of != rank]
	gho = own_of[own_of != rank]
	order = np.lexsort((gh, gho))
	gh, gho = gh[order], gho[order]
	node_global = np.concatenate([own, gh])
	loc = np.empty(ln.size, dtype=np.int64)
	loc[np.searchsorted(ln, node_global)] = np.arange(node_global.size)

	m = RankMesh()
	m.rank = rank
	m.vertex_global, m.node_global = lv, node_global
	m.coords = np.ascontiguousarray(np.stack([x, y, z], axis=1))
	m.cells = cells
	m.vertex_to_node = loc[np.searchsorted(ln, node_g)].astype(np.int32)
	m.n_nodes, m.n_owned = int(node_global.size), int(own.size)
	# basal facets: the two tets of every bottom hexahedron that contain v7 with a face on
	# the bed (tet 0 and tet 3 of the split), the facet opposite their highest vertex
	nb = ny * hx.size
	hex0 = np.arange(nb)
	fcell = np.concatenate([hex0 * 6 + 0, hex0 * 6 + 3])
	fcell.sort()
	m.facets = (fcell.astype(np.int32), np.full(fcell.size, 3, dtype=np.int32), np.full(fcell.size, 3, dtype=np.int32))
	Av = A(x, y) if callable(A) else np.full(lv.size, float(A))
	m.fields = dict(S=np.ascontiguousarray(Sv, dtype=np.float64), A=np.ascontiguousarray(Av, dtype=np.float64),
	                beta=np.ascontiguousarray(beta(x, y) * np.ones(lv.size), dtype=np.float64))
	# halo lists: I send the owned nodes that sit in the first / last owned column to the
	# rank owning the column to the left / right (periodic wrap); both sides order by global id
	nbrs = np.unique(gho).astype(np.int32)
	colg = node_global % nx
	send_nodes, send_ptr, recv_ptr = [], [0], [0]
	for k in nbrs:
		recv_ptr.append(recv_ptr[-1] + int((gho == k).sum()))
		# columns of rank k's hexes touch my node columns: k needs my column a if it owns
		# column a-1, and my column b-1 if it owns column b (its hex column b-1... is mine)
		need = []
		if col_rank[(a - 1) % nx] == k:
			need.append(a % nx)
		kcols = np.flatnonzero(col_rank == k)
		ka = int(kcols[0])
		if (ka - 1) % nx == (b - 1) % nx:
			need.append((b - 1) % nx)
		sel = np.flatnonzero(np.isin(colg[:m.n_owned], np.unique(need)))
		sel = sel[np.argsort(node_global[sel])]
		send_nodes.append(sel)
		send_ptr.append(send_ptr[-1] + sel.size)
	m.partition = dict(n_owned_nodes=m.n_owned, neighbor_rank=nbrs,
	                   send_ptr=np.array(send_ptr, dtype=np.int64),
	                   send_nodes=(np.concatenate(send_nodes) if send_nodes else np.zeros(0)).astype(np.int32),
	                   recv_ptr=np.array(recv_ptr, dtype=np.int64))
	return m
