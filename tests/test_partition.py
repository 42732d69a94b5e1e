"""Host-side partition logic (no GPU): ownership, ghost/send symmetry, and a
world_size-2 gloo run of a partitioned Jacobi-PCG whose compute is the oracle's."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from cases import ismip_model, marine_model, oracle_problem, state, relerr
from cslvr_b200.partition import partition_model, local_vector, gather_owned


@pytest.mark.parametrize('builder,nr', [(lambda: ismip_model('A', n=(8, 5, 3)), 2),
                                        (lambda: ismip_model('A', n=(9, 4, 2)), 3),
                                        (lambda: marine_model(n=(8, 4, 3)), 4)])
def test_partition_is_consistent(builder, nr):
	m = builder()
	rms = partition_model(m, nr)
	owned = np.concatenate([rm.node_global[:rm.n_owned] for rm in rms])
	assert np.array_equal(np.sort(owned), np.arange(m.n_nodes))            # disjoint cover
	cells = np.concatenate([rm.cell_global for rm in rms])
	assert np.array_equal(np.unique(cells), np.arange(m.num_cells))        # every tet somewhere
	for r, rm in enumerate(rms):
		p = rm.partition
		assert p['recv_ptr'][-1] == rm.n_nodes - rm.n_owned
		for k, nb in enumerate(p['neighbor_rank']):
			o = rms[nb].partition
			kk = list(o['neighbor_rank']).index(r)
			mine = rm.node_global[rm.n_owned + p['recv_ptr'][k]: rm.n_owned + p['recv_ptr'][k + 1]]
			theirs = rms[nb].node_global[o['send_nodes'][o['send_ptr'][kk]:o['send_ptr'][kk + 1]]]
			assert np.array_equal(mine, theirs)                            # same nodes, same order
		# every local tet touches an owned node; every owned row sees all its tets
		assert (rm.vertex_to_node[rm.cells] < rm.n_owned).any(axis=1).all()
		cn = m.vertex_to_node[m.mesh.cells()]
		n_inc_global = np.bincount(cn.ravel(), minlength=m.n_nodes)[rm.node_global[:rm.n_owned]]
		n_inc_local = np.bincount(rm.vertex_to_node[rm.cells].ravel(), minlength=rm.n_nodes)[:rm.n_owned]
		assert np.array_equal(n_inc_global, n_inc_local)


def _local_operator(model, rm, u_global):
	"""owned rows of the oracle Jacobian/residual in the rank's local numbering."""
	P = oracle_problem(model)
	J = P.jacobian_matrix(u_global).tocsr()
	F = P.residual(u_global)
	dof_g = (2 * rm.node_global[:, None] + np.arange(2)).ravel()
	return J[dof_g[:2 * rm.n_owned]][:, dof_g], F[dof_g[:2 * rm.n_owned]]


def _worker(rank, world, port, q):
	import torch.distributed as dist
	os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), CSLVR_B200_QUIET='1')
	dist.init_process_group('gloo', rank=rank, world_size=world)
	import torch
	m = ismip_model('A', n=(8, 5, 3))
	rm = partition_model(m, world, rank)
	u = state(2 * m.n_nodes, 3)
	A, b = _local_operator(m, rm, u)
	no = 2 * rm.n_owned
	p = rm.partition

	def halo(v):                       # v: local vector (owned + ghost), ghosts refreshed in place
		reqs, bufs = [], []
		for k, nb in enumerate(p['neighbor_rank']):
			s = p['send_nodes'][p['send_ptr'][k]:p['send_ptr'][k + 1]]
			sb = torch.from_numpy(np.ascontiguousarray(v.reshape(-1, 2)[s]))
			rb = torch.empty((int(p['recv_ptr'][k + 1] - p['recv_ptr'][k]), 2), dtype=torch.float64)
			reqs += [dist.isend(sb, int(nb)), dist.irecv(rb, int(nb))]
			bufs.append((k, rb))
		for r in reqs:
			r.wait()
		for k, rb in bufs:
			v.reshape(-1, 2)[rm.n_owned + p['recv_ptr'][k]: rm.n_owned + p['recv_ptr'][k + 1]] = rb.numpy()

	def dot(a, c):
		t = torch.tensor([float(a @ c)], dtype=torch.float64)
		dist.all_reduce(t)
		return t.item()

	# Jacobi-PCG on the partitioned operator
	d = A[:, :no].diagonal()
	x = np.zeros(2 * rm.n_nodes)
	r = b.copy()
	z = r / d
	pv = np.zeros(2 * rm.n_nodes)
	pv[:no] = z
	rz = dot(r, z)
	for it in range(2000):
		halo(pv)
		Ap = A @ pv
		alpha = rz / dot(pv[:no], Ap)
		x[:no] += alpha * pv[:no]
		r -= alpha * Ap
		z = r / d
		rz1 = dot(r, z)
		if np.sqrt(dot(r, r)) < 1e-12 * np.sqrt(dot(b, b)):
			break
		pv[:no] = z + (rz1 / rz) * pv[:no]
		rz = rz1
	q.put((rank, rm.node_global[:rm.n_owned].copy(), x[:no].copy(), it))
	dist.barrier()
	dist.destroy_process_group()


def test_gloo_world2_partitioned_pcg_matches_serial():
	import torch.multiprocessing as mp
	ctx = mp.get_context('spawn')
	q = ctx.Queue()
	port = 29500 + (os.getpid() % 2000)
	procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
	for p in procs:
		p.start()
	outs = [q.get(timeout=120) for _ in range(2)]
	for p in procs:
		p.join(timeout=60)
		assert p.exitcode == 0
	m = ismip_model('A', n=(8, 5, 3))
	P = oracle_problem(m)
	u = state(2 * m.n_nodes, 3)
	import scipy.sparse.linalg as spla
	ref = spla.spsolve(P.jacobian_matrix(u).tocsc(), P.residual(u))
	x = np.zeros((m.n_nodes, 2))
	for rank, nodes, xl, it in outs:
		x[nodes] = xl.reshape(-1, 2)
	assert relerr(x.ravel(), ref) < 1e-8


@pytest.mark.parametrize('nr', [2, 3, 4])
def test_direct_slab_builder_equals_generic_partitioner(nr):
	"""bench.py's per-rank mesh generator must give exactly what partitioning the
	global mesh gives (node ids, cells, halo lists, basal facets, fields)."""
	from cslvr_b200.partition import boxmesh_slab
	L, n = 20000.0, (9, 4, 3)
	m = ismip_model('C', L=L, n=n)
	a = 0.1 * np.pi / 180
	S = lambda x, y: -x * np.tan(a)
	B = lambda x, y: -x * np.tan(a) - 1000.0
	beta = lambda x, y: 1000.0 + 1000.0 * np.sin(2 * np.pi * x / L) * np.sin(2 * np.pi * y / L)
	rms = partition_model(m, nr)
	for r in range(nr):
		s, g = boxmesh_slab((L, L), n, r, nr, S, B, beta, 1e-16), rms[r]
		assert np.array_equal(s.node_global, g.node_global) and np.array_equal(s.vertex_global, g.vertex_global)
		assert np.abs(s.coords - g.coords).max() < 1e-9
		srt = lambda c: np.sort(np.ascontiguousarray(c).view('i4,i4,i4,i4'), axis=0)
		assert np.array_equal(srt(s.cells), srt(g.cells))
		for k in ('neighbor_rank', 'send_ptr', 'recv_ptr', 'send_nodes'):
			assert np.array_equal(s.partition[k], g.partition[k]), k
		fs = {(tuple(s.cells[c]), l) for c, l in zip(s.facets[0], s.facets[1])}
		fg = {(tuple(g.cells[c]), l) for c, l, mk in zip(*g.facets) if mk in (3, 5)}
		assert fs == fg
		for k in ('S', 'A', 'beta'):
			assert np.abs(s.fields[k] - g.fields[k]).max() <= 1e-9 * np.abs(g.fields[k]).max()
