"""Run under torchrun on >= 2 GPUs (not collected by pytest): the NCCL path of every solve
against the single-rank result of the same problem computed on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port 29533 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault('CSLVR_B200_QUIET', '1')


def _fresh_id(capi, dist, dev, rank):
	ids = [capi.comm_unique_id() if rank == 0 else None]
	dist.broadcast_object_list(ids, src=0, device=dev)
	return ids[0]


def main():
	import torch
	import torch.distributed as dist
	from cases import ismip_model, product_context, relerr
	from cslvr_b200 import capi
	from cslvr_b200.partition import partition_model, rank_context, local_vector
	rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
	torch.cuda.set_device(local)
	dev = torch.device('cuda', local)
	dist.init_process_group('nccl', device_id=dev)
	m = ismip_model('C', L=20000.0, n=(24, 10, 4))
	kw = dict(relative_tolerance=1e-11, maximum_iterations=40, krylov_rtol=1e-12)
	rm = partition_model(m, world, rank)
	c = rank_context(m, rm, device=local, **kw)
	c.set_field(capi.FIELD_B, m.vertex_field(m.B)[rm.vertex_global])
	ids = [capi.comm_unique_id() if rank == 0 else None]
	dist.broadcast_object_list(ids, src=0, device=dev)
	c.comm_init(ids[0], rank, world)
	ul = np.full(2 * rm.n_nodes, 3e-16)
	st = c.newton_solve(ul)
	p = c.solve_pressure(ul)
	uz = c.solve_vert_velocity(ul)
	# single-rank reference on every rank's own GPU (cheap at this size)
	s = product_context(m, device=local, **kw)
	s.set_field(capi.FIELD_B, m.vertex_field(m.B))
	ug = np.full(2 * m.n_nodes, 3e-16)
	sg = s.newton_solve(ug)
	pg = s.solve_pressure(ug)
	uzg = s.solve_vert_velocity(ug)
	own = slice(0, 2 * rm.n_owned)
	e = [relerr(ul[own], local_vector(rm, ug)[own]), relerr(p, pg[rm.vertex_global]), relerr(uz, uzg[rm.vertex_global])]
	t = torch.tensor(e, device=dev, dtype=torch.float64)
	dist.all_reduce(t, op=dist.ReduceOp.MAX)
	if rank == 0:
		print("multi-GPU check, %d ranks: newton its %d vs %d, |du| %.2e |dp| %.2e |du_z| %.2e" % (world, st['iterations'], sg['iterations'], *t.tolist()))
		assert st['converged'] and abs(st['iterations'] - sg['iterations']) <= 1
		assert t[0] < 1e-8 and t[1] < 1e-8 and t[2] < 1e-7, t.tolist()
		print("OK")
	c.close(); s.close()
	# ---- the reduced-Stokes model over NCCL: home-rolled Newton with the u_z callback, then its pressure
	kr = dict(relative_tolerance=1e-9, maximum_iterations=50, relaxation_parameter=0.8, krylov_rtol=1e-13, reduced_stokes=True)
	c = rank_context(m, rm, device=local, **kr)
	c.set_field(capi.FIELD_B, m.vertex_field(m.B)[rm.vertex_global])
	c.comm_init(ids[0] if world == 1 else _fresh_id(capi, dist, dev, rank), rank, world)
	ul, uzl = np.zeros(2 * rm.n_nodes), np.zeros(rm.coords.shape[0])
	st = c.rs_newton_solve(ul, uzl, atol=1e-6, callback=True)
	pl = c.rs_solve_pressure(ul, np.zeros(2 * rm.n_nodes))
	s = product_context(m, device=local, **kr)
	s.set_field(capi.FIELD_B, m.vertex_field(m.B))
	ug, uzg = np.zeros(2 * m.n_nodes), np.zeros(m.num_vertices)
	sg = s.rs_newton_solve(ug, uzg, atol=1e-6, callback=True)
	pg = s.rs_solve_pressure(ug, np.zeros(2 * m.n_nodes))
	e = [relerr(ul[own], local_vector(rm, ug)[own]), relerr(pl, pg[rm.vertex_global]), relerr(uzl, uzg[rm.vertex_global])]
	t = torch.tensor(e, device=dev, dtype=torch.float64)
	dist.all_reduce(t, op=dist.ReduceOp.MAX)
	if rank == 0:
		print("reduced-Stokes, %d ranks: newton its %d vs %d, |du| %.2e |dp| %.2e |du_z| %.2e" % (world, st['iterations'], sg['iterations'], *t.tolist()))
		assert st['converged'] and st['iterations'] == sg['iterations']
		assert t[0] < 1e-8 and t[1] < 1e-8 and t[2] < 1e-7, t.tolist()
		print("OK")
	c.close(); s.close()
	dist.barrier()
	dist.destroy_process_group()


if __name__ == '__main__':
	main()
