"""bench.py's N > 1 arm: one rank per GPU (torch.distributed.run), NCCL only for the
ghost-node exchange and the scalar reductions.  Weak scaling: every rank owns a slab of
260 ice-column rows of the ISMIP-HOM C domain extended in x (2.03 M tets per GPU)."""
import json
import os
import time

import numpy as np


def run_multi(args, rank, world, local):
	import torch
	import torch.distributed as dist
	from cslvr_b200 import capi
	from cslvr_b200.partition import boxmesh_slab
	import bench as B
	dev = torch.device('cuda', local)
	if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
		os.environ['NCCL_DEBUG'] = 'WARN'          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
	dist.init_process_group('nccl', device_id=dev)
	NXT = args.nx if args.nx > 0 else B.NX * world
	NYT = args.ny if args.ny > 0 else B.NY
	if args.exp == 'A':           # ISMIP-HOM A, L = 80 km (docs/get_started.rst:13-45 with L = 80 km), any resolution
		Lx = Ly = 80000.0
		a = 0.5 * np.pi / 180
		S = lambda x, y: -x * np.tan(a)
		Bed = lambda x, y: -x * np.tan(a) - 1000.0 + 500.0 * np.sin(2 * np.pi * x / Lx) * np.sin(2 * np.pi * y / Ly)
		beta = lambda x, y: 1000.0 + 0.0 * x
		wname = "ISMIP-HOM A (L=80 km) MomentumDukowiczBP, periodic BoxMesh %dx%dx%d = %.2f M tets in %d x-slabs" % (NXT, NYT, B.NZ, 6e-6 * NXT * NYT * B.NZ, world)
	else:
		Lx, Ly = B.L * NXT / B.NX, B.L * NYT / B.NY
		a = 0.1 * np.pi / 180
		S = lambda x, y: -x * np.tan(a)
		Bed = lambda x, y: -x * np.tan(a) - 1000.0
		beta = lambda x, y: 1000.0 + 1000.0 * np.sin(2 * np.pi * x / B.L) * np.sin(2 * np.pi * y / B.L)
		wname = B.workload_name(world) if (args.nx <= 0 and args.ny <= 0) else \
			"ISMIP-HOM C (L=20 km per 260 columns) MomentumDukowiczBP, periodic BoxMesh %dx%dx%d = %.2f M tets in %d x-slabs" % (NXT, NYT, B.NZ, 6e-6 * NXT * NYT * B.NZ, world)
	rm = boxmesh_slab((Lx, Ly), (NXT, NYT, B.NZ), rank, world, S, Bed, beta, 1e-16)
	params = dict(n=3.0, eps_reg=1e-15, rho_i=910.0, rhosw=1028.0, g=9.80665, periodic=True, use_pressure_bc=True)
	ctx = capi.BPContext(rm.coords, rm.cells, 2 * rm.n_nodes, rm.facets, params,
	                     vertex_to_node=rm.vertex_to_node, partition=rm.partition, device=local)
	for f, k in ((capi.FIELD_S, 'S'), (capi.FIELD_A, 'A'), (capi.FIELD_BETA, 'beta')):
		ctx.set_field(f, rm.fields[k])
	torch.cuda.set_stream(torch.cuda.Stream(dev))       # capturable (the legacy default stream is not), and torch.cuda.Event sees it
	ctx.use_torch_stream()
	clocks = B.ClockSampler(local) if rank == 0 else None      # started early: nvidia-smi needs a moment before its first sample
	ids = [capi.comm_unique_id() if rank == 0 else None]
	dist.broadcast_object_list(ids, src=0, device=dev)
	ctx.comm_init(ids[0], rank, world)
	nt_global = 6 * NXT * NYT * B.NZ

	# velocity state: a sheared sliding flow, deterministic in the global node id
	X = np.empty((rm.n_nodes, 3))
	X[rm.vertex_to_node] = rm.coords
	Sn = np.empty(rm.n_nodes); Sn[rm.vertex_to_node] = rm.fields['S']
	zrel = np.clip((X[:, 2] - (Sn - 1000.0)) / 1000.0, 0.0, 1.0)
	u_host = np.empty(2 * rm.n_nodes)
	u_host[0::2] = 15.0 + 10.0 * np.sin(2 * np.pi * X[:, 0] / B.L) * np.sin(2 * np.pi * X[:, 1] / B.L) + 20.0 * np.sqrt(zrel)
	u_host[1::2] = 2.0 * np.cos(2 * np.pi * X[:, 0] / B.L) * zrel
	u_host *= 1.0 + 1e-3 * np.sin(0.37 * rm.node_global.repeat(2))

	ctx.assemble(u_host, want_F=True, want_J=True, J=False)          # untimed: first halo exchange sets up the NCCL channels
	newton = None
	if not args.no_newton:
		pck = {'preconditioner': args.pc} if args.pc else {}
		ctx.set_params(relative_tolerance=1e-9, maximum_iterations=1, relaxation_parameter=1.0, krylov_rtol=1e-5, krylov_maxit=8, **pck)
		ctx.newton_solve(torch.full((2 * rm.n_nodes,), 3e-16, dtype=torch.float64, device=dev))   # untimed: sets up every NCCL collective used
		ctx.set_params(relative_tolerance=1e-9, maximum_iterations=25, relaxation_parameter=1.0, krylov_rtol=1e-5, krylov_maxit=10000)
		un = torch.full((2 * rm.n_nodes,), 3e-16, dtype=torch.float64, device=dev)
		dist.barrier(); torch.cuda.synchronize()
		t0 = time.perf_counter()
		st = ctx.newton_solve(un)
		torch.cuda.synchronize(); dist.barrier()
		t = torch.tensor([time.perf_counter() - t0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
		newton = {"solve_s": t.item(), "converged": st['converged'], "newton_iterations": st['iterations'],
		          "krylov_iterations": st['krylov_iterations'], "assemble_ms": st['t_assemble_ms'], "krylov_ms": st['t_krylov_ms'],
		          "preconditioner": ctx.pc_label(), "rtol": 1e-9, "krylov_rtol": 1e-5}

	u_dev = torch.from_numpy(u_host).to(dev)
	F_dev = torch.empty_like(u_dev)
	u_pin = torch.from_numpy(u_host).pin_memory()
	F_pin = torch.empty(u_host.size, dtype=torch.float64).pin_memory()
	step_dev = lambda: ctx.assemble(u_dev, want_F=True, want_J=True, F=F_dev, J=False)
	step_e2e = lambda: ctx.assemble(u_pin.numpy(), want_F=True, want_J=True, F=F_pin.numpy(), J=False)

	def timed(step):
		for _ in range(args.warmup):
			step()
		dist.barrier(); torch.cuda.synchronize()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		t0 = time.perf_counter()
		e0.record()
		for _ in range(args.steps):
			step()
		e1.record()
		torch.cuda.synchronize()
		wall = time.perf_counter() - t0
		dist.barrier()
		t = torch.tensor([e0.elapsed_time(e1) * 1e-3, wall], device=dev, dtype=torch.float64)
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
		return t[0].item() / args.steps, t[1].item() / args.steps

	if clocks: clocks.start()
	l0 = ctx.launch_count()
	s_dev, _ = timed(step_dev)
	launches = (ctx.launch_count() - l0) * args.steps // (args.steps + args.warmup)
	what = capi.ASSEMBLE_F | capi.ASSEMBLE_J
	t_cell = ctx.time_assemble(what | 4, repeats=max(min(args.steps, 50), 10), flush_l2=True)
	ctx.assemble(u_dev, want_F=True, want_J=True, F=F_dev, J=False)
	t_spmv = ctx.time_spmv(repeats=50, flush_l2=True)
	if clocks: clocks.stop()
	_, s_e2e = timed(step_e2e)
	tl = torch.tensor([float(launches)], device=dev); dist.all_reduce(tl)
	if rank == 0:
		peaks = {}
		try:
			peaks = json.load(open(os.path.join(B.ROOT, 'MEASURED_PEAKS.json')))
		except (OSError, ValueError):
			pass
		peak = float(peaks.get('hbm_gbs', 6650.0))
		nt, nv, nn = rm.cells.shape[0], rm.coords.shape[0], rm.n_owned
		b_cell = 16 * nt + 32 * nv + 24 * nn + 8 * ctx.nnz + 16 * nn
		b_spmv = 9 * ctx.nnz + 40 * nn
		line = {"metric": B.METRIC, "value": nt_global / s_dev / 1e6, "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
		        "warmup": args.warmup, "ms_per_step": s_dev * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
		        "dtype": "f64", "data": "synthetic",
		        "config": {"workload": wname, "scaling_mode": "weak (260 columns in x per GPU)" if args.nx <= 0 else "fixed total size", "n_tets": nt_global, "tets_per_rank_incl_halo": nt,
		                   "partition": "x-slabs of ice columns, 1 ghost node layer, halo of u by grouped ncclSend/ncclRecv before every assembly",
		                   "l2": "per-rank working set larger than L2; roofline launches flush L2"},
		        "e2e": {"value": nt_global / s_e2e / 1e6, "unit": B.UNIT, "h2d_bytes_per_step": 8 * u_host.size * world,
		                "d2h_bytes_per_step": 8 * u_host.size * world},
		        "gpu_launches": int(tl.item()),
		        "roofline": {"kernel": "k_rows<F,J> (rank 0)", "bound": "hbm", "achieved": b_cell / (t_cell * 1e-3) / 1e9, "peak": peak,
		                     "unit": "GB/s", "frac": b_cell / (t_cell * 1e-3) / 1e9 / peak, "traffic": None,
		                     "peak_source": "measured" if peaks else "fallback", "ms": t_cell},
		        "roofline_spmv": {"kernel": "k_spmv (rank 0)", "bound": "hbm", "achieved": b_spmv / (t_spmv * 1e-3) / 1e9, "peak": peak,
		                          "unit": "GB/s", "frac": b_spmv / (t_spmv * 1e-3) / 1e9 / peak, "ms": t_spmv},
		        "clocks": clocks.summary()}
		if newton:
			line["newton"] = newton
		B.emit(json.dumps(line))
	if clocks: clocks.close()
	ctx.close()
	dist.barrier()
	dist.destroy_process_group()
