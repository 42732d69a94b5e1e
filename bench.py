#!/usr/bin/env python
"""bench.py -- BP residual+Jacobian assembly throughput (Mtets/s) and Newton solve time, ISMIP-HOM A.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): "BP residual+Jacobian assembly Mtets/s; Newton solve time, ISMIP-HOM A".
Workload = BASELINE.json configs[4], the weak-scaling series of ISMIP-HOM A (L = 80 km,
docs/get_started.rst:13-45): a periodic BoxMesh of 5 M tetrahedra PER GPU --
408^2, 577^2, 816^2, 1155^2 x 5 layers at N = 1, 2, 4, 8 (N = 8 is the 40 M-tet north-star
case) -- split into N x-slabs of ice columns, one rank per GPU; NCCL only for the ghost-node
exchange of u before assembly and for the SpMV halo / reductions of the Newton solve.
A *step* is one fused assemble(F) + assemble(J) pass over the whole mesh at a fixed velocity.

value       device-resident: u already in HBM, F written to HBM, J stays resident.
e2e         the same pass through the C-ABI with HOST (pinned) buffers: u copied host->device and F
            device->host inside the timed region (J stays resident on the device exactly as it does
            between assemble() and solve() in the reference); e2e_with_J also exports the CSR values
            of J to the host.
roofline    the dominant kernel (cell assembly) timed alone with CUDA events on the library's stream,
            L2 flushed before every launch; algorithmic bytes per SURVEY.md §8d / DESIGN.md §4.
newton      the whole dolfin.solve(F == 0, u, J=J) call from the cold start, reference defaults.
parity      untimed, at FULL size: F, the CSR pattern and every entry of J against the oracle's C twin
            (N = 1), F and J x on every rank's rows (N > 1), and the oracle's own residual at the
            product's converged velocity.  F_rel also covers the F that the TIMED calls left behind: the
            device buffer of the `value` loop and the pinned host buffer of the `e2e` loop.
config.host_binding   every rank runs on the CPU cores next to its GPU (NVML's ideal affinity), as a launcher
            with NUMA binding would arrange; the CPU checker / baseline legs get the full affinity mask back.
cpu_baseline / cpu_baseline_newton / --impl reference
            the oracle's C/OpenMP restatement of the reference's CPU path ("port": the reference is
            Python on an un-vendored FEniCS stack and cannot run here) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault('CSLVR_B200_QUIET', '1')
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')      # stdout carries exactly one JSON line

_JSON_FD = None


def quiet_stdout():
	"""stdout carries exactly ONE JSON line: everything else a library prints there (NCCL's version
	banner, ...) is sent to stderr; emit() writes the line to the real stdout."""
	global _JSON_FD
	if _JSON_FD is None:
		sys.stdout.flush()
		_JSON_FD = os.dup(1)
		os.dup2(2, 1)


def emit(text):
	sys.stdout.flush()
	if _JSON_FD is None:
		print(text, flush=True)
	else:
		os.write(_JSON_FD, (text + "\n").encode())


METRIC = "BP residual+Jacobian assembly throughput"
UNIT = "Mtets/s"
NZ, L = 5, 80000.0
NXY = {1: 408, 2: 577, 4: 816, 8: 1155}          # 5 M tets per GPU (BASELINE.json configs[4])
CONSTS = dict(n=3.0, eps_reg=1e-15, rho_i=910.0, rhosw=1028.0, g=9.80665)
ALPHA = 0.5 * np.pi / 180


def grid(n_gpus):
	if os.environ.get('BENCH_NXY'):                 # tests only: a small mesh for the JSON-contract check
		return int(os.environ['BENCH_NXY'])
	if n_gpus in NXY:
		return NXY[n_gpus]
	return int(round(408 * np.sqrt(n_gpus)))


def S_xy(x, y):
	return -x * np.tan(ALPHA)


def B_xy(x, y):
	return -x * np.tan(ALPHA) - 1000.0 + 500.0 * np.sin(2 * np.pi * x / L) * np.sin(2 * np.pi * y / L)


def beta_xy(x, y):
	return 1000.0 + 0.0 * x


def workload_name(n):
	nxy = grid(n)
	return "ISMIP-HOM A (L=80 km) MomentumDukowiczBP, periodic BoxMesh %dx%dx%d = %.2f M tets%s" % (
		nxy, nxy, NZ, 6e-6 * nxy * nxy * NZ, "" if n == 1 else " in %d x-slabs (%.2f M per GPU)" % (n, 6e-6 * nxy * nxy * NZ / n))


def build_model(n_gpus=1, nxy=None):
	"""ISMIP-HOM A through the host mirror of the reference API (docs/get_started.rst:13-45)."""
	from cslvr_b200 import BoxMesh, Point, D3Model
	nxy = nxy or grid(n_gpus)
	mesh = BoxMesh(Point(0, 0, 0), Point(L, L, 1), nxy, nxy, NZ)
	model = D3Model(mesh, use_periodic=True)
	model.calculate_boundaries()
	model.deform_mesh_to_geometry(lambda x, y, z: S_xy(x, y), lambda x, y, z: B_xy(x, y))
	model.init_beta(1000.0)
	model.init_A(1e-16)
	return model


class Slab(object):
	"""what both arms need of one rank's share of the mesh, in local numbering."""
	pass


def slab_from_model(model):
	s = Slab()
	s.coords, s.cells = model.mesh.coordinates(), model.mesh.cells()
	s.vertex_to_node = model.vertex_to_node
	s.n_nodes = s.n_owned = model.n_nodes
	s.facets = (model.facet_cell, model.facet_local, model.ff)
	s.fields = dict(S=model.vertex_field(model.S), A=model.vertex_field(model.A), beta=model.vertex_field(model.beta))
	s.node_global = np.arange(model.n_nodes)
	s.partition = None
	return s


def slab_of_rank(rank, world):
	from cslvr_b200.partition import boxmesh_slab
	nxy = grid(world)
	rm = boxmesh_slab((L, L), (nxy, nxy, NZ), rank, world, S_xy, B_xy, beta_xy, 1e-16)
	s = Slab()
	s.coords, s.cells, s.vertex_to_node = rm.coords, rm.cells, rm.vertex_to_node
	s.n_nodes, s.n_owned, s.facets, s.fields = rm.n_nodes, rm.n_owned, rm.facets, rm.fields
	s.node_global, s.partition = rm.node_global, rm.partition
	return s


def node_coords(s):
	X = np.empty((s.n_nodes, 3))
	X[s.vertex_to_node] = s.coords
	return X


def synthetic_state(s, seed_scale=1e-3):
	"""a sheared sliding flow of ISMIP-HOM magnitude, perturbed; a function of the GLOBAL node id and the
	node coordinates only, so every rank holds the same values for shared nodes."""
	X = node_coords(s)
	Sn, Bn = S_xy(X[:, 0], X[:, 1]), B_xy(X[:, 0], X[:, 1])
	zrel = np.clip((X[:, 2] - Bn) / np.maximum(Sn - Bn, 1.0), 0.0, 1.0)
	xp = np.where(X[:, 0] >= L - 1e-6, X[:, 0] - L, X[:, 0])       # periodic images agree
	yp = np.where(X[:, 1] >= L - 1e-6, X[:, 1] - L, X[:, 1])
	u = np.empty(2 * s.n_nodes)
	u[0::2] = 15.0 + 10.0 * np.sin(2 * np.pi * xp / L) * np.sin(2 * np.pi * yp / L) + 20.0 * np.sqrt(zrel)
	u[1::2] = 2.0 * np.cos(2 * np.pi * xp / L) * zrel
	return u * (1.0 + seed_scale * np.sin(0.37 * np.repeat(s.node_global, 2) + 0.11 * np.tile([0.0, 1.0], s.n_nodes)))


# ---------------------------------------------------------------- the oracle side (checker / CPU baseline only)
def oracle_for_slab(s, threads=None):
	"""oracle/bp_oracle_c.c on the slab's arrays: rows of owned nodes are complete (every tet touching an
	owned node is local).  Only the parity / cpu_baseline / --impl reference legs call this."""
	from oracle import bp_oracle as bo
	from oracle.bp_oracle_cwrap import COracle
	n2 = 2 * s.vertex_to_node[s.cells]
	cell_dofs = np.concatenate([n2, n2 + 1], axis=1).astype(np.int32)
	fc, fl, ff = s.facets
	P = bo.BPProblem(s.coords, s.cells, cell_dofs, 2 * s.n_nodes, fc, fl, ff, s.fields['S'], s.fields['A'], s.fields['beta'],
	                 periodic=True, use_pressure_bc=True, consts=CONSTS)
	return COracle(P, threads=threads)


def relerr(a, b):
	return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def relerr_entrywise(a, b, floor=1e-10):
	"""largest |a - b| / |b| over the entries with |b| > floor * max|b| (small entries are absorbed by the
	max-norm measure; this one looks at every significant entry by itself)."""
	m = np.abs(b) > floor * np.abs(b).max()
	return float((np.abs(a - b)[m] / np.abs(b)[m]).max()) if m.any() else 0.0


def bind_to_gpu_numa_node(index):
	"""run this rank on the CPU cores next to its GPU (NVML's ideal affinity): the pinned host buffers of the end-to-end
	arm are then allocated on the NUMA node the GPU's PCIe root hangs off, instead of crossing the socket link.  What a
	launcher like `numactl` / `torchrun --numa-binding` does; BENCH_NO_BIND=1 leaves the affinity alone."""
	if os.environ.get('BENCH_NO_BIND'):
		return False
	try:
		import pynvml
		pynvml.nvmlInit()
		pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
		return True
	except Exception:
		return False


def host_threads():
	try:
		return len(os.sched_getaffinity(0))
	except AttributeError:
		return os.cpu_count() or 1


def cpu_newton_baseline(threads, nxy=130):
	"""Newton solve time of the CPU restatement (dolfin NewtonSolver semantics, Jacobi-PCG at KSP
	defaults) on ISMIP-HOM A: the reference's own 15x15x5 case and a 0.5 M-tet mesh."""
	out = []
	for n in (15, nxy):
		s = slab_from_model(build_model(1, nxy=n))
		co = oracle_for_slab(s, threads)
		t0 = time.perf_counter()
		_, st = co.newton()
		out.append({"workload": "ISMIP-HOM A %dx%dx%d = %d tets" % (n, n, NZ, s.cells.shape[0]), "solve_s": time.perf_counter() - t0,
		            "converged": st['converged'], "newton_iterations": st['iterations'], "krylov_iterations": st['krylov_iterations'],
		            "solver": "Newton rtol 1e-9 + Jacobi-PCG rtol 1e-5 (oracle/bp_oracle_c.c)", "cores": co.threads, "kind": "port"})
	return out


class ClockSampler(object):
	"""nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
	Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

	def __init__(self, index=0):
		self.p = None
		try:
			self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
			                          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
		except OSError:
			pass
		self.lines = []
		self.on = False
		if self.p:
			self.t = threading.Thread(target=self._read, daemon=True)
			self.t.start()

	def _read(self):
		for ln in self.p.stdout:
			if self.on:
				self.lines.append(ln)

	def start(self): self.on = True
	def stop(self): self.on = False

	def close(self):
		if self.p:
			self.p.terminate()

	def summary(self):
		sm, mx, reasons = [], 0.0, set()
		names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
		for ln in self.lines:
			f = [x.strip() for x in ln.split(',')]
			try:
				sm.append(float(f[0])); mx = max(mx, float(f[1]))
			except (ValueError, IndexError):
				continue
			for nme, v in zip(names, f[3:7]):
				if v.lower().startswith('active'):
					reasons.add(nme)
		return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
		        "reasons": sorted(reasons), "samples": len(sm)}


def measured_traffic(kernel, n_tets):
	"""dram__bytes_read.sum + dram__bytes_write.sum of one launch of ``kernel`` on a mesh of ``n_tets``
	cells, from the committed ncu captures listed in profiles/traffic.json (null when none matches)."""
	try:
		tab = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
	except (OSError, ValueError):
		return None, "profiles/traffic.json (absent)"
	for e in tab.get('entries', []):
		if e.get('kernel') == kernel and int(e.get('n_tets', -1)) == int(n_tets):
			return float(e['dram_bytes']), "profiles/traffic.json <- " + e.get('source', '?')
	return None, "profiles/traffic.json (no capture of %s at %d tets)" % (kernel, n_tets)


# ---------------------------------------------------------------- --impl reference
def run_reference(args):
	"""the reference's CPU implementation of the path: the C/OpenMP port of FFC tabulate_tensor + dolfin
	Assembler (oracle/bp_oracle_c.c) on all host cores.  Sample: rank 0's slab of the same mesh (the
	whole mesh at N = 1), F + J, every step."""
	rank = int(os.environ.get('RANK', '0'))
	if rank != 0:
		return
	threads = host_threads()
	os.environ['OMP_NUM_THREADS'] = str(threads)          # torch.distributed.run exports 1
	s = slab_from_model(build_model(1)) if args.gpus == 1 else slab_of_rank(0, args.gpus)
	u = synthetic_state(s)
	co = oracle_for_slab(s, threads)
	nt = s.cells.shape[0]
	budget = 150.0 / max(args.steps + args.warmup, 1)          # seconds per step
	_, _, t = co.assemble(u, n_cells=min(nt, 200000))
	ns = int(min(nt, max(100000, min(nt, 200000) / t * budget)))
	for _ in range(args.warmup):
		co.assemble(u, n_cells=ns)
	ts = []
	for _ in range(args.steps):
		_, _, t = co.assemble(u, n_cells=ns)
		ts.append(t)
	t = float(np.mean(ts))
	nt_global = 6 * grid(args.gpus) ** 2 * NZ
	base = {"value": ns / t / 1e6, "unit": UNIT, "cores": co.threads, "kind": "port",
	        "sample": "%d of the %d cells of %s per step (F + J, %d-point quadrature loop, CSR insertion by binary search, %d OpenMP threads, mean of %d steps)"
	                  % (ns, nt_global, "the mesh" if args.gpus == 1 else "the mesh: the first cells of rank 0's slab", len(co.q[1]), co.threads, args.steps)}
	line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
	        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
	        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
	        "config": {"workload": workload_name(args.gpus), "n_tets": nt_global, "sample_cells_per_step": ns},
	        "cpu_baseline": base,
	        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
	        "gpu_launches": 0}
	if not args.no_newton:
		line["cpu_baseline_newton"] = cpu_newton_baseline(threads, nxy=int(os.environ.get('BENCH_NEWTON_NXY', '130')))
	emit(json.dumps(line))


# ---------------------------------------------------------------- ours
def run_ours(args):
	import torch
	from cslvr_b200 import capi
	world = int(os.environ.get('WORLD_SIZE', '1'))
	rank = int(os.environ.get('RANK', '0'))
	local = int(os.environ.get('LOCAL_RANK', '0'))
	if world != args.gpus:
		raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run" % (args.gpus, world))
	if not torch.cuda.is_available():
		raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
	torch.cuda.set_device(local)
	dev = torch.device('cuda', local)
	multi = world > 1
	n_host_threads = host_threads()                 # before the binding below narrows the affinity mask
	affinity0 = os.sched_getaffinity(0) if hasattr(os, 'sched_getaffinity') else None
	bound = bind_to_gpu_numa_node(local)
	dist = None
	if multi:
		import torch.distributed as dist
		if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
			os.environ['NCCL_DEBUG'] = 'WARN'          # keep NCCL's version banner off stdout
		dist.init_process_group('nccl', device_id=dev)

	def barrier():
		if multi:
			dist.barrier()

	def allmax(*vals):
		t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
		if multi:
			dist.all_reduce(t, op=dist.ReduceOp.MAX)
		return [float(v) for v in t.tolist()]

	def allsum(*vals):
		t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
		if multi:
			dist.all_reduce(t, op=dist.ReduceOp.SUM)
		return [float(v) for v in t.tolist()]

	# ---- mesh + context ----------------------------------------------------------------------
	if multi:
		s = slab_of_rank(rank, world)
		params = dict(CONSTS, periodic=True, use_pressure_bc=True)
		ctx = capi.BPContext(s.coords, s.cells, 2 * s.n_nodes, s.facets, params, vertex_to_node=s.vertex_to_node,
		                     partition=s.partition, device=local)
		for f, k in ((capi.FIELD_S, 'S'), (capi.FIELD_A, 'A'), (capi.FIELD_BETA, 'beta')):
			ctx.set_field(f, s.fields[k])
	else:
		model = build_model(1)                       # the host mirror of the reference API
		s = slab_from_model(model)
		ctx = capi.context_from_model(model, device=local)
	# a non-default torch stream: the library runs on torch's current stream (so torch.cuda.Event sees its
	# kernels) and that stream can be captured into CUDA graphs (the legacy default stream cannot)
	torch.cuda.set_stream(torch.cuda.Stream(dev))
	ctx.use_torch_stream()
	clocks = ClockSampler(local) if rank == 0 else None      # started early: nvidia-smi needs a moment before its first sample
	if multi:
		ids = [capi.comm_unique_id() if rank == 0 else None]
		dist.broadcast_object_list(ids, src=0, device=dev)
		ctx.comm_init(ids[0], rank, world)
	nt_local = s.cells.shape[0]
	nxy = grid(world)
	nt_global = 6 * nxy * nxy * NZ
	nd = 2 * s.n_nodes
	u_host = synthetic_state(s)
	ctx.assemble(u_host, want_F=True, want_J=True, J=False)          # untimed: first halo exchange sets up the NCCL channels

	# ---- Newton solve (reference defaults: rtol 1e-9, relax 1.0, Krylov rtol 1e-5), cold start --
	newton, u_newton = None, None
	if not args.no_newton:
		pck = {'preconditioner': args.pc} if args.pc else {}
		# one-off set-up (multigrid aggregation on the host, graph capture, NCCL channels) is not part of a solve
		ctx.set_params(relative_tolerance=1e-9, maximum_iterations=1, relaxation_parameter=1.0, krylov_rtol=1e-5, krylov_maxit=8,
		               krylov_error_on_nonconvergence=False, **pck)
		ctx.newton_solve(torch.full((nd,), 3e-16, dtype=torch.float64, device=dev))
		ctx.set_params(relative_tolerance=1e-9, maximum_iterations=25, relaxation_parameter=1.0, krylov_rtol=1e-5, krylov_maxit=10000,
		               krylov_error_on_nonconvergence=True, **pck)
		un = torch.full((nd,), 3e-16, dtype=torch.float64, device=dev)
		barrier(); torch.cuda.synchronize()
		t0 = time.perf_counter()
		st = ctx.newton_solve(un)
		torch.cuda.synchronize(); barrier()
		t_dev, = allmax(time.perf_counter() - t0)
		u_newton = un.cpu().numpy()
		newton = {"solve_s": t_dev, "converged": st['converged'], "newton_iterations": st['iterations'],
		          "krylov_iterations": st['krylov_iterations'], "assemble_ms": st['t_assemble_ms'], "krylov_ms": st['t_krylov_ms'],
		          "preconditioner": ctx.pc_label(), "rtol": 1e-9, "krylov_rtol": 1e-5,
		          "gpu_launches": st['gpu_launches'], "host_launches": st.get('host_launches'),
		          "host_launches_per_krylov_iteration": (st['host_launches'] / max(st['krylov_iterations'], 1)) if st.get('host_launches') is not None else None,
		          "mixed_precision": "the multigrid smoother reads an FP32 copy of the level matrices; CG, its SpMV, the residual and all vectors are FP64"}
		if not multi:
			uh = np.full(nd, 3e-16)
			t0 = time.perf_counter()
			ctx.newton_solve(uh)
			newton["e2e_s"] = time.perf_counter() - t0

	u_dev = torch.from_numpy(u_host).to(dev)
	F_dev = torch.empty_like(u_dev)
	u_pin = torch.from_numpy(u_host).pin_memory()
	F_pin = torch.empty(u_host.size, dtype=torch.float64).pin_memory()

	def step_dev():
		ctx.assemble(u_dev, want_F=True, want_J=True, F=F_dev, J=False)

	def step_e2e():
		ctx.assemble(u_pin.numpy(), want_F=True, want_J=True, F=F_pin.numpy(), J=False)

	def timed(step, steps):
		for _ in range(args.warmup):
			step()
		barrier(); torch.cuda.synchronize()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		t0 = time.perf_counter()
		e0.record()
		for _ in range(steps):
			step()
		e1.record()
		torch.cuda.synchronize()
		wall = time.perf_counter() - t0
		barrier()
		a, b = allmax(e0.elapsed_time(e1) * 1e-3, wall)
		return a / steps, b / steps

	if clocks: clocks.start()
	l0 = ctx.launch_count()
	s_dev, _ = timed(step_dev, args.steps)
	launches, = allsum(float((ctx.launch_count() - l0) * args.steps // (args.steps + args.warmup)))
	# roofline: the cell kernel alone, CUDA events on the stream it is launched on, L2 flushed
	what = capi.ASSEMBLE_F | capi.ASSEMBLE_J
	reps = max(min(args.steps, 50), 10)
	t_cell = ctx.time_assemble(what | 4, repeats=reps, flush_l2=True)
	t_pass = ctx.time_assemble(what, repeats=reps, flush_l2=True)
	step_dev()                                                    # make the resident J valid again
	t_spmv = ctx.time_spmv(repeats=50, flush_l2=True)
	_, s_e2e = timed(step_e2e, args.steps)
	if clocks: clocks.stop()
	e2e_J = None
	if not multi and not os.environ.get('BENCH_NO_E2E_J'):
		J_pin = torch.empty(ctx.nnz_csr(), dtype=torch.float64).pin_memory()
		stepJ = lambda: ctx.assemble(u_pin.numpy(), want_F=True, want_J=True, F=F_pin.numpy(), J=J_pin.numpy())
		stepJ()
		k = max(2, min(args.steps, 5))
		t0 = time.perf_counter()
		for _ in range(k):
			stepJ()
		e2e_J = {"value": nt_global / ((time.perf_counter() - t0) / k) / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * u_host.size,
		         "d2h_bytes_per_step": 8 * u_host.size + 8 * ctx.nnz_csr(), "note": "F and the CSR values of J exported to pinned host memory every step"}
		del J_pin

	# ---- parity at full size (untimed): the oracle's C twin on this rank's arrays ----------------
	parity, cpu_base = None, None
	if not args.no_parity:
		if affinity0 is not None:
			os.sched_setaffinity(0, affinity0)                            # the CPU checker / baseline gets every core back
		co = oracle_for_slab(s, max(1, n_host_threads // world))
		Fo, Jo, t_cpu = co.assemble(u_host)
		own = slice(0, 2 * s.n_owned)
		Fp = ctx.assemble(u_host, want_F=True, want_J=True, J=False)[0]
		step_dev()
		eF = max(relerr(Fp[own], Fo[own]), relerr(F_pin.numpy()[own], Fo[own]), relerr(F_dev.cpu().numpy()[own], Fo[own]))   # ... and the F of the timed calls: pinned host buffer (e2e) and device buffer (value)
		x = np.sin(0.61 * np.repeat(s.node_global, 2) + np.tile([0.3, 1.1], s.n_nodes))      # a function of the global node id
		import scipy.sparse as sp
		Jx_o = sp.csr_matrix((Jo, co.indices, co.indptr), shape=(nd, nd)) @ x
		Jx_p = ctx.spmv(x)
		eJx = relerr(Jx_p[own], Jx_o[own])
		par = {"F_rel": eF, "Jx_rel": eJx}
		if not multi:
			ip, ix = ctx.csr_pattern()
			par["csr_pattern_identical"] = bool(np.array_equal(ip, co.indptr) and np.array_equal(ix, co.indices))
			_, Jp = ctx.assemble(u_host, want_F=False, want_J=True)
			par["J_rel"] = relerr(Jp, Jo) if par["csr_pattern_identical"] else None
			par["J_rel_entrywise"] = relerr_entrywise(Jp, Jo) if par["csr_pattern_identical"] else None
			del Jp
			cpu_base = {"value": nt_local / t_cpu / 1e6, "unit": UNIT, "cores": co.threads, "kind": "port",
			            "sample": "all %d cells of the same mesh, F + J once (%d-point quadrature loop, CSR insertion by binary search, %d OpenMP threads)"
			                      % (nt_local, len(co.q[1]), co.threads)}
		if u_newton is not None:
			# the oracle's own residual at the product's converged velocity, against its residual at the cold start
			r1 = co.assemble(u_newton, want_J=False)[0][own]
			r0 = co.assemble(np.full(nd, 3e-16), want_J=False)[0][own]
			a, b = allsum(float(r1 @ r1), float(r0 @ r0))
			par["newton_oracle_residual_rel"] = float(np.sqrt(a / b))
		vals = allmax(*[float(v) if v is not None else -1.0 for v in (par["F_rel"], par["Jx_rel"])])
		par["F_rel"], par["Jx_rel"] = vals
		par["checked"] = "every rank: F and J x on its owned rows vs oracle/bp_oracle_c.c on the rank's arrays (max over ranks)" if multi else \
			"F, CSR pattern and every J entry vs oracle/bp_oracle_c.c on the whole mesh"
		par["tolerance"] = 1e-12
		par["ok"] = bool(par["F_rel"] < 1e-12 and par["Jx_rel"] < 1e-12 and (multi or (par["csr_pattern_identical"] and par["J_rel"] < 1e-12))
		                 and par.get("newton_oracle_residual_rel", 0.0) < 1.01e-9)
		parity = par
		del co, Fo, Jo

	if rank == 0:
		peaks = {}
		try:
			peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
		except (OSError, ValueError):
			pass
		peak = float(peaks.get('hbm_gbs', 6650.0))
		nv, no = s.coords.shape[0], s.n_owned
		# every input read once, every output written once (this rank's share)
		b_cell = 16 * nt_local + 32 * nv + 24 * s.n_nodes + 8 * ctx.nnz + 16 * no
		b_full = b_cell + 8 * (ctx.n_basal // 2 + 1) + 16 * ctx.n_basal
		b_spmv = 9 * ctx.nnz + 40 * no
		kname = ctx.assembly_kernel_name() if hasattr(ctx, 'assembly_kernel_name') else "k_rows<F,J>"
		traffic, tsrc = measured_traffic(kname, nt_local)
		if kname == "k_tile":
			b_cell = b_full          # the tiled kernel also integrates the basal facets (folded into its first layer): the whole B_asm of SURVEY.md §8d
		line = {"metric": METRIC, "value": nt_global / s_dev / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
		        "ms_per_step": s_dev * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
		        "data": "synthetic",
		        "config": {"workload": workload_name(world), "n_tets": nt_global, "tets_per_rank_incl_halo": nt_local, "n_dofs_rank0": 2 * no, "nnz_rank0": ctx.nnz,
		                   "partition": "x-slabs of ice columns, 1 ghost node layer, halo of u by grouped ncclSend/ncclRecv before every assembly" if multi else "one GPU",
		                   "l2": "per-rank working set (J values %d MB) larger than L2; roofline launches flush L2" % (8 * ctx.nnz >> 20),
		                   "state": "synthetic sheared sliding flow, a function of the global node id",
		                   "host_binding": "rank bound to its GPU's NUMA node (NVML ideal affinity)" if bound else "none"},
		        "e2e": {"value": nt_global / s_e2e / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * u_host.size * world, "d2h_bytes_per_step": 8 * u_host.size * world},
		        "gpu_launches": int(launches),
		        "roofline": {"kernel": kname + (" (rank 0)" if multi else ""), "bound": "hbm", "achieved": b_cell / (t_cell * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
		                     "frac": b_cell / (t_cell * 1e-3) / 1e9 / peak, "traffic": traffic, "traffic_source": tsrc,
		                     "peak_source": "measured" if peaks else "fallback", "bytes_per_tet": b_cell / nt_local, "ms": t_cell,
		                     "pass_ms": t_pass, "pass_frac": b_full / (t_pass * 1e-3) / 1e9 / peak},
		        "roofline_spmv": {"kernel": "k_spmv" + (" (rank 0)" if multi else ""), "bound": "hbm", "achieved": b_spmv / (t_spmv * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
		                          "frac": b_spmv / (t_spmv * 1e-3) / 1e9 / peak, "ms": t_spmv},
		        "clocks": clocks.summary()}
		if e2e_J:
			line["e2e_with_J"] = e2e_J
		if newton:
			line["newton"] = newton
		if parity:
			line["parity"] = parity
		if cpu_base and not args.no_cpu:
			line["cpu_baseline"] = cpu_base
		if not multi and not args.no_cpu and not args.no_newton:
			line["cpu_baseline_newton"] = cpu_newton_baseline(n_host_threads, nxy=int(os.environ.get('BENCH_NEWTON_NXY', '130')))
		emit(json.dumps(line))
	if clocks: clocks.close()
	ctx.close()
	if multi:
		dist.barrier()
		dist.destroy_process_group()


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('--gpus', type=int, default=1)
	ap.add_argument('--steps', type=int, default=200)
	ap.add_argument('--warmup', type=int, default=3)
	ap.add_argument('--impl', default='ours')
	ap.add_argument('--no-newton', action='store_true')
	ap.add_argument('--no-cpu', action='store_true')
	ap.add_argument('--no-parity', action='store_true')
	ap.add_argument('--pc', default='', help='preconditioner of the Newton solve (default: the library default)')
	args = ap.parse_args()
	quiet_stdout()
	args.warmup = max(args.warmup, 3) if args.impl != 'reference' else args.warmup
	if args.impl == 'reference':
		return run_reference(args)
	return run_ours(args)


if __name__ == '__main__':
	main()
