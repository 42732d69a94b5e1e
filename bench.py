#!/usr/bin/env python
"""bench.py -- BP residual+Jacobian assembly throughput (Mtets/s) and Newton solve time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): "BP residual+Jacobian assembly Mtets/s; Newton solve time".
A *step* is one fused assemble(F) + assemble(J) pass of the hot path over the whole
mesh at a fixed velocity state.  N = 1 runs BASELINE.json configs[1], ISMIP-HOM C
(L = 20 km, basal sliding) on a 260x260x5 BoxMesh (2.03 M tets); N > 1 is weak scaling:
the same 2.03 M tets per GPU, the domain extended in x and split into N slabs of ice
columns (one rank per GPU; NCCL only for the ghost-node exchange of u before assembly
and for the SpMV halo / reductions in the Newton solve).

value       device-resident: u already in HBM, F written to HBM, J stays resident.
e2e         the same pass through the C-ABI with HOST (pinned) buffers: u copied
            host->device and F device->host inside the timed region (J stays resident on
            the device exactly as it does between assemble() and solve() in the reference).
roofline    the dominant kernel (cell assembly) timed alone with CUDA events on the
            library's stream, L2 flushed before every launch; algorithmic bytes per
            SURVEY.md §8d / DESIGN.md.
cpu_baseline / --impl reference
            the oracle's C/OpenMP restatement of the reference's CPU path ("port": the
            reference is Python on an un-vendored FEniCS stack and cannot run here) on
            all host cores, same config, bounded sample of the cells.
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
sys.path.insert(0, os.path.join(ROOT, 'tests'))
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
	fd = _JSON_FD
	if fd is None:                 # imported as a module by bench_multi while the script itself runs as __main__
		fd = getattr(sys.modules.get('__main__'), '_JSON_FD', None)
	if fd is None:
		print(text, flush=True)
	else:
		os.write(fd, (text + "\n").encode())


METRIC = "BP residual+Jacobian assembly throughput"
UNIT = "Mtets/s"
NX, NY, NZ, L = 260, 260, 5, 20000.0
if os.environ.get('BENCH_NXY'):                 # tests only: a small mesh for the JSON-contract check
	NX = NY = int(os.environ['BENCH_NXY'])


def build_model(n_slabs=1):
	"""ISMIP-HOM C (simulations/ismip_hom/ismip_hom_c.py:4-25), periodic, nx = 260 per slab."""
	from cslvr_b200 import BoxMesh, Point, D3Model
	Lx = L * n_slabs
	mesh = BoxMesh(Point(0, 0, 0), Point(Lx, L, 1), NX * n_slabs, NY, NZ)
	model = D3Model(mesh, use_periodic=True)
	model.calculate_boundaries()
	a = 0.1 * np.pi / 180
	model.deform_mesh_to_geometry(lambda x, y, z: -x * np.tan(a), lambda x, y, z: -x * np.tan(a) - 1000.0)
	model.init_beta(lambda x, y, z: 1000.0 + 1000.0 * np.sin(2 * np.pi * x / L) * np.sin(2 * np.pi * y / L))
	model.init_A(1e-16)
	return model


def workload_name(n):
	return "ISMIP-HOM C (L=20 km, basal sliding) MomentumDukowiczBP, periodic BoxMesh %dx%dx%d = %.2f M tets%s" % (
		NX * n, NY, NZ, 6e-6 * NX * n * NY * NZ, "" if n == 1 else " in %d x-slabs" % n)


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


def algorithmic_bytes(model, ctx):
	"""every input read once, every output written once."""
	nt, nv, nn, nd = model.num_cells, model.num_vertices, model.n_nodes, 2 * model.n_nodes
	nnz = ctx.nnz
	cell = 16 * nt + 32 * nv + 24 * nn + 8 * nnz + 8 * nd
	nb = ctx.n_basal
	full = cell + 8 * (nb // 2 + 1) + 16 * nb
	spmv_bsr = 9 * nnz + 20 * nd
	return cell, full, spmv_bsr


def cpu_reference(model, u, sample_cells=None, steps=1, warmup=0):
	"""time the oracle's C/OpenMP restatement on the host cores."""
	from cases import oracle_problem
	from oracle.bp_oracle_cwrap import COracle
	P = oracle_problem(model)
	co = COracle(P)
	nt = P.cells.shape[0]
	ns = nt if sample_cells is None else min(nt, int(sample_cells))
	for _ in range(warmup):
		co.assemble(u, n_cells=ns)
	ts = []
	for _ in range(steps):
		_, _, t = co.assemble(u, n_cells=ns)
		ts.append(t)
	t = float(np.mean(ts))
	return {"value": ns / t / 1e6, "unit": UNIT, "cores": co.threads, "kind": "port",
	        "sample": "first %d of %d cells of the same mesh, F+J, %d-point quadrature loop, CSR insertion by binary search, %d OpenMP threads, mean of %d"
	                  % (ns, nt, len(co.q[1]), co.threads, steps)}, t


def synthetic_state(model, seed=1234):
	"""a sheared sliding flow of ISMIP-HOM magnitude, perturbed."""
	X = model.mesh.coordinates()[model.node_vertex]
	zrel = (X[:, 2] - model.B.values[model.node_vertex]) / np.maximum(model.S.values[model.node_vertex] - model.B.values[model.node_vertex], 1.0)
	ux = 15.0 + 10.0 * np.sin(2 * np.pi * X[:, 0] / L) * np.sin(2 * np.pi * X[:, 1] / L) + 20.0 * zrel ** 0.5
	uy = 2.0 * np.cos(2 * np.pi * X[:, 0] / L) * zrel
	u = np.empty(2 * model.n_nodes)
	u[0::2], u[1::2] = ux, uy
	return u * (1.0 + 1e-3 * np.random.default_rng(seed).normal(size=u.size))


def run_reference(args):
	rank = int(os.environ.get('RANK', '0'))
	if rank != 0:
		return
	model = build_model(args.gpus)
	u = synthetic_state(model)
	budget = 150.0 / max(args.steps + args.warmup, 1)          # seconds per step
	probe, t = cpu_reference(model, u, sample_cells=200000)
	sample = int(min(model.num_cells, max(100000, probe["value"] * 1e6 * budget)))
	base, t = cpu_reference(model, u, sample_cells=sample, steps=args.steps, warmup=args.warmup)
	line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
	        "warmup": args.warmup, "ms_per_step": t * 1e3 * model.num_cells / sample, "higher_is_better": True,
	        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
	        "config": {"workload": workload_name(args.gpus)},
	        "cpu_baseline": base,
	        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
	        "gpu_launches": 0}
	emit(json.dumps(line))


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('--gpus', type=int, default=1)
	ap.add_argument('--steps', type=int, default=500)
	ap.add_argument('--warmup', type=int, default=3)
	ap.add_argument('--impl', default='ours')
	ap.add_argument('--no-newton', action='store_true')
	ap.add_argument('--no-cpu', action='store_true')
	ap.add_argument('--pc', default='', help='preconditioner of the Newton solve (default: the library default)')
	ap.add_argument('--exp', default='C', help='ISMIP-HOM experiment of the multi-rank arm: C (default, L=20 km per 260 columns) or A (L=80 km total)')
	ap.add_argument('--nx', type=int, default=0, help='multi-rank arm: TOTAL columns in x (default 260 per GPU); strong-scaling studies fix it')
	ap.add_argument('--ny', type=int, default=0)
	args = ap.parse_args()
	quiet_stdout()
	args.warmup = max(args.warmup, 3) if args.impl != 'reference' else args.warmup
	if args.impl == 'reference':
		return run_reference(args)

	import torch
	from cslvr_b200 import capi
	from cases import product_context
	world = int(os.environ.get('WORLD_SIZE', '1'))
	rank = int(os.environ.get('RANK', '0'))
	local = int(os.environ.get('LOCAL_RANK', '0'))
	if world != args.gpus:
		raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run" % (args.gpus, world))
	if not torch.cuda.is_available():
		raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
	torch.cuda.set_device(local)
	dev = torch.device('cuda', local)
	if world > 1:
		from bench_multi import run_multi
		return run_multi(args, rank, world, local)

	model = build_model(1)
	ctx = product_context(model)
	ctx.device = local
	# a non-default torch stream: the library runs on torch's current stream (so torch.cuda.Event sees its
	# kernels) and that stream can be captured into CUDA graphs (the legacy default stream cannot)
	torch.cuda.set_stream(torch.cuda.Stream(dev))
	ctx.use_torch_stream()
	nt = model.num_cells
	u_host = synthetic_state(model)

	# ---- Newton solve (reference defaults: rtol 1e-9, relax 1.0, Krylov rtol 1e-5), cold start
	newton = None
	if not args.no_newton:
		pck = {'preconditioner': args.pc} if args.pc else {}
		# one-off set-up (multigrid aggregation on the host, graph capture) is not part of a solve
		ctx.set_params(relative_tolerance=1e-9, maximum_iterations=1, relaxation_parameter=1.0, krylov_rtol=1e-5, krylov_maxit=4, **pck)
		ctx.newton_solve(torch.full((2 * model.n_nodes,), 3e-16, dtype=torch.float64, device=dev))
		ctx.set_params(relative_tolerance=1e-9, maximum_iterations=25, relaxation_parameter=1.0, krylov_rtol=1e-5, krylov_maxit=10000, **pck)
		un = torch.full((2 * model.n_nodes,), 3e-16, dtype=torch.float64, device=dev)
		torch.cuda.synchronize()
		t0 = time.perf_counter()
		st = ctx.newton_solve(un)
		torch.cuda.synchronize()
		t_dev = time.perf_counter() - t0
		uh = np.full(2 * model.n_nodes, 3e-16)
		t0 = time.perf_counter()
		st2 = ctx.newton_solve(uh)
		t_e2e = time.perf_counter() - t0
		newton = {"solve_s": t_dev, "e2e_s": t_e2e, "converged": st['converged'], "newton_iterations": st['iterations'],
		          "krylov_iterations": st['krylov_iterations'], "assemble_ms": st['t_assemble_ms'], "krylov_ms": st['t_krylov_ms'],
		          "preconditioner": ctx.pc_label(), "rtol": 1e-9, "krylov_rtol": 1e-5,
		          "gpu_launches": st['gpu_launches']}
		if st['converged']:
			u_host = uh * (1.0 + 1e-3 * np.random.default_rng(1234).normal(size=uh.size))

	u_dev = torch.from_numpy(u_host).to(dev)
	F_dev = torch.empty_like(u_dev)
	u_pin = torch.from_numpy(u_host).pin_memory()
	F_pin = torch.empty(u_host.size, dtype=torch.float64).pin_memory()

	def step_dev():
		ctx.assemble(u_dev, want_F=True, want_J=True, F=F_dev, J=False)

	def step_e2e():
		ctx.assemble(u_pin.numpy(), want_F=True, want_J=True, F=F_pin.numpy(), J=False)

	clocks = ClockSampler(local)
	for _ in range(args.warmup):
		step_dev()
	torch.cuda.synchronize()
	l0 = ctx.launch_count()
	clocks.start()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(args.steps):
		step_dev()
	e1.record()
	torch.cuda.synchronize()
	ms = e0.elapsed_time(e1) / args.steps
	launches = ctx.launch_count() - l0
	# roofline: the cell kernel alone, CUDA events on the stream it is launched on, L2 flushed
	what = capi.ASSEMBLE_F | capi.ASSEMBLE_J
	t_cell = ctx.time_assemble(what | 4, repeats=max(args.steps, 10), flush_l2=True)
	t_pass = ctx.time_assemble(what, repeats=max(args.steps, 10), flush_l2=True)
	ctx.assemble(u_dev, want_F=True, want_J=True, F=F_dev, J=False)       # make the resident J valid again
	t_spmv = ctx.time_spmv(repeats=50, flush_l2=True)
	for _ in range(args.warmup):
		step_e2e()
	t0 = time.perf_counter()
	for _ in range(args.steps):
		step_e2e()
	t_e2e = (time.perf_counter() - t0) / args.steps
	clocks.stop()
	clk = clocks.summary()
	clocks.close()

	peaks = {}
	try:
		peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
	except (OSError, ValueError):
		pass
	peak = float(peaks.get('hbm_gbs', 6650.0))
	b_cell, b_full, b_spmv = algorithmic_bytes(model, ctx)
	line = {"metric": METRIC, "value": nt / ms / 1e3, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
	        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
	        "data": "synthetic",
	        "config": {"workload": workload_name(1), "n_tets": nt, "n_dofs": 2 * model.n_nodes, "nnz": ctx.nnz,
	                   "l2": "working set (J values %d MB + slots %d MB) larger than L2; roofline launches flush L2" % (8 * ctx.nnz >> 20, 64 * nt >> 20),
	                   "state": "converged Newton solution x (1 + 1e-3 N(0,1))" if newton and newton['converged'] else "synthetic sheared flow"},
	        "e2e": {"value": nt / t_e2e / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * u_host.size, "d2h_bytes_per_step": 8 * u_host.size},
	        "gpu_launches": int(launches),
	        "roofline": {"kernel": "k_rows<F,J> (cell assembly, row gather)", "bound": "hbm", "achieved": b_cell / (t_cell * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
	                     "frac": b_cell / (t_cell * 1e-3) / 1e9 / peak,
	                     # dram__bytes_read.sum + dram__bytes_write.sum of one k_rows<1,1,0> launch on this very workload,
	                     # ncu --set full, profiles/r1_h_ncu_k_rows.csv (254.5 + 147.1 MB); the excess over the
	                     # algorithmic bytes is the incidence list (24 B per (tet, node) pair) and the mirror slots
	                     "traffic": 401.6e6 if nt == 2028000 else None,
	                     "peak_source": "measured" if peaks else "fallback", "bytes_per_tet": b_cell / nt, "ms": t_cell,
	                     "pass_ms": t_pass, "pass_frac": b_full / (t_pass * 1e-3) / 1e9 / peak},
	        "roofline_spmv": {"kernel": "k_spmv", "bound": "hbm", "achieved": b_spmv / (t_spmv * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
	                          "frac": b_spmv / (t_spmv * 1e-3) / 1e9 / peak, "ms": t_spmv},
	        "clocks": clk}
	if newton:
		line["newton"] = newton
	if not args.no_cpu:
		base, _ = cpu_reference(model, u_host, sample_cells=None if nt <= 2500000 else 2000000, steps=1, warmup=0)
		line["cpu_baseline"] = base
	emit(json.dumps(line))


if __name__ == '__main__':
	main()
