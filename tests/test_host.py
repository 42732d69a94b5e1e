"""Host logic without a GPU: the D3Model mirror against the oracle's independent mesh
code, the reference's parameter dictionary, and the C-ABI library (loads, exports every
symbol of include/cslvr_b200.h, refuses to compute without a device)."""
import ctypes
import os
import re

import numpy as np
import pytest

from cases import ismip_model, marine_model, oracle_problem
from oracle import bp_oracle as bo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mirror_mesh_equals_oracle_mesh():
	m = ismip_model('A')
	P = bo.ismip_hom('A')
	assert np.array_equal(P.cells, m.mesh.cells()) and np.array_equal(P.coords, m.mesh.coordinates())
	assert np.array_equal(P.cell_dofs, m.cell_dofs)
	assert np.array_equal(P.fc, m.facet_cell) and np.array_equal(P.fl, m.facet_local) and np.array_equal(P.ff, m.ff)
	assert np.array_equal(P.beta, m.vertex_field(m.beta)) and np.array_equal(P.S, m.vertex_field(m.S))


def test_markers_follow_d3model_rules():
	m = marine_model()
	# all six marker classes of cslvr/d3model.py:410-455 that the defaults can produce
	assert m.N_GAMMA_U_GND > 0 and m.N_GAMMA_U_FLT > 0 and m.N_GAMMA_B_GND > 0 and m.N_GAMMA_B_FLT > 0
	assert m.N_GAMMA_L_OVR > 0 and m.N_GAMMA_L_UDR > 0 and m.N_OMEGA_FLT > 0
	nx, ny, nz = m.mesh.shape
	assert m.N_GAMMA_B_GND + m.N_GAMMA_B_FLT == 2 * nx * ny == m.N_GAMMA_U_GND + m.N_GAMMA_U_FLT
	assert m.N_GAMMA_L_OVR + m.N_GAMMA_L_UDR == 2 * nz * 2 * (nx + ny)
	# deform_mesh_to_geometry re-marks with the default mask: everything grounded (d3model.py:667)
	m.deform_mesh_to_geometry(m.S.values.copy(), m.B.values.copy())
	assert m.N_GAMMA_B_FLT == 0


def test_periodic_map_and_dof_counts():
	m = ismip_model('A')
	assert m.n_nodes == 15 * 15 * 6 and m.QM2.dim() == 2700
	X = m.flat_coordinates
	img = m.node_vertex[m.vertex_to_node]
	d = np.abs(X - X[img])
	assert np.all((d[:, 0] < 1e-9) | (np.abs(d[:, 0] - 80000.0) < 1e-6)) and np.all(d[:, 2] < 1e-12)


def test_default_solve_params_match_reference():
	"""cslvr/momentumbp.py:165-203."""
	from cslvr_b200.momentumbp import MomentumBPBase

	class Probe(MomentumBPBase):
		def __init__(self, model):
			self.model, self.linear = model, False
	for per, relax in ((True, 1.0), (False, 0.7)):
		p = Probe(ismip_model('A', n=(3, 3, 2), periodic=per)).default_solve_params()
		ns = p['solver']['newton_solver']
		assert ns['relative_tolerance'] == 1e-9 and ns['maximum_iterations'] == 25
		assert ns['relaxation_parameter'] == relax and ns['error_on_nonconvergence'] is False
		assert ns['linear_solver'] == 'cg' and ns['preconditioner'] == 'hypre_amg'
		assert p['solve_vert_velocity'] and p['solve_pressure'] and p['vert_solve_method'] == 'mumps'


def test_reduced_stokes_defaults_match_reference():
	"""cslvr/momentumstokes.py:420-447 (relaxation 0.8, 50 iterations) and :396-404 (get_velocity)."""
	from cslvr_b200 import MomentumDukowiczStokesReduced, MomentumBPBase
	m = ismip_model('A', n=(3, 3, 2))
	mom = MomentumDukowiczStokesReduced(m)               # construction needs no device
	assert isinstance(mom, MomentumBPBase) and mom._reduced_stokes and mom.get_residual() == 'rs-dukowicz'
	ns = mom.solve_params['solver']['newton_solver']
	assert ns['relaxation_parameter'] == 0.8 and ns['maximum_iterations'] == 50 and ns['relative_tolerance'] == 1e-9
	assert ns['linear_solver'] == 'cg' and ns['preconditioner'] == 'hypre_amg' and ns['error_on_nonconvergence'] is False
	assert mom.solve_params['solve_vert_velocity'] and mom.solve_params['solve_pressure']
	ux, uy, uz = mom.get_velocity()
	assert ux.size == uy.size == uz.size == m.n_nodes and uz is mom.u_z.values
	with pytest.raises(NotImplementedError):             # an external solver cannot tape
		mom.solve(annotate=True)


def test_misuse_errors_like_the_reference():
	from cslvr_b200 import MomentumDukowiczBP, D3Model, BoxMesh, Point
	with pytest.raises(SystemExit):                      # cslvr/momentumbp.py:30-33
		MomentumDukowiczBP(object())
	with pytest.raises(NotImplementedError):             # P1 only
		D3Model(BoxMesh(Point(0, 0, 0), Point(1, 1, 1), 2, 2, 2), order=2)
	m = ismip_model('A', n=(3, 3, 2))
	with pytest.raises(SystemExit):                      # cslvr/momentum.py:49-53
		MomentumDukowiczBP(m, solve_params=3)
	with pytest.raises(TypeError):                       # use_lat_bcs is broken upstream too
		MomentumDukowiczBP(m, use_lat_bcs=True)


def test_library_exports_every_declared_symbol(lib_built):
	hdr = open(os.path.join(ROOT, 'include', 'cslvr_b200.h')).read()
	declared = set(re.findall(r'\b(bp_[a-z_0-9]+)\s*\(', hdr))
	from cslvr_b200 import capi
	assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
	lib = ctypes.CDLL(lib_built)
	for s in sorted(declared):
		assert hasattr(lib, s), s


def test_no_cpu_fallback(lib_built):
	import torch
	if torch.cuda.is_available():
		pytest.skip("a GPU is visible")
	from cases import product_context
	from cslvr_b200.capi import BPError
	with pytest.raises(BPError) as e:
		product_context(ismip_model('A', n=(3, 3, 2)))
	assert 'no CPU path' in str(e.value)


def test_product_does_not_import_the_oracle():
	import glob
	for f in glob.glob(os.path.join(ROOT, 'cslvr_b200', '**', '*'), recursive=True):
		if f.endswith(('.py', '.cu', '.h')):
			src = open(f).read()
			assert 'import oracle' not in src and 'from oracle' not in src and 'bp_oracle' not in src, f


def test_bench_reference_arm_prints_one_json_line():
	"""bench.py --impl reference (the reference's CPU path, oracle port): exactly one JSON line on stdout
	with the keys of the bench contract; a small mesh through BENCH_NXY keeps it to seconds."""
	import json
	import subprocess
	import sys
	env = dict(os.environ, BENCH_NXY='20', BENCH_NEWTON_NXY='15', CSLVR_B200_QUIET='1', OMP_NUM_THREADS='1')   # as under torch.distributed.run
	r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
	                   capture_output=True, text=True, env=env, timeout=600)
	assert r.returncode == 0, r.stderr[-2000:]
	lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
	assert len(lines) == 1, r.stdout
	d = json.loads(lines[0])
	for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
	          'vs_baseline', 'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
		assert k in d, k
	assert d['impl'] == 'reference' and d['unit'] == 'Mtets/s' and d['value'] > 0 and d['vs_baseline'] is None
	assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and 'workload' in d['config']
	assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
	assert 'ISMIP-HOM A' in d['config']['workload']
	# all host cores even though the launcher exported OMP_NUM_THREADS=1
	assert d['cpu_baseline']['cores'] == len(os.sched_getaffinity(0))
	nb = d['cpu_baseline_newton']
	assert nb[0]['converged'] and nb[0]['newton_iterations'] > 3 and nb[0]['solve_s'] > 0


@pytest.mark.parametrize('case', ['ismip', 'marine', 'delaunay', 'eismint', 'greenland'])
def test_facet_markers_of_the_mirror_equal_the_oracles(case):
	"""a13: ``D3Model.calculate_boundaries`` (cslvr/d3model.py:337-475) of the host mirror against the oracle's own
	exterior-facet search and marker rules on every mesh family of the test suite, shelf masks included -- the
	oracle is NOT handed the product's markers here."""
	from cases import delaunay_extruded_model, eismint_dome_model, greenland_like_model
	m = {'ismip': lambda: ismip_model('A', n=(6, 5, 3)), 'marine': marine_model, 'delaunay': delaunay_extruded_model,
	     'eismint': eismint_dome_model, 'greenland': greenland_like_model}[case]()
	X, cells = m.mesh.coordinates(), m.mesh.cells()
	fc, fl = bo.exterior_facets(cells)
	ff = bo.mark_facets(X, cells, fc, fl, mask=getattr(m, 'mask_callable', None))
	mine = {(int(c), int(l)): int(k) for c, l, k in zip(m.facet_cell, m.facet_local, m.ff)}
	theirs = {(int(c), int(l)): int(k) for c, l, k in zip(fc, fl, ff)}
	assert mine == theirs
	if case in ('marine', 'delaunay'):
		assert 5 in mine.values() and 3 in mine.values() and 10 in mine.values()       # floating bed, grounded bed, under-water sides


def test_basin_divide_markers():
	"""mark_divide=True with the ice-sheet contour (cslvr/d3model.py:352-364, 441-448): lateral facets further than 1 km
	from the contour become GAMMA_L_DVD = 7; without the contour the reference exits."""
	m = marine_model()
	X, cells = m.mesh.coordinates(), m.mesh.cells()
	L = X[:, 0].max()
	# the "true" ice-sheet margin: only the x = L side; the other three sides are cut through the ice (divides)
	contour = np.column_stack([np.full(200, L), np.linspace(X[:, 1].min(), X[:, 1].max(), 200)])
	m.calculate_boundaries(mask=m.mask_callable, mark_divide=True, contour=contour)
	assert m.N_GAMMA_L_DVD > 0 and m.N_GAMMA_L_OVR + m.N_GAMMA_L_UDR > 0
	fc, fl = bo.exterior_facets(cells)
	ff = bo.mark_facets(X, cells, fc, fl, mask=m.mask_callable, contour=contour)
	assert {(int(c), int(l)): int(k) for c, l, k in zip(m.facet_cell, m.facet_local, m.ff)} == \
	       {(int(c), int(l)): int(k) for c, l, k in zip(fc, fl, ff)}
	with pytest.raises(SystemExit):
		m.calculate_boundaries(mark_divide=True)


def test_lateral_dirichlet_conditions_of_momentum_bp():
	"""cslvr/momentumbp.py:101-112: with mark_divide and use_lat_bcs MomentumBP constrains (u_x, u_y) on the nodes of
	the GAMMA_L_DVD facets to (u_x_lat, u_y_lat); MomentumDukowiczBP raises there (momentumbp.py:495), as upstream."""
	from cslvr_b200 import MomentumBP, MomentumDukowiczBP
	m = marine_model()
	X = m.mesh.coordinates()
	L = X[:, 0].max()
	contour = np.column_stack([np.full(200, L), np.linspace(X[:, 1].min(), X[:, 1].max(), 200)])
	m.calculate_boundaries(mask=m.mask_callable, mark_divide=True, contour=contour)
	nodes = m.divide_nodes()
	# the divide is every lateral side but x = L: its nodes are the mesh vertices on x = 0, y = 0 or y = max (all layers)
	xf = m.flat_coordinates
	on = (np.abs(xf[:, 0]) < 1e-9) | (np.abs(xf[:, 1]) < 1e-9) | (np.abs(xf[:, 1] - xf[:, 1].max()) < 1e-9)
	assert np.array_equal(nodes, np.flatnonzero(on))
	m.init_u_lat(lambda x, y, z: 3.0 + 0.0 * x)
	m.init_u_y_lat(-1.5)
	mom = MomentumBP(m, use_lat_bcs=True)
	bcs = mom.get_boundary_conditions()
	assert len(bcs) == 2 and np.array_equal(bcs[0][0], 2 * nodes) and np.array_equal(bcs[1][0], 2 * nodes + 1)
	assert np.allclose(bcs[0][1].values[bcs[0][2]], 3.0) and np.allclose(bcs[1][1].values[bcs[1][2]], -1.5)
	assert MomentumBP(m).get_boundary_conditions() == []              # not without use_lat_bcs
	with pytest.raises(TypeError):
		MomentumDukowiczBP(m, use_lat_bcs=True)
