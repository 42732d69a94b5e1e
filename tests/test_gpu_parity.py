"""GPU parity: the CUDA path through the C-ABI vs the oracle on the same inputs.

Tolerances are BASELINE.json's: CSR sparsity / dof map bit-exact, assembled
residual and Jacobian to a relative 1e-12, converged velocity to a relative 1e-8.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from cases import (ismip_model, marine_model, oracle_problem, product_context, state, relerr)

pytestmark = pytest.mark.gpu

TOL_ASM = 1e-12
TOL_U = 1e-8


def _check_assembly(model, **kw):
	P = oracle_problem(model, **{k: v for k, v in kw.items() if k in ('use_pressure_bc', 'quad_degree')})
	pk = {}
	if 'quad_degree' in kw and kw['quad_degree'] is not None:
		pk['quadrature_degree'] = kw['quad_degree']
	oip, oix = P.pattern()
	for mode in ('gather', 'scatter'):
		c = product_context(model, use_pressure_bc=kw.get('use_pressure_bc', True), assembly=mode, **pk)
		ip, ix = c.csr_pattern()
		assert np.array_equal(ip, oip) and np.array_equal(ix, oix)       # bit-exact sparsity
		for seed, scale in ((1, 30.0), (2, 1e-3), (3, 0.0)):
			u = state(P.n_dofs, seed, scale) + 3e-16
			F, J = c.assemble(u, want_F=True, want_J=True)
			Fo, Jo = P.residual(u), P.jacobian(u)
			assert relerr(F, Fo) < TOL_ASM, (mode, seed, relerr(F, Fo))
			assert relerr(J, Jo) < TOL_ASM, (mode, seed, relerr(J, Jo))
			# F-only and J-only launches give the same numbers as the fused one
			F2, _ = c.assemble(u, want_F=True, want_J=False)
			_, J2 = c.assemble(u, want_F=False, want_J=True)
			assert relerr(F2, F) < 1e-14 and relerr(J2, J) < 1e-14
			if mode == 'gather':           # no atomics in the cell part: reproducible run to run
				F3, J3 = c.assemble(u, want_F=True, want_J=True)   # (facet terms still add atomically)
				touched = np.zeros(P.n_dofs, dtype=bool)
				touched[P.cell_dofs[P.fc].ravel()] = True
				assert np.array_equal(F3[~touched], F[~touched]) and relerr(F3, F) < 1e-15 and relerr(J3, J) < 1e-15
		c.close()
	return P


def test_assembly_ismip_a_periodic():
	_check_assembly(ismip_model('A'))


def test_assembly_ismip_c_periodic():
	_check_assembly(ismip_model('C', L=20000.0, n=(12, 9, 4)))


def test_assembly_nonperiodic_marine_variable_A():
	m = marine_model()
	assert m.N_GAMMA_B_FLT > 0 and m.N_GAMMA_L_UDR > 0 and m.N_GAMMA_L_OVR > 0
	_check_assembly(m)
	_check_assembly(m, use_pressure_bc=False)


@pytest.mark.parametrize('deg', [1, 2, 3, 4, 6, 9])
def test_assembly_quadrature_degrees(deg):
	_check_assembly(marine_model(n=(5, 4, 3)), quad_degree=deg)


def test_assembly_glen_exponent_generic():
	m = marine_model(n=(5, 4, 3))
	m.n = 2.5
	_check_assembly(m)


def test_arbitrary_cell_dofs_numbering():
	"""a dolfin-like reordered dof map: sparsity must follow the caller's numbering."""
	m = ismip_model('A', n=(6, 5, 3))
	nd = 2 * m.n_nodes
	perm = np.random.default_rng(7).permutation(nd).astype(np.int32)
	cd = perm[m.cell_dofs]
	P = oracle_problem(m, cell_dofs=cd, n_dofs=nd)
	c = product_context(m, cell_dofs=cd)
	ip, ix = c.csr_pattern()
	oip, oix = P.pattern()
	assert np.array_equal(ip, oip) and np.array_equal(ix, oix)
	u = state(nd, 5)
	F, J = c.assemble(u, want_F=True, want_J=True)
	assert relerr(F, P.residual(u)) < TOL_ASM and relerr(J, P.jacobian(u)) < TOL_ASM
	c.close()


def test_spmv_and_krylov_against_scipy():
	m = marine_model()
	c = product_context(m, krylov_rtol=1e-13)
	u = state(2 * m.n_nodes, 11)
	_, J = c.assemble(u, want_F=True, want_J=True)
	ip, ix = c.csr_pattern()
	A = sp.csr_matrix((J, ix, ip), shape=(2 * m.n_nodes,) * 2)
	x = state(2 * m.n_nodes, 12, 1.0)
	assert relerr(c.spmv(x), A @ x) < 1e-13
	b = state(2 * m.n_nodes, 13, 1e6)
	dx, its, rr = c.krylov_solve(b)
	import scipy.sparse.linalg as spla
	ref = spla.spsolve(A.tocsc(), b)
	assert relerr(dx, ref) < 1e-8 and its > 0
	c.close()


@pytest.mark.parametrize('exp,L,n', [('A', 80000.0, (15, 15, 5)), ('C', 20000.0, (10, 10, 4))])
def test_newton_velocity_matches_oracle(exp, L, n):
	"""tightened tolerances on both sides (SURVEY §7 hard parts): Newton rtol 1e-11
	(two orders above the rounding floor of |F|/|F0|), Krylov rtol 1e-13 vs sparse LU,
	relax 1.0."""
	m = ismip_model(exp, L=L, n=n)
	P = oracle_problem(m)
	uo, io = P.newton(rtol=1e-11, maxit=40)
	assert io['converged']
	c = product_context(m, relative_tolerance=1e-11, maximum_iterations=40, krylov_rtol=1e-13)
	u = np.full(P.n_dofs, 3e-16)
	st = c.newton_solve(u)
	assert st['converged']
	assert relerr(u, uo) < TOL_U, relerr(u, uo)
	assert abs(st['iterations'] - io['iterations']) <= 1
	c.close()


def test_newton_default_tolerances_and_marine():
	"""reference defaults: rtol 1e-9, Krylov rtol 1e-5, relax 0.7 (non-periodic) -- same convergence history as the oracle;
	and, with both sides converged tightly (Newton 1e-12, Krylov 1e-13), the velocity to north_star's 1e-8."""
	m = marine_model()
	P = oracle_problem(m)
	uo, io = P.newton(rtol=1e-9, relax=0.7, maxit=60)
	c = product_context(m, relative_tolerance=1e-9, relaxation_parameter=0.7, maximum_iterations=60)
	u = np.full(P.n_dofs, 3e-16)
	st = c.newton_solve(u)
	assert st['converged'] == io['converged'] and abs(st['iterations'] - io['iterations']) <= 1
	c.close()
	uo, io = P.newton(rtol=1e-12, relax=0.7, maxit=100)
	c = product_context(m, relative_tolerance=1e-12, relaxation_parameter=0.7, maximum_iterations=100, krylov_rtol=1e-13)
	u = np.full(P.n_dofs, 3e-16)
	st = c.newton_solve(u)
	assert st['converged'] and io['converged']
	assert relerr(u, uo) < TOL_U, relerr(u, uo)
	c.close()


def test_reference_script_runs_unchanged():
	"""docs/get_started.rst:13-50 line for line (Expressions -> callables)."""
	from cslvr_b200 import BoxMesh, Point, D3Model, MomentumDukowiczBP, MomentumBP
	L = 8000
	mesh = BoxMesh(Point(0.0, 0.0, 0.0), Point(L, L, 1), 15, 15, 5)
	model = D3Model(mesh, out_dir='./results/', use_periodic=True)
	model.calculate_boundaries()
	a = 0.5 * np.pi / 180
	surface = lambda x, y, z: -x * np.tan(a)
	bed = lambda x, y, z: -x * np.tan(a) - 1000.0 + 500.0 * np.sin(2 * np.pi * x / L) * np.sin(2 * np.pi * y / L)
	model.deform_mesh_to_geometry(surface, bed)
	model.init_beta(1000)
	model.init_A(1e-16)
	mom = MomentumDukowiczBP(model)
	mom.solve()
	assert mom.stats['converged']
	ux = model.u.sub(0)
	top = model.mesh.coordinates()[model.node_vertex, 2] >= model.S.values[model.node_vertex] - 1e-6
	# docs/get_started.rst:102-104 contour levels for this set-up (full Stokes): 87-92 m/a
	assert 60.0 < ux[top].min() and ux[top].max() < 120.0
	mom2 = MomentumBP(model)
	mom2.solve()
	assert relerr(mom2.get_unknown().values, mom.get_unknown().values) < 1e-12


@pytest.mark.parametrize('name', ['ismip_a_6x5x3', 'marine_6x5x3'])
def test_against_committed_golden_vectors(name):
	"""tests/golden/*.npz (oracle outputs, see make_golden.py) through the C-ABI."""
	import os
	from golden.make_golden import build
	g = np.load(os.path.join(os.path.dirname(__file__), 'golden', name + '.npz'))
	m = build(name)
	c = product_context(m, relative_tolerance=1e-11, maximum_iterations=80, krylov_rtol=1e-13,
	                    relaxation_parameter=float(g['relax']))
	ip, ix = c.csr_pattern()
	assert np.array_equal(ip, g['indptr']) and np.array_equal(ix, g['indices'])
	F, J = c.assemble(g['u'], want_F=True, want_J=True)
	assert relerr(F, g['F']) < TOL_ASM and relerr(J, g['J']) < TOL_ASM
	u = np.full(2 * m.n_nodes, 3e-16)
	st = c.newton_solve(u)
	assert st['converged'] and relerr(u, g['u_newton']) < TOL_U
	assert abs(st['iterations'] - int(g['newton_iterations'])) <= 1
	c.close()


def test_empty_and_ragged_inputs_are_rejected():
	"""C-ABI misuse: integer return + message, no crash (SURVEY §8b errors)."""
	from cslvr_b200 import capi
	m = ismip_model('A', n=(3, 3, 2))
	c = product_context(m)
	with pytest.raises(capi.BPError):
		c.set_field(capi.FIELD_A, np.ones(5))                       # ragged field
	with pytest.raises(capi.BPError):
		c.spmv(np.ones(2 * m.n_nodes)) if False else capi.BPContext(  # empty mesh
			np.zeros((0, 3)), np.zeros((0, 4), dtype=np.int32), 0, ([], [], []), c.params)
	bad = m.cell_dofs.copy()
	bad[0, 4] = bad[0, 0]                                           # u_y dof equal to u_x dof
	with pytest.raises(capi.BPError):
		product_context(m, cell_dofs=bad)
	c2 = capi.BPContext(m.mesh.coordinates(), m.mesh.cells(), 2 * m.n_nodes, (m.facet_cell, m.facet_local, m.ff),
	                    c.params, vertex_to_node=m.vertex_to_node)
	with pytest.raises(capi.BPError):                               # fields not set yet
		c2.assemble(np.zeros(2 * m.n_nodes))
	with pytest.raises(capi.BPError):                               # no resident Jacobian yet
		c.spmv(np.ones(2 * m.n_nodes))
	c.close(); c2.close()


@pytest.mark.parametrize('pc', ['none', 'jacobi', 'block_jacobi', 'chebyshev', 'column', 'mg'])
def test_every_preconditioner_solves_the_same_system(pc):
	"""bp_krylov_solve with each BP_PC_* against a sparse direct solve of the exported J."""
	import scipy.sparse.linalg as spla
	m = ismip_model('A', n=(10, 8, 5))                     # dx = 8 km >> dz = 200 m: strongly anisotropic
	c = product_context(m, krylov_rtol=1e-12, preconditioner=pc, **({} if pc == 'mg' else dict(pc_degree=4)))
	u = state(2 * m.n_nodes, 31)
	F, J = c.assemble(u, want_F=True, want_J=True)
	ip, ix = c.csr_pattern()
	A = sp.csr_matrix((J, ix, ip), shape=(2 * m.n_nodes,) * 2)
	ref = spla.spsolve(A.tocsc(), F)
	dx, its, rr = c.krylov_solve(F)
	assert relerr(dx, ref) < 1e-8, (pc, its, relerr(dx, ref))
	test_every_preconditioner_solves_the_same_system.its = getattr(test_every_preconditioner_solves_the_same_system, 'its', {})
	test_every_preconditioner_solves_the_same_system.its[pc] = its
	c.close()


def test_preconditioner_quality_ordering():
	"""on an anisotropic ice mesh the vertical-line solve must beat block Jacobi by a wide
	margin, and Chebyshev(4) must cut the CG iteration (= all-reduce) count."""
	its = {}
	m = ismip_model('A', n=(12, 12, 5))
	u = state(2 * m.n_nodes, 32)
	for pc in ('block_jacobi', 'chebyshev', 'column'):
		c = product_context(m, krylov_rtol=1e-8, preconditioner=pc, pc_degree=4)
		F, _ = c.assemble(u, want_F=True, want_J=True, J=False)
		_, its[pc], _ = c.krylov_solve(F)
		c.close()
	assert its['column'] * 3 < its['block_jacobi'], its
	assert its['chebyshev'] * 2 < its['block_jacobi'], its


@pytest.mark.parametrize('pc', ['chebyshev', 'column', 'mg'])
def test_newton_with_other_preconditioners(pc):
	m = ismip_model('C', L=20000.0, n=(10, 10, 4))
	P = oracle_problem(m)
	uo, io = P.newton(rtol=1e-11, maxit=40)
	c = product_context(m, relative_tolerance=1e-11, maximum_iterations=40, krylov_rtol=1e-12, preconditioner=pc)
	u = np.full(P.n_dofs, 3e-16)
	st = c.newton_solve(u)
	assert st['converged'] and relerr(u, uo) < TOL_U
	c.close()


@pytest.mark.parametrize('builder', [lambda: ismip_model('A', n=(9, 8, 4)), lambda: marine_model(n=(7, 6, 4))])
def test_solve_pressure_matches_oracle(builder):
	"""SURVEY §8f rank 2, cslvr/momentumbp.py:205-229: L2 projection of the BP pressure."""
	m = builder()
	P = oracle_problem(m)
	u = state(P.n_dofs, 41)
	po = P.pressure(u)                                   # per node (model.Q dofs)
	c = product_context(m)
	pv = c.solve_pressure(u)
	assert relerr(pv[m.node_vertex], po) < TOL_U, relerr(pv[m.node_vertex], po)
	assert np.array_equal(pv, pv[m.node_vertex][m.vertex_to_node])     # periodic images carry the master value
	assert c.last_projection_iterations > 0
	c.close()


def test_solve_runs_pressure_step_like_the_reference():
	from cslvr_b200 import MomentumDukowiczBP
	m = ismip_model('A', n=(8, 8, 4))
	mom = MomentumDukowiczBP(m)
	mom.solve_params['solve_vert_velocity'] = False
	mom.solve()
	P = oracle_problem(m)
	assert relerr(m.p.values, P.pressure(mom.get_unknown().values)) < TOL_U


@pytest.mark.parametrize('builder', [lambda: ismip_model('A', n=(9, 8, 4)), lambda: ismip_model('C', L=20000.0, n=(8, 8, 5)),
                                     lambda: marine_model(n=(7, 6, 4))])
def test_solve_vert_velocity_matches_oracle(builder):
	"""SURVEY §8f rank 1, cslvr/momentumbp.py:231-253: u_z from the continuity system."""
	from cslvr_b200 import capi
	m = builder()
	P = oracle_problem(m)
	u = state(P.n_dofs, 43)
	uzo, uzbo = P.vert_velocity(u, m.vertex_field(m.B))
	c = product_context(m)
	c.set_field(capi.FIELD_B, m.vertex_field(m.B))
	uz = c.solve_vert_velocity(u)
	assert relerr(uz[m.node_vertex], uzo) < TOL_U, (relerr(uz[m.node_vertex], uzo), c.last_vert_iterations)
	c.close()


def test_default_solve_runs_all_three_steps():
	"""MomentumBPBase.solve(): Newton, then u_z, then p (cslvr/momentumbp.py:255-273)."""
	from cslvr_b200 import MomentumDukowiczBP
	m = ismip_model('A', n=(8, 8, 4))
	mom = MomentumDukowiczBP(m)
	mom.solve()
	P = oracle_problem(m)
	uzo, _ = P.vert_velocity(mom.get_unknown().values, m.vertex_field(m.B))
	assert relerr(m.u.sub(2), uzo) < 1e-7 and relerr(m.p.values, P.pressure(mom.get_unknown().values)) < TOL_U


@pytest.mark.parametrize('mode', ['gather', 'scatter'])
def test_linear_solve_matches_oracle(mode):
	"""linear=True -- one solve with eta frozen at eta(model.u)."""
	m = marine_model(n=(7, 6, 4))
	P = oracle_problem(m)
	ul = state(P.n_dofs, 51, 20.0)
	ref = P.linear_solve(ul)
	c = product_context(m, krylov_rtol=1e-13, assembly=mode)
	out = c.linear_solve(ul)
	assert relerr(out, ref) < TOL_U, relerr(out, ref)
	c.close()


def test_linearize_viscosity_like_the_reference():
	"""Momentum.linearize_viscosity (cslvr/momentum.py:97-138): after a non-linear solve, the
	linearised problem at model.u reproduces that velocity (fixed point) in ONE linear solve."""
	from cslvr_b200 import MomentumDukowiczBP
	m = ismip_model('A', n=(8, 8, 4))
	mom = MomentumDukowiczBP(m)
	mom.solve_params['solver']['newton_solver']['relative_tolerance'] = 1e-12
	mom.solve_params['solver']['newton_solver']['krylov_solver'] = {'relative_tolerance': 1e-13}
	mom.solve()
	u_nl = mom.get_unknown().values.copy()
	mom.linearize_viscosity()
	assert mom.linear and mom.solve_params['solve_pressure'] is False and mom.solve_params['solve_vert_velocity'] is False
	mom.solve()
	assert relerr(mom.get_unknown().values, u_nl) < 1e-4       # linearize_viscosity restores the ORIGINAL params: Krylov rtol 1e-5
	P = oracle_problem(m)
	assert relerr(mom.get_unknown().values, P.linear_solve(u_nl)) < 1e-4      # default Krylov rtol 1e-5 on the linear path


@pytest.mark.parametrize('deg', [None, 4])
def test_energy_dependent_rate_factor_at_quadrature_points(deg):
	"""A(T', W, E) evaluated at the quadrature points
	(cslvr/model.py:1815-1859), UFL degree 9 (125-point rule) unless the script sets it
	(simulations/greenland/nioghalvfjerdsbrae/data_assimilation.py:5 uses 4)."""
	from cslvr_b200 import capi
	m = marine_model(n=(6, 5, 3))
	X = m.mesh.coordinates()
	Tp = 250.0 + 20.0 * (1.0 - (X[:, 2] + 500.0) / 1100.0) + 4.0 * np.sin(5 * X[:, 0] / 40000.0)     # straddles 263.15 K
	W = np.clip(0.02 * (Tp - 262.0) / 8.0, 0.0, 0.03)                                                # straddles 0.01
	E = 1.0 + 0.5 * np.cos(3 * X[:, 1] / 40000.0)
	P = oracle_problem(m, energy=dict(Tp=Tp, W=W, E=E), quad_degree=deg)
	kw = dict(energy_dependent_rate_factor=True)
	if deg is not None:
		kw['quadrature_degree'] = deg
	for mode in ('gather', 'scatter'):
		c = product_context(m, assembly=mode, **kw)
		for f, v in ((capi.FIELD_TP, Tp), (capi.FIELD_W, W), (capi.FIELD_E, E)):
			c.set_field(f, v)
		u = state(P.n_dofs, 61)
		F, J = c.assemble(u, want_F=True, want_J=True)
		assert relerr(F, P.residual(u)) < TOL_ASM and relerr(J, P.jacobian(u)) < TOL_ASM
		c.close()


def test_energy_dependent_model_through_the_mirror():
	from cslvr_b200 import MomentumDukowiczBP
	m = ismip_model('A', n=(7, 7, 4))
	X = m.mesh.coordinates()[m.node_vertex]
	m.init_Tp(255.0 + 12.0 * (m.S.values[m.node_vertex] - X[:, 2]) / 1500.0)
	m.init_T(m.Tp.values)
	m.init_W(0.0)
	m.init_E(1.0)
	m.form_energy_dependent_rate_factor()
	mom = MomentumDukowiczBP(m)
	assert mom.solve_params['solver']['newton_solver']['relaxation_parameter'] == 0.7     # non-uniform T: momentumbp.py:169-175
	mom.solve_params['solve_pressure'] = False
	mom.solve_params['solver']['newton_solver']['maximum_iterations'] = 100
	mom.solve_params['solver']['newton_solver']['relative_tolerance'] = 1e-12           # both sides converged tightly:
	mom.solve_params['solver']['newton_solver']['krylov_solver'] = {'relative_tolerance': 1e-13}   # the velocity bar is 1e-8
	mom.solve()
	P = oracle_problem(m, energy=dict(Tp=m.vertex_field(m.Tp), W=m.vertex_field(m.W), E=m.vertex_field(m.E)))
	uo, io = P.newton(rtol=1e-12, relax=0.7, maxit=100)
	assert mom.stats['converged'] and relerr(mom.get_unknown().values, uo) < TOL_U


def test_multigrid_iteration_count_is_nearly_mesh_independent():
	"""the point of BP_PC_MG: refining the footprint 2x must not double the CG iterations."""
	its = {}
	for nxy in (24, 48):
		m = ismip_model('C', L=20000.0 * nxy / 260, n=(nxy, nxy, 5))
		u = state(2 * m.n_nodes, 71, 5.0) + 20.0
		out = {}
		for pc in ('mg', 'chebyshev'):
			c = product_context(m, krylov_rtol=1e-8, preconditioner=pc)
			F, _ = c.assemble(u, want_F=True, want_J=True, J=False)
			_, out[pc], _ = c.krylov_solve(F)
			c.close()
		its[nxy] = out
	assert its[48]['mg'] <= 1.35 * its[24]['mg'] + 3, its
	assert its[48]['mg'] < its[48]['chebyshev'], its


def test_unstructured_extruded_mesh_all_paths():
	"""nothing in the library may depend on BoxMesh structure or vertex numbering: random
	Delaunay footprint, prism -> 3 tets split, shuffled vertex ids, floating tongue, P1 A."""
	from cases import delaunay_extruded_model
	from cslvr_b200 import capi
	m = delaunay_extruded_model()
	assert m.N_GAMMA_B_FLT > 0 and m.N_GAMMA_L_UDR > 0
	P = oracle_problem(m)
	for mode in ('gather', 'scatter'):
		c = product_context(m, assembly=mode)
		ip, ix = c.csr_pattern()
		assert np.array_equal(ip, P.pattern()[0]) and np.array_equal(ix, P.pattern()[1])
		u = state(P.n_dofs, 81)
		F, J = c.assemble(u, want_F=True, want_J=True)
		assert relerr(F, P.residual(u)) < TOL_ASM and relerr(J, P.jacobian(u)) < TOL_ASM
		c.close()
	uo, io = P.newton(rtol=1e-12, relax=0.7, maxit=120)
	for pc in ('mg', 'chebyshev', 'column'):
		c = product_context(m, relative_tolerance=1e-12, relaxation_parameter=0.7, maximum_iterations=120, krylov_rtol=1e-13, preconditioner=pc)
		u = np.full(P.n_dofs, 3e-16)
		st = c.newton_solve(u)
		assert st['converged'] and relerr(u, uo) < TOL_U, (pc, relerr(u, uo))
		c.close()
	c = product_context(m)
	c.set_field(capi.FIELD_B, m.vertex_field(m.B))
	assert relerr(c.solve_pressure(uo)[m.node_vertex], P.pressure(uo)) < TOL_U
	assert relerr(c.solve_vert_velocity(uo)[m.node_vertex], P.vert_velocity(uo, m.vertex_field(m.B))[0]) < 1e-7
	c.close()


@pytest.mark.parametrize('name', ['eismint', 'greenland'])
def test_baseline_configs_3_and_4_at_test_size(name):
	"""BASELINE.json configs[2] (EISMINT-II-style dome, thermomechanical A, frozen bed, no
	lateral pressure) and configs[3] (Greenland-shaped, marine margin, variable beta):
	assembly parity and a converged Newton velocity against the oracle."""
	from cases import eismint_dome_model, greenland_like_model
	m = eismint_dome_model() if name == 'eismint' else greenland_like_model()
	upb = (name != 'eismint')                                  # eismint_ii_a.py: use_pressure_bc=False
	P = oracle_problem(m, use_pressure_bc=upb)
	if name == 'greenland':
		assert m.N_GAMMA_L_UDR > 0 and (P.D > 0).any()         # water pressure on the lateral boundary is active
	c = product_context(m, use_pressure_bc=upb, relative_tolerance=1e-12, relaxation_parameter=0.7, maximum_iterations=120, krylov_rtol=1e-13)
	u = state(P.n_dofs, 91, 5.0)
	F, J = c.assemble(u, want_F=True, want_J=True)
	assert relerr(F, P.residual(u)) < TOL_ASM and relerr(J, P.jacobian(u)) < TOL_ASM
	uo, io = P.newton(rtol=1e-12, relax=0.7, maxit=120)
	un = np.full(P.n_dofs, 3e-16)
	st = c.newton_solve(un)
	assert io['converged'] and st['converged'] and abs(st['iterations'] - io['iterations']) <= 1
	assert relerr(un, uo) < TOL_U, relerr(un, uo)
	c.close()


# ---- SURVEY §8f rank 4 (second half): MomentumDukowiczStokesReduced ------------------------
def _rs_context(m, **params):
	from cslvr_b200 import capi
	c = product_context(m, reduced_stokes=True, **params)
	c.set_field(capi.FIELD_B, m.vertex_field(m.B))
	return c


@pytest.mark.parametrize('builder', [lambda: ismip_model('A', n=(9, 8, 4)), lambda: marine_model(n=(7, 6, 4)),
                                     lambda: ismip_model('A', n=(7, 6, 3), periodic=False)])
def test_reduced_stokes_assembly_matches_oracle(builder):
	"""cslvr/momentumstokes.py:300-400: u_z in the strain-rate tensor, u_z_b in the sliding term
	(grounded bed only), F and J to 1e-12, both assembly modes, sparsity unchanged."""
	from cases import rs_problem
	from cslvr_b200 import capi
	m = builder()
	rng = np.random.default_rng(61)
	Br = rng.normal(size=m.num_vertices) * 0.2
	P = rs_problem(m, B_ring=Br)
	uz = rng.normal(size=P.n_dofs // 2)[m.vertex_to_node] * 4.0       # a P1 function on model.Q, per vertex
	P.set_uz(uz)
	for mode in ('gather', 'scatter'):
		c = _rs_context(m, assembly=mode)
		c.set_field(capi.FIELD_UZ, uz)
		c.set_field(capi.FIELD_B_RING, Br)
		ip, ix = c.csr_pattern()
		assert np.array_equal(ip, P.pattern()[0]) and np.array_equal(ix, P.pattern()[1])
		for seed, scale in ((1, 30.0), (2, 1e-3), (3, 0.0)):
			u = state(P.n_dofs, seed, scale)
			F, J = c.assemble(u, want_F=True, want_J=True)
			assert relerr(F, P.residual(u)) < TOL_ASM, (mode, seed, relerr(F, P.residual(u)))
			assert relerr(J, P.jacobian(u)) < TOL_ASM, (mode, seed, relerr(J, P.jacobian(u)))
			F2, _ = c.assemble(u, want_F=True, want_J=False)
			_, J2 = c.assemble(u, want_F=False, want_J=True)
			assert relerr(F2, F) < 1e-14 and relerr(J2, J) < 1e-14
		# the same context serves the BP forms again once the flag is cleared
		c.set_params(reduced_stokes=False)
		Pb = oracle_problem(m)
		u = state(P.n_dofs, 4)
		F, J = c.assemble(u, want_F=True, want_J=True)
		assert relerr(F, Pb.residual(u)) < TOL_ASM and relerr(J, Pb.jacobian(u)) < TOL_ASM
		c.close()


@pytest.mark.parametrize('builder,cb', [(lambda: ismip_model('A', n=(8, 7, 4)), True), (lambda: marine_model(n=(7, 6, 4)), True),
                                        (lambda: ismip_model('A', n=(8, 7, 4)), False)])
def test_reduced_stokes_newton_matches_oracle(builder, cb):
	"""Model.home_rolled_newton_method (cslvr/model.py:2610-2730) with solve_vert_velocity as the
	callback: same iteration count, residual history, velocity (1e-8) and u_z."""
	from cases import rs_problem
	m = builder()
	P = rs_problem(m)
	uo, io = P.newton_rs(rtol=1e-9, atol=1e-6, maxit=50, relax=0.8, callback=cb)
	c = _rs_context(m, relative_tolerance=1e-9, maximum_iterations=50, relaxation_parameter=0.8, krylov_rtol=1e-13)
	u = np.zeros(P.n_dofs)
	uz = np.zeros(m.num_vertices)
	st = c.rs_newton_solve(u, uz, atol=1e-6, callback=cb)
	assert st['converged'] and io['converged'] and st['iterations'] == io['iterations']
	assert relerr(st['residuals'], io['residuals']) < 1e-6
	assert relerr(u, uo) < TOL_U, relerr(u, uo)
	assert relerr(uz, P.uz) < 1e-7 or not cb
	c.close()


def test_reduced_stokes_class_solves_like_the_reference():
	"""MomentumDukowiczStokesReduced.solve() (cslvr/momentumstokes.py:449-486): Newton with the
	u_z callback, the pressure with the viscosity of the velocity the model held BEFORE the solve,
	then update_model_var."""
	from cases import rs_problem
	from cslvr_b200 import MomentumDukowiczStokesReduced
	m = ismip_model('A', n=(8, 7, 4))
	mom = MomentumDukowiczStokesReduced(m)
	assert mom.solve_params['solver']['newton_solver']['relaxation_parameter'] == 0.8
	assert mom.solve_params['solver']['newton_solver']['maximum_iterations'] == 50
	mom.solve_params['solver']['newton_solver']['krylov_solver'] = {'relative_tolerance': 1e-13}
	mom.solve()
	P = rs_problem(m)
	uo, io = P.newton_rs()
	assert mom.stats['converged'] and mom.stats['iterations'] == io['iterations']
	u = mom.get_unknown().values
	assert relerr(u, uo) < TOL_U
	assert relerr(m.u.sub(0), uo[0::2]) < TOL_U and relerr(m.u.sub(1), uo[1::2]) < TOL_U      # update_model_var
	assert relerr(m.u.sub(2), P.uz[m.node_vertex]) < 1e-7                                         # the callback's u_z
	po = P.pressure_rs(uo, np.zeros(P.n_dofs), P.uz)            # model.u was zero before the solve
	assert relerr(m.p.values, po) < TOL_U, relerr(m.p.values, po)
	# residual / Jacobian of the class (with the u_z it now holds)
	ur = state(P.n_dofs, 7)
	assert relerr(mom.assemble_residual(ur), P.residual(ur)) < TOL_ASM
	ip, ix, J = mom.assemble_jacobian()
	assert relerr(J, P.jacobian(u)) < TOL_ASM


@pytest.mark.parametrize('mode', ['gather', 'scatter'])
def test_reduced_stokes_linear_solve_matches_oracle(mode):
	"""linear=True for the reduced model: viscosity frozen at (u_l, u_z), the right-hand side
	keeps the u_z part of d(epsdot) and the B_ring part of the sliding term."""
	from cases import rs_problem
	from cslvr_b200 import capi
	m = marine_model(n=(7, 6, 4))
	rng = np.random.default_rng(71)
	Br = rng.normal(size=m.num_vertices) * 0.2
	P = rs_problem(m, B_ring=Br)
	uz = rng.normal(size=P.n_dofs // 2)[m.vertex_to_node] * 2.0
	P.set_uz(uz)
	ul = state(P.n_dofs, 51, 20.0)
	ref = P.linear_solve(ul)
	c = _rs_context(m, krylov_rtol=1e-13, assembly=mode)
	c.set_field(capi.FIELD_UZ, uz)
	c.set_field(capi.FIELD_B_RING, Br)
	out = c.linear_solve(ul)
	assert relerr(out, ref) < TOL_U, relerr(out, ref)
	c.close()


@pytest.mark.parametrize('name', ['rs_ismip_a_6x5x3', 'rs_marine_6x5x3'])
def test_reduced_stokes_against_committed_golden_vectors(name):
	"""the CUDA path against tests/golden/rs_*.npz (oracle outputs, see make_golden.py::main_rs)."""
	import os
	from cslvr_b200 import capi
	from golden.make_golden import build
	g = np.load(os.path.join(os.path.dirname(__file__), 'golden', name + '.npz'))
	m = build(name[3:])
	c = _rs_context(m, relative_tolerance=1e-9, maximum_iterations=50, relaxation_parameter=0.8, krylov_rtol=1e-13)
	c.set_field(capi.FIELD_UZ, g['uz'])
	c.set_field(capi.FIELD_B_RING, g['B_ring'])
	F, J = c.assemble(g['u'], want_F=True, want_J=True)
	assert relerr(F, g['F']) < TOL_ASM and relerr(J, g['J']) < TOL_ASM
	u, uz = np.zeros(g['u'].size), np.zeros(g['uz'].size)
	st = c.rs_newton_solve(u, uz, atol=1e-6, callback=True)
	assert st['converged'] and st['iterations'] == int(g['newton_iterations'])
	assert relerr(st['residuals'], g['residuals']) < 1e-6
	assert relerr(u, g['u_newton']) < TOL_U and relerr(uz, g['uz_newton']) < 1e-7
	c.close()


def test_reduced_stokes_entry_points_refuse_misuse():
	"""bp_rs_newton_solve / bp_rs_solve_pressure: the model flag and the bed field are required; an
	in-process partition is refused (one rank per process serves partitioned contexts)."""
	from cslvr_b200 import capi
	from cslvr_b200.capi import BPError
	from cslvr_b200.partition import partition_model, rank_context
	m = ismip_model('A', n=(6, 5, 3))
	c = product_context(m)                               # Blatter-Pattyn context: flag not set
	u, uz = np.zeros(2 * m.n_nodes), np.zeros(m.num_vertices)
	with pytest.raises(BPError) as e:
		c.rs_newton_solve(u, uz)
	assert e.value.code == capi.BP_ERR_STATE and 'reduced_stokes' in str(e.value)
	with pytest.raises(BPError):
		c.rs_solve_pressure(u, u)
	c.set_params(reduced_stokes=True)
	with pytest.raises(BPError) as e:                    # solve_vert_velocity callback needs B
		c.rs_newton_solve(u, uz)
	assert 'BP_FIELD_B' in str(e.value)
	st = c.rs_newton_solve(u, uz, callback=False)        # cb_ftn = None works without it
	assert st['converged'] and np.all(uz == 0.0)
	with pytest.raises(ValueError):
		c.rs_newton_solve(u.astype(np.float32), uz)
	c.close()
	rm = partition_model(m, 2, 0)
	cp = rank_context(m, rm, reduced_stokes=True)
	with pytest.raises(BPError) as e:
		cp.rs_newton_solve(np.zeros(2 * rm.n_nodes), np.zeros(rm.coords.shape[0]), callback=False)
	assert 'communicator' in str(e.value)
	cp.close()


@pytest.mark.parametrize('case', ['ismip', 'marine', 'delaunay'])
def test_tiled_cell_kernel_matches_oracle_and_row_gather(case):
	"""assembly='tiled' (one thread per prism column, every tet evaluated once; the default on extruded meshes)
	against the oracle to 1e-12 and against the row gather, bit-for-bit reproducible from run to run."""
	from cases import delaunay_extruded_model
	m = {'ismip': lambda: ismip_model('A', n=(40, 33, 4)), 'marine': lambda: marine_model(n=(21, 17, 5)),
	     'delaunay': lambda: delaunay_extruded_model(n_pts=300, n_layers=5)}[case]()
	P = oracle_problem(m)
	u = state(P.n_dofs, 7)
	c = product_context(m, assembly='tiled')
	F1, J1 = c.assemble(u, want_F=True, want_J=True)
	F2, J2 = c.assemble(u, want_F=True, want_J=True)
	c.close()
	assert np.array_equal(F1, F2) and np.array_equal(J1, J2)
	assert relerr(F1, P.residual(u)) < TOL_ASM and relerr(J1, P.jacobian(u)) < TOL_ASM
	c = product_context(m, assembly='gather')
	F3, J3 = c.assemble(u, want_F=True, want_J=True)
	c.close()
	assert relerr(F1, F3) < TOL_ASM and relerr(J1, J3) < TOL_ASM


def test_default_assembly_is_atomic_free_and_bitwise_reproducible(monkeypatch):
	"""the default path has no atomics anywhere: cell integrals by row gather (or the tiled cell kernel),
	facet integrals by row gather (one thread per touched node, plain read-modify-write of its own rows
	after the cell kernel) -- F and J identical bit for bit from run to run, with no environment switch;
	BP_FACET_ATOMIC=1 gives the atomic facet kernel back (same numbers to rounding)."""
	m = marine_model()
	P = oracle_problem(m)
	u = state(P.n_dofs, 5)
	monkeypatch.delenv('BP_FACET_ATOMIC', raising=False)
	c = product_context(m)
	F1, J1 = c.assemble(u, want_F=True, want_J=True)
	F2, J2 = c.assemble(u, want_F=True, want_J=True)
	assert np.array_equal(F1, F2) and np.array_equal(J1, J2)
	assert relerr(F1, P.residual(u)) < TOL_ASM and relerr(J1, P.jacobian(u)) < TOL_ASM
	c.close()
	monkeypatch.setenv('BP_FACET_ATOMIC', '1')
	c = product_context(m)
	F3, J3 = c.assemble(u, want_F=True, want_J=True)
	assert relerr(F3, F1) < TOL_ASM and relerr(J3, J1) < TOL_ASM
	c.close()


def test_reduced_stokes_with_the_energy_dependent_rate_factor():
	"""the two model switches together: A(T', W, E) at the quadrature points (cslvr/model.py:1815-1859)
	inside the reduced-Stokes forms (cslvr/momentumstokes.py:300-400), generic Glen exponent included."""
	from cases import rs_problem
	from cslvr_b200 import capi
	m = marine_model(n=(6, 5, 3))
	X = m.mesh.coordinates()
	Tp = 250.0 + 20.0 * (1.0 - (X[:, 2] + 500.0) / 1100.0) + 4.0 * np.sin(5 * X[:, 0] / 40000.0)
	W = np.clip(0.02 * (Tp - 262.0) / 8.0, 0.0, 0.03)
	E = 1.0 + 0.5 * np.cos(3 * X[:, 1] / 40000.0)
	rng = np.random.default_rng(81)
	uz = rng.normal(size=m.n_nodes)[m.vertex_to_node] * 3.0
	P = rs_problem(m, energy=dict(Tp=Tp, W=W, E=E), quad_degree=4)
	P.set_uz(uz)
	c = _rs_context(m, energy_dependent_rate_factor=True, quadrature_degree=4)
	for f, v in ((capi.FIELD_TP, Tp), (capi.FIELD_W, W), (capi.FIELD_E, E), (capi.FIELD_UZ, uz)):
		c.set_field(f, v)
	u = state(P.n_dofs, 62)
	F, J = c.assemble(u, want_F=True, want_J=True)
	assert relerr(F, P.residual(u)) < TOL_ASM and relerr(J, P.jacobian(u)) < TOL_ASM
	c.close()


def test_newton_with_lateral_dirichlet_conditions_matches_oracle():
	"""MomentumBP(use_lat_bcs=True) on a mesh with basin divides (cslvr/momentumbp.py:101-112, d3model.py:441-448): the
	Newton solve with DirichletBCs on GAMMA_L_DVD -- dolfin applies them to the update (SURVEY A-N) -- against the
	oracle's constrained Newton (identity rows, sparse LU): same iteration count, velocity to 1e-8, constrained dofs
	exact; the linear=True solve obeys the conditions as well."""
	from cslvr_b200 import MomentumBP
	m = marine_model(n=(10, 8, 4))
	X = m.mesh.coordinates()
	L = X[:, 0].max()
	contour = np.column_stack([np.full(400, L), np.linspace(X[:, 1].min(), X[:, 1].max(), 400)])
	m.calculate_boundaries(mask=m.mask_callable, mark_divide=True, contour=contour)
	m.init_u_lat(lambda x, y, z: 40.0 + 10.0 * np.sin(3 * y / L))
	m.init_u_y_lat(lambda x, y, z: 5.0 * np.cos(2 * x / L))
	mom = MomentumBP(m, use_lat_bcs=True)
	ns = mom.solve_params['solver']['newton_solver']
	ns.update(relative_tolerance=1e-11, maximum_iterations=60, relaxation_parameter=0.7)
	ns['krylov_solver'] = {'relative_tolerance': 1e-13, 'monitor_convergence': False}
	mom.solve_params['solve_vert_velocity'] = False
	mom.solve_params['solve_pressure'] = False
	mom.solve()
	u = mom.get_unknown().values.copy()
	bcs = mom.get_boundary_conditions()
	dofs = np.concatenate([b[0] for b in bcs])
	vals = np.concatenate([b[1].values[b[2]] for b in bcs])
	assert dofs.size > 0
	P = oracle_problem(m)
	uo, io = P.newton(rtol=1e-11, maxit=60, relax=0.7, bcs=(dofs, vals))
	assert io['converged'] and abs(io['iterations'] - mom.stats['iterations']) <= 1
	assert relerr(u, uo) < 1e-8
	assert np.abs(u[dofs] - vals).max() < 1e-9 * np.abs(vals).max()
	# without the conditions the solution is a different one (the test has teeth)
	un, _ = P.newton(rtol=1e-11, maxit=60, relax=0.7)
	assert relerr(un, uo) > 1e-3
	# u_z with its lateral condition DirichletBC(Q, u_z_lat, ff, GAMMA_L_DVD) on top of the basal ones (momentumbp.py:110-111)
	m.init_u_z_lat(lambda x, y, z: -0.3 + 1e-5 * x)
	mom.solve_vert_velocity()
	nodes = m.divide_nodes()
	uz_o, _ = P.vert_velocity(u, m.vertex_field(m.B), lateral_bc=(nodes, m.u_z_lat.values[nodes]))
	assert relerr(mom.u_z.values, uz_o) < 1e-7
	assert np.abs(mom.u_z.values[nodes] - m.u_z_lat.values[nodes]).max() < 1e-9
	# linear=True: K(u_lin) u = f with the same rows constrained
	m.u.values[0::3], m.u.values[1::3] = u[0::2], u[1::2]
	lin = MomentumBP(m, use_lat_bcs=True, linear=True)
	lin.solve_params['solve_vert_velocity'] = False
	lin.solve_params['solve_pressure'] = False
	lin.solve_params['solver']['krylov_solver'] = {'relative_tolerance': 1e-13}      # the rows are only as exact as the Krylov solve
	lin.solve()
	ul = lin.get_unknown().values
	assert np.abs(ul[dofs] - vals).max() < 1e-9 * np.abs(vals).max()


def test_krylov_tolerance_change_is_honoured_by_the_iteration_graph():
	"""the CG iteration replays as a CUDA graph whose kernels carry the tolerances as arguments: a changed krylov_rtol on
	the same context must rebuild it (PETSc reads the tolerances at every KSPSolve)."""
	m = ismip_model('A', n=(16, 16, 4))
	c = product_context(m, krylov_rtol=1e-3, preconditioner='mg')
	u = state(2 * m.n_nodes, 5)
	F, _ = c.assemble(u, want_F=True, want_J=True, J=False)
	_, it1, rr1 = c.krylov_solve(F)
	_, it1b, rr1b = c.krylov_solve(F)                    # second solve: the graph exists now
	c.set_params(krylov_rtol=1e-11, preconditioner='mg')
	_, it2, rr2 = c.krylov_solve(F)
	assert rr1 < 1e-3 and rr1b < 1e-3 and it1b == it1
	assert rr2 < 1e-11 and it2 > it1
	c.close()
