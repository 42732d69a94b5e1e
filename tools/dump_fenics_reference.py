#!/usr/bin/env python
"""Run ON A FEniCS 2017.2 BOX WITH cslvr INSTALLED (not possible in the build container):
dumps everything needed to pin parity against the real reference as one .npz --
mesh, dof map, markers, coefficient vertex values, a velocity state, assemble(F),
assemble(J) in CSR form, and the converged Newton solution.

    python tools/dump_fenics_reference.py ismip_a.npz

tests/test_gpu_parity.py::test_against_committed_golden_vectors accepts such a file
unchanged (same keys as tests/golden/make_golden.py plus the mesh arrays)."""
import sys

import numpy as np


def main(out):
	from cslvr import D3Model, MomentumDukowiczBP          # noqa: needs FEniCS 2017.2
	from dolfin import Point, BoxMesh, Expression, assemble, as_backend_type, parameters, pi
	parameters['form_compiler']['quadrature_degree'] = 5
	L = 80000.0
	mesh = BoxMesh(Point(0.0, 0.0, 0.0), Point(L, L, 1), 15, 15, 5)
	model = D3Model(mesh, out_dir='./dump/', use_periodic=True)
	model.calculate_boundaries()
	a = 0.5 * pi / 180
	S = Expression('- x[0] * tan(a)', a=a, element=model.Q.ufl_element())
	B = Expression('- x[0] * tan(a) - 1000.0 + 500.0 * sin(2*pi*x[0]/L) * sin(2*pi*x[1]/L)', a=a, L=L,
	               element=model.Q.ufl_element())
	model.deform_mesh_to_geometry(S, B)
	model.init_beta(1000)
	model.init_A(1e-16)
	mom = MomentumDukowiczBP(model)
	Q = mom.Q
	dm = Q.dofmap()
	cell_dofs = np.array([dm.cell_dofs(c) for c in range(mesh.num_cells())], dtype=np.int32)
	u = mom.get_unknown()
	rng = np.random.default_rng(99)
	u0 = rng.normal(size=Q.dim()) * 30.0
	u.vector().set_local(u0); u.vector().apply('insert')
	F = assemble(mom.get_residual()).get_local()
	J = as_backend_type(assemble(mom.get_jacobian())).mat()
	ip, ix, jv = J.getValuesCSR()
	mesh.init(2, 3)
	ff = model.ff.array()
	idx = np.flatnonzero(ff > 0)
	f2c, c2f = mesh.topology()(2, 3), mesh.topology()(3, 2)
	fcell = np.array([f2c(int(f))[0] for f in idx], dtype=np.int32)
	flocal = np.array([list(c2f(int(c))).index(int(f)) for c, f in zip(fcell, idx)], dtype=np.int32)
	mom.solve_params['solver']['newton_solver'].update(relative_tolerance=1e-11, maximum_iterations=80)
	mom.solve_params['solve_vert_velocity'] = False
	mom.solve_params['solve_pressure'] = False
	mom.solve()
	extra = dump_reduced_stokes(model, u0)
	np.savez_compressed(out, coords=mesh.coordinates(), cells=mesh.cells(), cell_dofs=cell_dofs,
	                    facet_cell=fcell, facet_local=flocal, facet_marker=ff[idx].astype(np.int32),
	                    S=model.S.compute_vertex_values(mesh), A=model.A.compute_vertex_values(mesh),
	                    beta=model.beta.compute_vertex_values(mesh), B=model.B.compute_vertex_values(mesh),
	                    u=u0, F=F, indptr=ip, indices=ix, J=jv, u_newton=u.vector().get_local(), relax=1.0, **extra)


def dump_reduced_stokes(model, u0):
	"""the same for MomentumDukowiczStokesReduced (cslvr/momentumstokes.py:300-507): F and J at u0 with a
	random frozen u_z, then the class's own solve() (home-rolled Newton with the u_z callback); keys match
	tests/golden/make_golden.py::main_rs (rs_F, rs_J, rs_uz, rs_u_newton, rs_uz_newton, rs_newton_iterations)."""
	from cslvr import MomentumDukowiczStokesReduced
	from dolfin import assemble, as_backend_type
	mom = MomentumDukowiczStokesReduced(model)
	mesh = model.mesh
	rng = np.random.default_rng(7)
	uz0 = rng.normal(size=model.Q.dim()) * 3.0
	mom.u_z.vector().set_local(uz0); mom.u_z.vector().apply('insert')
	u = mom.get_unknown()
	u.vector().set_local(u0); u.vector().apply('insert')
	F = assemble(mom.get_residual()).get_local()
	ip, ix, jv = as_backend_type(assemble(mom.get_jacobian())).mat().getValuesCSR()
	uz_vertex = mom.u_z.compute_vertex_values(mesh)
	u.vector().zero(); mom.u_z.vector().zero()
	mom.solve_params['solve_pressure'] = False
	mom.solve()
	return dict(rs_F=F, rs_J=jv, rs_uz=uz_vertex, rs_u_newton=u.vector().get_local(),
	            rs_uz_newton=mom.u_z.compute_vertex_values(mesh))


if __name__ == '__main__':
	main(sys.argv[1] if len(sys.argv) > 1 else 'fenics_reference.npz')
