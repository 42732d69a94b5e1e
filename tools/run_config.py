#!/usr/bin/env python
"""Solve one BASELINE.json configuration at a chosen size on one GPU and print the solver
statistics (not a bench line: set-up is numpy, the solve is the library).

    python tools/run_config.py eismint 577 577 5        # config 3, 9.99 M tets
    python tools/run_config.py greenland 260 520 5
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ.setdefault('CSLVR_B200_QUIET', '1')


def main():
	from cases import eismint_dome_model, greenland_like_model, product_context
	name = sys.argv[1]
	n = tuple(int(a) for a in sys.argv[2:5])
	t0 = time.time()
	m = eismint_dome_model(n=n) if name == 'eismint' else greenland_like_model(n=n)
	t_model = time.time() - t0
	t0 = time.time()
	c = product_context(m, use_pressure_bc=(name != 'eismint'), relative_tolerance=1e-9, relaxation_parameter=0.7,
	                    maximum_iterations=60, krylov_rtol=1e-5)
	t_ctx = time.time() - t0
	u = np.full(2 * m.n_nodes, 3e-16)
	t0 = time.time()
	st = c.newton_solve(u)
	t_solve = time.time() - t0
	ms = c.time_assemble(3, repeats=10, flush_l2=True)
	out = dict(config=name, mesh=n, n_tets=m.num_cells, n_dofs=2 * m.n_nodes, host_model_s=t_model, context_setup_s=t_ctx,
	           newton_solve_s=t_solve, converged=st['converged'], newton_iterations=st['iterations'],
	           krylov_iterations=st['krylov_iterations'], assemble_ms_in_solve=st['t_assemble_ms'], krylov_ms=st['t_krylov_ms'],
	           assembly_pass_ms=ms, assembly_mtets_s=m.num_cells / ms / 1e3, preconditioner=c.pc_label(),
	           u_max=float(np.abs(u).max()), residual0=st['residual_norm0'], residual=st['residual_norm'])
	print(json.dumps(out))


if __name__ == '__main__':
	main()
