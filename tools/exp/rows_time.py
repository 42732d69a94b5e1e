# time the cell-assembly kernel alone on the bench workload (2.03 M tets) and print checksums of F / J
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); os.environ['CSLVR_B200_QUIET'] = '1'
import numpy as np, bench
from cases import product_context
m = bench.build_model(1); c = product_context(m)
u = bench.synthetic_state(m)
F, _ = c.assemble(u, want_F=True, want_J=True, J=False)
import torch
y = c.spmv(np.ones(2 * m.n_nodes))
print('checksum F %.15e  J.1 %.15e' % (np.abs(F).sum(), np.abs(y).sum()))
for rep in range(3):
    print('rows  %.4f ms   pass %.4f ms' % (c.time_assemble(3 | 4, repeats=20, flush_l2=True), c.time_assemble(3, repeats=20, flush_l2=True)))
c.assemble(u, want_F=True, want_J=True, J=False)
print('spmv  %.4f ms' % c.time_spmv(repeats=50, flush_l2=True))
