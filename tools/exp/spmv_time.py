import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); os.environ['CSLVR_B200_QUIET']='1'
import numpy as np, bench
from cases import product_context
m = bench.build_model(1); c = product_context(m)
u = bench.synthetic_state(m)
c.assemble(u, want_F=True, want_J=True, J=False)
for fl in (1, 3, 0, 2):
    print('flush', fl&1, 'with_dot', fl>>1, c.time_spmv(repeats=50, flush_l2=fl), 'ms')
