# needs the library built with BP_EXPERIMENTS=1 (python cslvr_b200/_build.py) for the proxy kernel (flag 8)
import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); os.environ['CSLVR_B200_QUIET']='1'
import numpy as np, bench
from cases import product_context
m = bench.build_model(1); c = product_context(m)
u = bench.synthetic_state(m)
c.assemble(u, want_F=True, want_J=True, J=False)
for name, what in (('rows',3|4),('proxy_smem_cas',3|8)):
    for rep in range(2):
        print(name, c.time_assemble(what, repeats=10, flush_l2=True), 'ms')
c.set_params(assembly='scatter')
print('scatter', c.time_assemble(3|4, repeats=10, flush_l2=True))
