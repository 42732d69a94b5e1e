import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); os.environ['CSLVR_B200_QUIET'] = '1'
import numpy as np
from cases import ismip_model, product_context, state
m = ismip_model('A')
for mode in ('scatter', 'gather'):
    c = product_context(m, assembly=mode)
    u = state(2 * m.n_nodes, 1)
    for kw in (dict(want_F=True, want_J=False), dict(want_F=False, want_J=True), dict(want_F=True, want_J=True)):
        try:
            F, J = c.assemble(u, **kw)
            print(mode, kw, 'ok')
        except Exception as e:
            print(mode, kw, 'FAIL', str(e)[:200]); break
