import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); os.environ['CSLVR_B200_QUIET']='1'
import numpy as np, torch, bench
from cases import product_context
m = bench.build_model(1); c = product_context(m)
u = bench.synthetic_state(m)
u_pin = torch.from_numpy(u).pin_memory(); F_pin = torch.empty(u.size, dtype=torch.float64).pin_memory()
un, Fn = u_pin.numpy(), F_pin.numpy()
for k in sys.argv[1:]:
    os.environ['BP_PIPE_K'] = k
    for _ in range(5): c.assemble(un, want_F=True, want_J=True, F=Fn, J=False)
    t0 = time.perf_counter()
    for _ in range(200): c.assemble(un, want_F=True, want_J=True, F=Fn, J=False)
    t = (time.perf_counter() - t0) / 200
    print('K', k, 'ms/step %.4f' % (t * 1e3), 'Mtets/s %.0f' % (m.num_cells / t / 1e6), flush=True)
