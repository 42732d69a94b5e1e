"""Time the host-buffer bp_assemble (u pinned in, F pinned out, J resident) on an ISMIP-HOM A mesh.

    python tools/prof_e2e.py [nxy=408] [steps=20]        (BP_ASM_TILED=0 / BP_NO_PIPELINE=1 / BP_PIPE_K=k select the path)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('CSLVR_B200_QUIET', '1')
import bench
from cslvr_b200 import capi

nxy = int(sys.argv[1]) if len(sys.argv) > 1 else 408
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
model = bench.build_model(1, nxy=nxy)
s = bench.slab_from_model(model)
u = bench.synthetic_state(s)
ctx = capi.context_from_model(model, verbose=True)
u_pin = torch.from_numpy(u).pin_memory()
F_pin = torch.empty(u.size, dtype=torch.float64).pin_memory()
F_ref, _ = ctx.assemble(torch.from_numpy(u).cuda(), want_F=True, want_J=True, J=False)
for _ in range(3):
	ctx.assemble(u_pin.numpy(), want_F=True, want_J=True, F=F_pin.numpy(), J=False)
t0 = time.perf_counter()
for _ in range(steps):
	ctx.assemble(u_pin.numpy(), want_F=True, want_J=True, F=F_pin.numpy(), J=False)
t = (time.perf_counter() - t0) / steps
err = float(np.abs(F_pin.numpy() - F_ref.cpu().numpy()).max() / np.abs(F_ref.cpu().numpy()).max())
print("e2e %.4f ms/step  %.1f Mtets/s   |F_host - F_dev| %.1e   env: tiled=%s nopipe=%s K=%s" % (
	t * 1e3, model.num_cells / t / 1e6, err, os.environ.get('BP_ASM_TILED', '-'), os.environ.get('BP_NO_PIPELINE', '-'), os.environ.get('BP_PIPE_K', '-')), flush=True)
