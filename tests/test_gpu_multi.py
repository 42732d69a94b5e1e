"""The NCCL path under pytest: on a box with >= 2 GPUs this launches tests/multi_gpu_check.py as one rank per GPU
(torch.distributed.run, rendezvous on 127.0.0.1) and requires its "OK": Newton, u_z, p and the reduced-Stokes solve
of 2 ranks against the single-rank result to 1e-8.  Skipped on single-GPU boxes (the in-process group tests of
test_gpu_partition.py cover the partitioned code path there)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_ranks_over_nccl_equal_one_rank():
	import torch
	if torch.cuda.device_count() < 2:
		pytest.skip("needs >= 2 GPUs")
	cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
	       '--master-port', '29533', os.path.join(HERE, 'multi_gpu_check.py')]
	r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
	assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
	assert r.stdout.strip().endswith("OK"), r.stdout[-3000:]
