"""Build libcslvr_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, 'csrc', f) for f in
       ('bp_setup.cu', 'bp_kernels.cu', 'bp_tiles.cu', 'bp_mg.cu', 'bp_comm.cu', 'bp_api.cu')]
HDR = [os.path.join(HERE, 'csrc', 'bp_internal.h'), os.path.join(ROOT, 'include', 'cslvr_b200.h')]
LIB = os.environ.get('BP_LIB_PATH') or os.path.join(HERE, 'lib', 'libcslvr_b200.so')


def lib_path():
	return LIB


def needs_build():
	if not os.path.exists(LIB):
		return True
	t = os.path.getmtime(LIB)
	return any(os.path.getmtime(f) > t for f in SRC + HDR)


def build(force=False, verbose=False):
	"""nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> cslvr_b200/lib/libcslvr_b200.so"""
	if not force and not needs_build():
		return LIB
	os.makedirs(os.path.dirname(LIB), exist_ok=True)
	nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
	cmd = [nvcc, '-O3', '-std=c++17', '-lineinfo',
	       '-gencode', 'arch=compute_100a,code=sm_100a',
	       '-Xcompiler', '-fPIC,-fopenmp,-O3', '-shared',
	       '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(HERE, 'csrc'),
	       '-Xptxas', '-v' if verbose else '-O3'] + (['-DBP_EXPERIMENTS'] if os.environ.get('BP_EXPERIMENTS') == '1' else []) + os.environ.get('BP_NVCC_FLAGS', '').split() + [
	       '-o', LIB] + SRC + ['-lcudart', '-ldl', '-lgomp']
	r = subprocess.run(cmd, capture_output=True, text=True)
	if r.returncode != 0:
		sys.stderr.write(r.stdout + r.stderr)
		raise RuntimeError('nvcc failed building libcslvr_b200.so')
	if verbose:
		sys.stderr.write(r.stdout + r.stderr)
	return LIB


if __name__ == '__main__':
	print(build(force=True, verbose='-v' in sys.argv))
