"""gcc recipe for the oracle's C twin -> oracle/_c/libbp_oracle.so (TEST INFRASTRUCTURE).

The reference itself is pure Python over an un-vendored FEniCS stack,
so there is nothing under /root/reference to compile into oracle/_ref; the CPU baseline
kind is therefore "port"."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'bp_oracle_c.c')
LIB = os.path.join(HERE, '_c', 'libbp_oracle.so')


def build(force=False):
	if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
		return LIB
	os.makedirs(os.path.dirname(LIB), exist_ok=True)
	cmd = ['gcc', '-O3', '-march=x86-64-v3', '-fopenmp', '-fPIC', '-shared', '-o', LIB, SRC, '-lm']
	r = subprocess.run(cmd, capture_output=True, text=True)
	if r.returncode != 0:
		raise RuntimeError('gcc failed on the oracle: ' + r.stderr)
	return LIB


if __name__ == '__main__':
	print(build(force=True))
