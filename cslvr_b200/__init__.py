"""cslvr_b200 -- B200-native Blatter-Pattyn momentum path behind cslvr's Physics API.

``from cslvr_b200 import *`` gives a cslvr BP script the names it uses
(docs/get_started.rst:11-50): Point, BoxMesh, D3Model, MomentumBP,
MomentumDukowiczBP (and MomentumDukowiczStokesReduced, the reduced full-Stokes model on
the same unknown).  The numerics run in ``lib/libcslvr_b200.so`` (hand-written
sm_100a CUDA, reached through ctypes); there is no CPU fallback.
"""
from .mesh import Point, Mesh, BoxMesh
from .d3model import D3Model, Function, DOLFIN_EPS
from .momentumbp import MomentumBPBase, MomentumBP, MomentumDukowiczBP
from .momentumstokes import MomentumDukowiczStokesReduced
from .inputoutput import print_text, print_min_max

__all__ = ['Point', 'Mesh', 'BoxMesh', 'D3Model', 'Function', 'DOLFIN_EPS',
           'MomentumBPBase', 'MomentumBP', 'MomentumDukowiczBP', 'MomentumDukowiczStokesReduced',
           'print_text', 'print_min_max']
