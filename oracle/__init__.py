"""CPU oracle for the Blatter-Pattyn momentum hot path of pf4d/cslvr.

TEST INFRASTRUCTURE ONLY.  Nothing in ``cslvr_b200/`` (the product) may import,
link or execute anything in this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs do, and there only as the checker / the reported CPU baseline.

PARITY UNPINNED: the reference (cslvr on FEniCS 2017.2.0) cannot be imported in
this container and ships no tests or golden vectors for this path
(SURVEY.md §4, §8c).  The oracle is therefore a restatement of the math written
in ``cslvr/momentumbp.py``, ``cslvr/momentum.py``, ``cslvr/physics.py`` and
``cslvr/d3model.py`` plus the recalled behaviour of the un-vendored third-party
stack; it earns trust through the self-consistency
checks in ``tests/test_oracle_*.py`` (sympy re-derivation of the UFL form,
finite differences, symmetry/SPD, quadrature exactness, MMS-free sanity ranges).
"""
