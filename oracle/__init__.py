"""ORACLE — test infrastructure only.

CPU restatements of the reference (ZihaoW123/UniMM) algorithms on the generative-scoring hot path.
Nothing under ``unimm_b200/`` imports this package; it exists so that ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs can check and
time the CUDA path against an independent implementation.  See each module's header for the
reference file:line it follows and for how it is pinned (tests/golden/).
"""
