"""CPU oracle for the RAFT correlation hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker (or the timed CPU baseline),
never as the thing shipped.  The product path (``optical-flow-dnn_b200``) never
imports this package and fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``oracle.corr_oracle`` (numpy) and ``oracle/corr_oracle.c``
(plain C) are checked in ``tests/test_oracle_golden.py`` against golden vectors
produced by importing the reference's own ``raft/core/corr.py`` in the build
container (``tests/golden/make_golden.py``; the reference itself has no tests or
fixtures for this path).
"""
